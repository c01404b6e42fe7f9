// Gradient all-reduce over NVLink peer memory.  One process per GPU; every rank owns an exchange buffer and
// two flag arrays in cudaMalloc memory that the other ranks map through CUDA IPC.  An all-reduce of n floats is three launches on the
// caller's stream, no host synchronisation and no NCCL call:
//   publish : grads -> own exchange buffer; the last CTA raises flagA[rank] = epoch in every peer              (st.release.sys)
//   reduce  : waits until every peer's flagA reached the epoch, then rank r sums slice r of ALL exchange buffers in rank order (P2P
//             loads through NVLink / NVSwitch) and stores the sum into slice r of every rank's buffer (P2P stores); last CTA raises flagB
//   collect : waits for every peer's flagB, copies the exchange buffer (now the full sum) back into grads
// Every element is summed by exactly one rank in a fixed order, so all replicas receive bit-identical gradients (NCCL's ring / tree
// order is not specified).  When a rank passes collect, every peer has finished reading this rank's buffer and writing into it, so
// the next publish may overwrite it.  A spin that exceeds ~20 s (a rank that never arrives) sets an error word instead of hanging the GPU.
//
// Registered, in-place variant (the one training uses): dtc_dp_register maps every rank's GRADIENT BUFFER itself into the others
// (IPC handle of the allocation that contains it + offset), and the all-reduce is ONE cooperative kernel with no staging copies:
//   flagA handshake (every rank's backward has finished) -> rank r sums slice r straight out of all ranks' gradient buffers and stores
//   the sum straight into all of them -> grid barrier -> flagB handshake (every slice has landed everywhere).
#include <cuda.h>

#include "dtc_common.cuh"

#define DP_MAX_WORLD 16
#define DP_BLOCKS 128
#define DP_THREADS 256

struct dtc_dp {
  int rank, world;
  int64_t n;                       // floats in every exchange buffer
  float* buf[DP_MAX_WORLD];        // [rank] = own (cudaMalloc), others = IPC mappings
  uint32_t* flags[DP_MAX_WORLD];   // per rank: flagA[DP_MAX_WORLD] | flagB[DP_MAX_WORLD] | done counter | error word
  bool opened[DP_MAX_WORLD];
  uint32_t epoch;
  // registered in-place range: reg[rank] = local base (caller's memory), others = IPC mappings of the peers' ranges
  float* reg[DP_MAX_WORLD];
  void* reg_map[DP_MAX_WORLD];     // what cudaIpcOpenMemHandle returned (allocation base), for closing
  int64_t reg_n;
  bool reg_open[DP_MAX_WORLD];
  int coop_blocks;
  uint32_t bar_target;
};
struct DpPeers {
  float* buf[DP_MAX_WORLD];
  uint32_t* flags[DP_MAX_WORLD];
};
#define DP_FLAG_A 0
#define DP_FLAG_B DP_MAX_WORLD
#define DP_COUNTER (2 * DP_MAX_WORLD)
#define DP_ERROR (2 * DP_MAX_WORLD + 1)
#define DP_GRIDBAR (2 * DP_MAX_WORLD + 2)  // arrival count of the in-place kernel's grid barrier (monotonic)
#define DP_FLAG_WORDS (2 * DP_MAX_WORLD + 3)

__device__ __forceinline__ void dp_store_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t dp_load_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 dp_load_peer(const float* p) {  // never served from a stale L1 line of an earlier epoch
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// thread 0 of every CTA waits until all `world` flags of the own array reached `epoch`, then the CTA proceeds
__device__ __forceinline__ void dp_wait_all(uint32_t* own_flags, int which, int world, uint32_t epoch) {
  if ((int)threadIdx.x < world) {  // one thread per peer: the polls run side by side
    const long long t0 = clock64();
    while ((int32_t)(dp_load_acquire_sys(own_flags + which + threadIdx.x) - epoch) < 0) {
      if (clock64() - t0 > 40000000000ll) { own_flags[DP_ERROR] = 1u; break; }  // ~20 s: give up rather than hang the device
      __nanosleep(20);
    }
  }
  __syncthreads();
}
// one thread per peer raises flag[which][rank] = epoch there; the caller has fenced (system scope) whatever the flag publishes
__device__ __forceinline__ void dp_raise(const DpPeers& P, int which, int rank, int world, uint32_t epoch) {
  if ((int)threadIdx.x < world) {
    volatile uint32_t* f = P.flags[threadIdx.x] + which + rank;
    *f = epoch;
  }
}
// the last CTA of the grid to get here raises flag[which][rank] = epoch in every rank's flag array
__device__ __forceinline__ void dp_signal_all(const DpPeers& P, uint32_t* own_flags, int which, int rank, int world, uint32_t epoch) {
  __threadfence_system();  // this CTA's stores (own and peer memory) are visible system-wide before the counter / flags move
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(own_flags + DP_COUNTER, 1u);
    if (prev == gridDim.x - 1) {
      own_flags[DP_COUNTER] = 0u;
      __threadfence_system();  // one fence, then posted stores: a release per flag would serialise world round trips
      for (int p = 0; p < world; ++p) *((volatile uint32_t*)(P.flags[p] + which + rank)) = epoch;
    }
  }
}

__global__ void __launch_bounds__(DP_THREADS) k_dp_publish(const float* __restrict__ data, int64_t n4, DpPeers P, int rank, int world, uint32_t epoch) {
  float4* dst = reinterpret_cast<float4*>(P.buf[rank]);
  const float4* src = reinterpret_cast<const float4*>(data);
  for (int64_t i = blockIdx.x * (int64_t)DP_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * DP_THREADS) dst[i] = src[i];
  dp_signal_all(P, P.flags[rank], DP_FLAG_A, rank, world, epoch);
}
__global__ void __launch_bounds__(DP_THREADS) k_dp_reduce(int64_t n4, DpPeers P, int rank, int world, uint32_t epoch) {
  dp_wait_all(P.flags[rank], DP_FLAG_A, world, epoch);
  const int64_t s0 = n4 * rank / world, s1 = n4 * (rank + 1) / world;  // this rank's slice, in float4 units
  for (int64_t i = s0 + blockIdx.x * (int64_t)DP_THREADS + threadIdx.x; i < s1; i += (int64_t)gridDim.x * DP_THREADS) {
    float4 v[DP_MAX_WORLD];
#pragma unroll
    for (int p = 0; p < DP_MAX_WORLD; ++p)
      if (p < world) v[p] = dp_load_peer(P.buf[p] + 4 * i);  // all loads in flight before the first add
    float4 acc = v[0];
#pragma unroll
    for (int p = 1; p < DP_MAX_WORLD; ++p)
      if (p < world) { acc.x += v[p].x; acc.y += v[p].y; acc.z += v[p].z; acc.w += v[p].w; }
#pragma unroll
    for (int p = 0; p < DP_MAX_WORLD; ++p)
      if (p < world) *reinterpret_cast<float4*>(P.buf[p] + 4 * i) = acc;
  }
  dp_signal_all(P, P.flags[rank], DP_FLAG_B, rank, world, epoch);
}
__global__ void __launch_bounds__(DP_THREADS) k_dp_collect(float* __restrict__ data, int64_t n4, DpPeers P, int rank, int world, uint32_t epoch) {
  dp_wait_all(P.flags[rank], DP_FLAG_B, world, epoch);
  const float* src = P.buf[rank];
  for (int64_t i = blockIdx.x * (int64_t)DP_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * DP_THREADS)
    reinterpret_cast<float4*>(data)[i] = dp_load_peer(src + 4 * i);  // peers wrote most of it: bypass L1
}

// ---- registered, in-place: one cooperative launch
__device__ __forceinline__ void dp_grid_barrier(uint32_t* counter, uint32_t target) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(counter, 1u);
    while ((int32_t)(dp_load_acquire_sys(counter) - target) < 0) __nanosleep(32);
  }
  __syncthreads();
}
template <int WORLD>
__global__ void __launch_bounds__(512) k_dp_allreduce_inplace(DpPeers P /* buf[] = the ranks' registered ranges at the call's offset */, int64_t n4,
                                                              int rank, uint32_t epoch, uint32_t barrier_target) {
  uint32_t* own = P.flags[rank];
  // A: this rank's gradients are complete (stream order) -> tell everyone; wait until everyone's are
  if (blockIdx.x == 0) {
    __threadfence_system();
    dp_raise(P, DP_FLAG_A, rank, WORLD, epoch);
  }
  dp_wait_all(own, DP_FLAG_A, WORLD, epoch);
  const int64_t s0 = n4 * rank / WORLD, s1 = n4 * (rank + 1) / WORLD;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = s0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < s1; i += 2 * stride) {
    const int64_t j = i + stride;  // two independent elements per trip: 2 x WORLD loads in flight per thread
    const bool two = j < s1;
    float4 a[WORLD], b[WORLD];
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      a[p] = dp_load_peer(P.buf[p] + 4 * i);
      b[p] = two ? dp_load_peer(P.buf[p] + 4 * j) : a[p];
    }
    float4 sa = a[0], sb = b[0];
#pragma unroll
    for (int p = 1; p < WORLD; ++p) {
      sa.x += a[p].x; sa.y += a[p].y; sa.z += a[p].z; sa.w += a[p].w;
      sb.x += b[p].x; sb.y += b[p].y; sb.z += b[p].z; sb.w += b[p].w;
    }
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      *reinterpret_cast<float4*>(P.buf[p] + 4 * i) = sa;
      if (two) *reinterpret_cast<float4*>(P.buf[p] + 4 * j) = sb;
    }
  }
  // B: every CTA's stores are out -> one signal per rank -> wait until every rank's slice has landed here
  dp_grid_barrier(own + DP_GRIDBAR, barrier_target);
  if (blockIdx.x == 0) {
    __threadfence_system();  // orders the flags after everything the grid barrier made visible to this CTA (cumulativity)
    dp_raise(P, DP_FLAG_B, rank, WORLD, epoch);
  }
  dp_wait_all(own, DP_FLAG_B, WORLD, epoch);
}
template <int WORLD>
static int dp_launch_inplace(dtc_dp* d, const DpPeers& P, int64_t n4, uint32_t epoch, cudaStream_t st) {
  if (!d->coop_blocks) {
    int dev = 0, sms = 0, per = 0;
    DTC_CUDA(cudaGetDevice(&dev));
    DTC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DTC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_dp_allreduce_inplace<WORLD>, 512, 0));
    if (per < 1) DTC_FAIL(DTC_ERR_CUDA, "k_dp_allreduce_inplace does not fit");
    d->coop_blocks = sms;  // one CTA per SM; the cooperative launch guarantees they are co-resident (the grid barrier spins)
  }
  d->bar_target += (uint32_t)d->coop_blocks;  // the barrier word only ever grows: after this launch it stands at bar_target
  DpPeers Pv = P;
  int64_t n4v = n4;
  int rank = d->rank;
  uint32_t ep = epoch, target = d->bar_target;
  void* args[] = {&Pv, &n4v, &rank, &ep, &target};
  DTC_CUDA(cudaLaunchCooperativeKernel((const void*)k_dp_allreduce_inplace<WORLD>, dim3(d->coop_blocks), dim3(512), args, 0, st));
  g_dtc_launches++;
  return DTC_OK;
}

extern "C" int dtc_dp_create(int32_t rank, int32_t world, int64_t max_floats, dtc_dp** out) {
  if (!out || world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world || max_floats <= 0)
    DTC_FAIL(DTC_ERR_ARG, "dtc_dp_create: rank %d world %d floats %lld", rank, world, (long long)max_floats);
  dtc_dp* d = new dtc_dp();
  memset(d, 0, sizeof(*d));
  d->rank = rank; d->world = world; d->n = (max_floats + 3) & ~(int64_t)3;
  if (cudaMalloc(&d->buf[rank], (size_t)d->n * sizeof(float)) != cudaSuccess || cudaMalloc(&d->flags[rank], DP_FLAG_WORDS * sizeof(uint32_t)) != cudaSuccess) {
    delete d;
    DTC_FAIL(DTC_ERR_CUDA, "dtc_dp_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  DTC_CUDA(cudaMemset(d->flags[rank], 0, DP_FLAG_WORDS * sizeof(uint32_t)));
  DTC_CUDA(cudaDeviceSynchronize());
  d->opened[rank] = true;
  *out = d;
  return DTC_OK;
}
// two cudaIpcMemHandle_t (64 bytes each): the exchange buffer and the flag array of this rank
extern "C" int dtc_dp_handles(dtc_dp* d, void* buf_handle64, void* flags_handle64) {
  if (!d || !buf_handle64 || !flags_handle64) DTC_FAIL(DTC_ERR_ARG, "dtc_dp_handles: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DTC_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(buf_handle64), d->buf[d->rank]));
  DTC_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(flags_handle64), d->flags[d->rank]));
  return DTC_OK;
}
extern "C" int dtc_dp_open(dtc_dp* d, int32_t peer, const void* buf_handle64, const void* flags_handle64) {
  if (!d || peer < 0 || peer >= d->world || peer == d->rank || !buf_handle64 || !flags_handle64) DTC_FAIL(DTC_ERR_ARG, "dtc_dp_open: bad arguments");
  if (d->opened[peer]) return DTC_OK;
  cudaIpcMemHandle_t hb, hf;
  memcpy(&hb, buf_handle64, 64); memcpy(&hf, flags_handle64, 64);
  DTC_CUDA(cudaIpcOpenMemHandle((void**)&d->buf[peer], hb, cudaIpcMemLazyEnablePeerAccess));
  DTC_CUDA(cudaIpcOpenMemHandle((void**)&d->flags[peer], hf, cudaIpcMemLazyEnablePeerAccess));
  d->opened[peer] = true;
  return DTC_OK;
}
// ---- registration of the caller's own buffer (the flat gradient buffer) for the in-place kernel
// handle of the ALLOCATION that contains `base` (cudaMalloc memory, e.g. a block of torch's caching allocator) + the byte offset of
// `base` inside it; the same [base, base + nfloats) range must be registered on every rank
extern "C" int dtc_dp_register(dtc_dp* d, float* base, int64_t nfloats, void* handle64, int64_t* offset_bytes) {
  if (!d || !base || nfloats <= 0 || !handle64 || !offset_bytes || ((uintptr_t)base & 15)) DTC_FAIL(DTC_ERR_ARG, "dtc_dp_register: bad arguments");
  typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  DTC_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) DTC_FAIL(DTC_ERR_CUDA, "cuMemGetAddressRange not available");
  CUdeviceptr abase = 0; size_t asize = 0;
  if (((PFN_range)fn)(&abase, &asize, (CUdeviceptr)(uintptr_t)base) != CUDA_SUCCESS) DTC_FAIL(DTC_ERR_CUDA, "cuMemGetAddressRange failed");
  if ((uintptr_t)base + (size_t)nfloats * 4 > (uintptr_t)abase + asize) DTC_FAIL(DTC_ERR_ARG, "dtc_dp_register: range crosses its allocation");
  DTC_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), (void*)(uintptr_t)abase));
  *offset_bytes = (int64_t)((uintptr_t)base - (uintptr_t)abase);
  d->reg[d->rank] = base; d->reg_n = nfloats; d->reg_open[d->rank] = true;
  return DTC_OK;
}
extern "C" int dtc_dp_open_registered(dtc_dp* d, int32_t peer, const void* handle64, int64_t offset_bytes) {
  if (!d || peer < 0 || peer >= d->world || peer == d->rank || !handle64 || offset_bytes < 0) DTC_FAIL(DTC_ERR_ARG, "dtc_dp_open_registered: bad arguments");
  if (d->reg_open[peer]) return DTC_OK;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DTC_CUDA(cudaIpcOpenMemHandle(&d->reg_map[peer], h, cudaIpcMemLazyEnablePeerAccess));
  d->reg[peer] = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->reg_map[peer]) + offset_bytes);
  d->reg_open[peer] = true;
  return DTC_OK;
}
static bool dp_inplace_ready(const dtc_dp* d, const float* data, int64_t n) {
  if (d->world > 8 || !d->reg[d->rank] || data < d->reg[d->rank] || data + n > d->reg[d->rank] + d->reg_n) return false;
  for (int p = 0; p < d->world; ++p)
    if (!d->reg_open[p]) return false;
  return true;
}
// in-place SUM over the ranks of data[0, n): n % 4 == 0, data 16-byte aligned; every rank must call it in the same order with the same
// offset inside its registered range (then: the single in-place kernel) or with n <= max_floats (else: the exchange-buffer path)
extern "C" int dtc_dp_allreduce(dtc_dp* d, float* data, int64_t n, void* stream) {
  DTC_NVTX("dtc_dp_allreduce");
  if (!d || !data || n <= 0 || (n & 3) || ((uintptr_t)data & 15)) DTC_FAIL(DTC_ERR_ARG, "dtc_dp_allreduce: bad arguments (n = %lld)", (long long)n);
  for (int p = 0; p < d->world; ++p)
    if (!d->opened[p]) DTC_FAIL(DTC_ERR_STATE, "dtc_dp_allreduce: peer %d not opened", p);
  DpPeers P;
  for (int p = 0; p < DP_MAX_WORLD; ++p) { P.buf[p] = d->buf[p]; P.flags[p] = d->flags[p]; }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  if (dp_inplace_ready(d, data, n)) {
    const int64_t off = data - d->reg[d->rank];
    for (int p = 0; p < d->world; ++p) P.buf[p] = d->reg[p] + off;
    const uint32_t ep = ++d->epoch;
    switch (d->world) {
      case 2: return dp_launch_inplace<2>(d, P, n4, ep, st);
      case 3: return dp_launch_inplace<3>(d, P, n4, ep, st);
      case 4: return dp_launch_inplace<4>(d, P, n4, ep, st);
      case 5: return dp_launch_inplace<5>(d, P, n4, ep, st);
      case 6: return dp_launch_inplace<6>(d, P, n4, ep, st);
      case 7: return dp_launch_inplace<7>(d, P, n4, ep, st);
      default: return dp_launch_inplace<8>(d, P, n4, ep, st);
    }
  }
  if (n > d->n) DTC_FAIL(DTC_ERR_ARG, "dtc_dp_allreduce: %lld floats exceed the exchange buffer (%lld)", (long long)n, (long long)d->n);
  const uint32_t epoch = ++d->epoch;
  const int blocks = (int)(n4 < (int64_t)DP_BLOCKS * DP_THREADS ? (n4 + DP_THREADS - 1) / DP_THREADS : DP_BLOCKS);
  k_dp_publish<<<blocks, DP_THREADS, 0, st>>>(data, n4, P, d->rank, d->world, epoch);
  DTC_CHECK_LAUNCH("k_dp_publish");
  k_dp_reduce<<<blocks, DP_THREADS, 0, st>>>(n4, P, d->rank, d->world, epoch);
  DTC_CHECK_LAUNCH("k_dp_reduce");
  k_dp_collect<<<blocks, DP_THREADS, 0, st>>>(data, n4, P, d->rank, d->world, epoch);
  DTC_CHECK_LAUNCH("k_dp_collect");
  return DTC_OK;
}
// 0 = no rank ever timed out waiting for a peer (reads the error word: synchronises the stream it is given)
extern "C" int dtc_dp_error(dtc_dp* d, void* stream) {
  if (!d) return DTC_ERR_ARG;
  uint32_t e = 0;
  if (cudaMemcpyAsync(&e, d->flags[d->rank] + DP_ERROR, sizeof(e), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return DTC_ERR_CUDA;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return DTC_ERR_CUDA;
  if (e) DTC_FAIL(DTC_ERR_STATE, "dtc_dp: a rank waited more than ~20 s for a peer's flag (mismatched all-reduce sequence?)");
  return DTC_OK;
}
extern "C" void dtc_dp_destroy(dtc_dp* d) {
  if (!d) return;
  cudaDeviceSynchronize();
  for (int p = 0; p < d->world; ++p) {
    if (p == d->rank || !d->opened[p]) continue;
    cudaIpcCloseMemHandle(d->buf[p]);
    cudaIpcCloseMemHandle(d->flags[p]);
    if (d->reg_map[p]) cudaIpcCloseMemHandle(d->reg_map[p]);
  }
  cudaFree(d->buf[d->rank]);
  cudaFree(d->flags[d->rank]);
  delete d;
}
