// Standalone probe for the Blackwell pieces the tensor-core GEMM is built from (run on the B200 box):
//   tc_probe tma                      TMA 2-D box load, fp32, SWIZZLE_128B -> verifies the shared-memory swizzle pattern
//   tc_probe mma <amaj> <bmaj> <K>    one 128x128 tile: D = A B^T with tcgen05.mma kind::tf32, K-major (0) / MN-major (1)
//                                     operands staged by TMA; reports the error against truncated- and rounded-TF32 references
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_probe tools/tc_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
  return (PFN_encodeTiled)fn;
}
// 2-D fp32 tensor [rows, cols] row-major with leading dimension ld; box {box_cols (inner), box_rows}; 128-byte swizzle
static CUtensorMap make_map(const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows,
                            CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
  return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\tbra.uni WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ test 1: TMA swizzle pattern
__global__ void k_tma(const __grid_constant__ CUtensorMap map, float* out, int c0, int c1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ __align__(8) uint64_t bar;
  float* tile = (float*)smem;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    mbar_expect_tx(&bar, 128 * 32 * 4);
    tma_load_2d(tile, &map, &bar, c0, c1);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) out[i] = tile[i];
}

static int test_tma() {
  const int R = 256, C = 96;
  std::vector<float> h(R * C);
  for (int i = 0; i < R * C; ++i) h[i] = (float)i;
  float *d, *o;
  CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&o, 128 * 32 * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap m = make_map(d, R, C, C, 32, 128);
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024));
  k_tma<<<1, 128, 32768 + 1024>>>(m, o, 32, 64);
  CK(cudaDeviceSynchronize());
  std::vector<float> r(128 * 32);
  CK(cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int row = 0; row < 128; ++row)
    for (int c = 0; c < 32; ++c) {
      int off = row * 32 + (((c / 4) ^ (row % 8)) * 4) + (c % 4);
      float want = (float)((64 + row) * C + 32 + c);
      if (r[off] != want) { if (bad < 5) printf("mismatch row %d col %d: got %f want %f\n", row, c, r[off], want); ++bad; }
    }
  printf("TMA swizzle-128B test: %s (%d mismatches)\n", bad ? "FAIL" : "PASS", bad);
  return bad != 0;
}

// ------------------------------------------------------------------ test 2: tcgen05.mma kind::tf32 single tile
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version 1 (Blackwell)
  d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (MN-major tf32)
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}

// A tile: 128 (M) x 32 (K) floats per k-block; B tile the same.  K-major: one TMA box {32, 128}; MN-major: 4 boxes {32 mn, 32 k}.
template <int AMAJ, int BMAJ>
__global__ void __launch_bounds__(256) k_mma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* D,
                                             int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  float* sA = (float*)smem;            // 16 KB
  float* sB = (float*)(smem + 16384);  // 16 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), majors, N=128 (16<<17), M=128 (8<<24)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)AMAJ << 15) | ((uint32_t)BMAJ << 16) | ((128u >> 3) << 17) |
                         ((128u >> 4) << 24);
  const int nkb = K / 32;
  for (int kb = 0; kb < nkb; ++kb) {
    if (warp == 0 && elect_one()) {
      mbar_expect_tx(&bar_full, 32768);
      if (AMAJ == 0) tma_load_2d(sA, &mapA, &bar_full, kb * 32, 0);
      else for (int j = 0; j < 4; ++j) tma_load_2d(sA + j * 1024, &mapA, &bar_full, j * 32, kb * 32);
      if (BMAJ == 0) tma_load_2d(sB, &mapB, &bar_full, kb * 32, 0);
      else for (int j = 0; j < 4; ++j) tma_load_2d(sB + j * 1024, &mapB, &bar_full, j * 32, kb * 32);
    }
    if (warp == 1) {
      mbar_wait(&bar_full, kb & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        for (int k = 0; k < 4; ++k) {  // 4 MMAs of K = 8 per 32-wide k-block
          uint64_t ad = AMAJ == 0 ? make_desc(smem_u32(sA) + k * 32, 16, 1024) : make_desc(smem_u32(sA) + k * 1024, 4096, 512, 1);
          uint64_t bd = BMAJ == 0 ? make_desc(smem_u32(sB) + k * 32, 16, 1024) : make_desc(smem_u32(sB) + k * 1024, 4096, 512, 1);
          mma_tf32(tmem, ad, bd, idesc, (kb | k) ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
      }
      __syncwarp();
    }
    // everybody waits until this k-block's MMAs have consumed the tiles (single-stage probe)
    mbar_wait(&bar_mma, kb & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp >= 4) {
    const int q = warp & 3;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
          "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = q * 32 + lane;
      for (int j = 0; j < 32; ++j) D[row * 128 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

static float tf32_trunc(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xFFFFE000u; memcpy(&x, &b, 4); return x; }
static float tf32_rn(float x) { uint32_t b; memcpy(&b, &x, 4); b += 0x00000FFFu + ((b >> 13) & 1u); b &= 0xFFFFE000u; memcpy(&x, &b, 4); return x; }

static int test_mma(int amaj, int bmaj, int K) {
  const int M = 128, N = 128;
  std::vector<float> A(M * K), B(N * K);  // logical A[m][k], B[n][k]
  srand(7);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  std::vector<float> Ag(M * K), Bg(N * K);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Ag[amaj ? k * M + m : m * K + k] = A[m * K + k];
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bg[bmaj ? k * N + n : n * K + k] = B[n * K + k];
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, Ag.size() * 4)); CK(cudaMalloc(&dB, Bg.size() * 4)); CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, Ag.data(), Ag.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bg.data(), Bg.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, M * N * 4));
  CUtensorMap mA = amaj ? make_map(dA, K, M, M, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) : make_map(dA, M, K, K, 32, 128);
  CUtensorMap mB = bmaj ? make_map(dB, K, N, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) : make_map(dB, N, K, K, 32, 128);
  auto launch = [&](auto kern) {
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 34816));
    kern<<<1, 256, 34816>>>(mA, mB, dD, K);
  };
  if (!amaj && !bmaj) launch(k_mma<0, 0>);
  else if (!amaj && bmaj) launch(k_mma<0, 1>);
  else if (amaj && !bmaj) launch(k_mma<1, 0>);
  else launch(k_mma<1, 1>);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("MMA a_major=%d b_major=%d K=%d: CUDA error %s\n", amaj, bmaj, K, cudaGetErrorString(e)); return 1; }
  std::vector<float> D(M * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double et = 0, er = 0, ef = 0, mag = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double st = 0, sr = 0, sf = 0;
      for (int k = 0; k < K; ++k) {
        st += (double)tf32_trunc(A[m * K + k]) * (double)tf32_trunc(B[n * K + k]);
        sr += (double)tf32_rn(A[m * K + k]) * (double)tf32_rn(B[n * K + k]);
        sf += (double)A[m * K + k] * (double)B[n * K + k];
      }
      double d = D[m * N + n];
      et = fmax(et, fabs(d - st)); er = fmax(er, fabs(d - sr)); ef = fmax(ef, fabs(d - sf)); mag = fmax(mag, fabs(sf));
    }
  printf("MMA a_major=%d b_major=%d K=%d: max|D-ref| trunc-inputs %.3e  rn-inputs %.3e  exact-inputs %.3e  (max|ref| %.3f) -> %s\n", amaj,
         bmaj, K, et, er, ef, mag, (et < 1e-3 * mag || er < 1e-3 * mag) ? "PASS" : "FAIL");
  return !(et < 1e-3 * mag || er < 1e-3 * mag);
}

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: tc_probe tma | mma amaj bmaj K\n"); return 1; }
  if (!strcmp(argv[1], "tma")) return test_tma();
  if (!strcmp(argv[1], "mma")) return test_mma(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
  return 1;
}
