// Standalone throughput probe for tcgen05.mma kind::tf32 issue patterns (run on the B200 box): how many cycles does the tensor
// pipe need per 128xNx8 MMA when (a) every MMA accumulates into one TMEM tile, (b) MMAs alternate between two tiles the way the
// 3xTF32 GEMM does (main, correction, correction), (c) N = 256, (d) one big commit vs a commit every 12 MMAs.
// Operands are whatever is in shared memory (zeros): only timing matters.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o mma_rate_probe tools/mma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\tbra.uni WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// mode 0: one accumulator, N=128.  1: main/corr/corr pattern, N=128.  2: one accumulator, N=256.  3: main/corr/corr, N=256
// mode 4: like 1 but three distinct accumulators (main, corr1, corr2).  5: like 0 but a commit+wait every 12 MMAs
__global__ void __launch_bounds__(128, 1) k_probe(int mode, int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const uint32_t s0 = (smem_u32(raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(raw + (s0 - smem_u32(raw)))[i] = 0.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (warp == 1) {
    if (elect_one()) {
      const int N = (mode == 2 || mode == 3) ? 256 : 128;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t hi = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));
      auto lo = [&](uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | ((16u >> 4) << 16); };
      // four "stages" of operands so consecutive k-blocks read different shared memory
      const long long t0 = clock64();
      uint32_t par = 0;
      for (int r = 0; r < reps; ++r) {
        const uint32_t st = s0 + (r & 3) * (32 * 1024);
        const uint32_t a0 = lo(st), al0 = lo(st + 4 * 1024), b0 = lo(st + 8 * 1024), bl0 = lo(st + 12 * 1024);  // overlapping tiles: timing only
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {
          const uint32_t acc = (r == 0 && k8 == 0) ? 0u : 1u;
          if (mode == 0 || mode == 2 || mode == 5) {
            mma(tmem, a0 + 2 * k8, b0 + 2 * k8, hi, idesc, acc);
            mma(tmem, al0 + 2 * k8, b0 + 2 * k8, hi, idesc, 1u);
            mma(tmem, a0 + 2 * k8, bl0 + 2 * k8, hi, idesc, 1u);
          } else if (mode == 1 || mode == 3) {
            mma(tmem, a0 + 2 * k8, b0 + 2 * k8, hi, idesc, acc);
            mma(tmem + N, al0 + 2 * k8, b0 + 2 * k8, hi, idesc, acc);
            mma(tmem + N, a0 + 2 * k8, bl0 + 2 * k8, hi, idesc, 1u);
          } else {
            mma(tmem, a0 + 2 * k8, b0 + 2 * k8, hi, idesc, acc);
            mma(tmem + 128, al0 + 2 * k8, b0 + 2 * k8, hi, idesc, acc);
            mma(tmem + 256, a0 + 2 * k8, bl0 + 2 * k8, hi, idesc, acc);
          }
        }
        if (mode == 5) { commit(&bar); mbar_wait(&bar, par); par ^= 1; }
      }
      if (mode != 5) { commit(&bar); mbar_wait(&bar, 0); }
      const long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 2000;
  const int ctas = argc > 2 ? atoi(argv[2]) : 148;
  long long* d;
  CK(cudaMalloc(&d, ctas * sizeof(long long)));
  const int smem = 161 * 1024 + 1024;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 6; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      k_probe<<<ctas, 128, smem>>>(mode, reps, d);
      CK(cudaDeviceSynchronize());
    }
    long long h[256];
    CK(cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0, mn = 1ll << 62;
    for (int i = 0; i < ctas; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
    const double mmas = 12.0 * reps;
    const int N = (mode == 2 || mode == 3) ? 256 : 128;
    printf("mode %d N=%d ctas=%d: %.1f .. %.1f cycles per MMA  (ideal %d)\n", mode, N, ctas, mn / mmas, mx / mmas, N / 2);
  }
  return 0;
}
