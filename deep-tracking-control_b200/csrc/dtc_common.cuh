// Shared device/host helpers for libdtc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/dtc_b200.h"

#define NP 693  // 33 x 21 height points
#define GXN 33
#define GYN 21

extern thread_local char g_dtc_err[512];
extern int64_t g_dtc_launches;

#define DTC_FAIL(code, ...)                                  \
  do {                                                       \
    snprintf(g_dtc_err, sizeof(g_dtc_err), __VA_ARGS__);     \
    return (code);                                           \
  } while (0)

#define DTC_CHECK_LAUNCH(name)                                                            \
  do {                                                                                    \
    g_dtc_launches++;                                                                     \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) DTC_FAIL(DTC_ERR_CUDA, "%s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

#define DTC_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) DTC_FAIL(DTC_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e));     \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// optional per-launch CUDA-event timing (bench.py's roofline leg): kind 0 = GEMM family (work = flops), 1 = foothold kernel
extern int g_dtc_prof;
void dtc_prof_begin(cudaStream_t st, int kind, double work);
void dtc_prof_tag(int m, int n, int k, int layout);  // shape of the launch opened by the last dtc_prof_begin (per-shape dump)
void dtc_prof_end(cudaStream_t st);

// NVTX range over a C-ABI entry point (the reference has no tracing; a timeline tool sees the env / learner
// phases by name).  Header-only NVTX3: a pointer check when no tool is attached.
#include <nvtx3/nvToolsExt.h>
struct DtcNvtxRange {
  explicit DtcNvtxRange(const char* name) { nvtxRangePushA(name); }
  ~DtcNvtxRange() { nvtxRangePop(); }
};
#define DTC_NVTX(name) DtcNvtxRange dtc_nvtx_range_(name)

#define RETURN_IF_ERR(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

// 3xTF32 companion value: the tensor core truncates an fp32 operand x to TF32; x_lo = rn_tf32(x - trunc_tf32(x)) carries
// the next 11 bits (dtc_gemm_tc.cu)
__host__ __device__ __forceinline__ float tf32_lo(float x) {
#ifdef __CUDA_ARCH__
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  const float r = x - hi;
  return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
#else
  uint32_t b; memcpy(&b, &x, 4); b &= 0xFFFFE000u; float hi; memcpy(&hi, &b, 4);
  float r = x - hi; memcpy(&b, &r, 4); b = (b + 0x1000u) & 0xFFFFE000u; memcpy(&r, &b, 4); return r;
#endif
}

// ------------------------------------------------------------------ warp reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// lexicographic (value, index) minimum: lowest index wins ties (CPU topk/min semantics)
__device__ __forceinline__ void warp_argmin(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// block-wide sum of doubles; result valid in thread 0. `sh` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? sh[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// ------------------------------------------------------------------ Philox4x32-10 counter RNG
struct Philox {
  uint32_t c[4], k[2];
  __device__ __forceinline__ Philox(uint64_t seed, uint64_t ctr_hi, uint64_t ctr_lo) {
    k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
    c[0] = (uint32_t)ctr_lo; c[1] = (uint32_t)(ctr_lo >> 32);
    c[2] = (uint32_t)ctr_hi; c[3] = (uint32_t)(ctr_hi >> 32);
  }
  __device__ __forceinline__ uint4 next() {
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    c[0]++;  // advance the low counter word for the next draw
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
  float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.28318530717958647692f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// ------------------------------------------------------------------ exact-op quaternion helpers.
// Operation order follows oracle/env_oracle.py (measured ATen CPU behaviour): cross components are one FMA,
// fma(a1,b2,-(a2*b1)); the 3-dot is ((a0b0+a1b1)+a2b2) with rounded products; everything else rounds per op.
__device__ __forceinline__ void cross_exact(const float a[3], const float b[3], float c[3]) {
  c[0] = __fmaf_rn(a[1], b[2], -__fmul_rn(a[2], b[1]));
  c[1] = __fmaf_rn(a[2], b[0], -__fmul_rn(a[0], b[2]));
  c[2] = __fmaf_rn(a[0], b[1], -__fmul_rn(a[1], b[0]));
}
__device__ __forceinline__ void quat_rotate_inverse_exact(const float q[4], const float v[3], float out[3]) {
  float qw = q[3];
  float s = __fsub_rn(__fmul_rn(2.0f, __fmul_rn(qw, qw)), 1.0f);
  float cr[3];
  cross_exact(q, v, cr);
  float dot = __fadd_rn(__fadd_rn(__fmul_rn(q[0], v[0]), __fmul_rn(q[1], v[1])), __fmul_rn(q[2], v[2]));
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = __fmul_rn(v[i], s);
    float b = __fmul_rn(__fmul_rn(cr[i], qw), 2.0f);
    float c = __fmul_rn(__fmul_rn(q[i], dot), 2.0f);
    out[i] = __fadd_rn(__fsub_rn(a, b), c);
  }
}
__device__ __forceinline__ void quat_apply_exact(const float q[4], const float b[3], float out[3]) {
  float c[3], t[3], d[3];
  cross_exact(q, b, c);
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = __fmul_rn(c[i], 2.0f);
  cross_exact(q, t, d);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = __fadd_rn(__fadd_rn(b[i], __fmul_rn(q[3], t[i])), d[i]);
}
// yaw-only normalised quaternion (0,0,z,w) of legged_gym/utils/math.py:8-12
__device__ __forceinline__ void yaw_quat_exact(float qz, float qw, float& yz, float& yw) {
  float n = __fsqrt_rn(__fadd_rn(__fmul_rn(qz, qz), __fmul_rn(qw, qw)));
  n = fmaxf(n, 1e-9f);
  yz = __fdiv_rn(qz, n);
  yw = __fdiv_rn(qw, n);
}
// R_yaw * (gx, gy, 0): the two non-trivial components of quat_apply((0,0,yz,yw), (gx,gy,0))
__device__ __forceinline__ void yaw_apply_exact(float yz, float yw, float gx, float gy, float& rx, float& ry) {
  float t0 = __fmul_rn(-__fmul_rn(yz, gy), 2.0f);  // 2 * fma(0,0,-(yz*gy))
  float t1 = __fmul_rn(__fmul_rn(yz, gx), 2.0f);   // 2 * fma(yz,gx,-0)
  float d0 = -__fmul_rn(yz, t1);                   // fma(0,t2,-(yz*t1))
  float d1 = __fmul_rn(yz, t0);                    // fma(yz,t0,-(0*t2))
  rx = __fadd_rn(__fadd_rn(gx, __fmul_rn(yw, t0)), d0);
  ry = __fadd_rn(__fadd_rn(gy, __fmul_rn(yw, t1)), d1);
}
