// FP32 GEMM family for the MLP stack (forward / dgrad / wgrad) - interface.
#pragma once
#include "dtc_common.cuh"

enum GemmEpi {
  EPI_STORE = 0,      // C = acc
  EPI_BIAS = 1,       // C = acc + bias[n]
  EPI_BIAS_RELU = 2,  // C = relu(acc + bias[n])
  EPI_BIAS_ELU = 3,   // C = elu(acc + bias[n])
  EPI_DRELU = 4,      // C = acc * (act_src > 0)
  EPI_DELU = 5,       // C = acc * (act_src > 0 ? 1 : act_src + 1)      (d/dx elu(x) written with y = elu(x))
};

struct GemmArgs {
  // C[m,n] = epi( sum_k A(m,k) * B(n,k) )
  const float* A; int lda; bool a_kc;  // a_kc: A(m,k) = A[m*lda + k]   else A[k*lda + m]
  const float* B; int ldb; bool b_kc;  // b_kc: B(n,k) = B[n*ldb + k]   else B[k*ldb + n]
  float* C; int ldc;
  int M, N, K;        // logical extents.  Storage contract: every base pointer is 16-byte aligned and every leading
                      // dimension is a multiple of 4 floats; tails inside a float4 are masked by the loaders.  Columns
                      // [N, round4(N)) of C are either left alone or (tensor-core path, ldc == round4(N)) set to zero.
  const float* bias;  // [N] for EPI_BIAS*
  const float* act_src; int ld_act;  // [M, ld_act] for EPI_D*
  int epi;
  bool accumulate;    // C += result (dgrad fan-in)
  int splits;         // >1: split the reduction; partial tiles go to `ws` [splits][M][ldc] and are summed into C
  float* ws;
  int k_per_split;    // filled by the launcher
  // 3xTF32 companions (dtc_gemm_tc.cu): x_lo = rn_tf32(x - trunc_tf32(x)).  A_lo / B_lo feed the tensor-core path (NULL = that
  // correction term is skipped); C_lo, when set, receives the companion of the result from either path's epilogue.
  const float* A_lo; const float* B_lo; float* C_lo;
  // a_split / b_split: that operand has no companion array; the tensor-core kernels compute its companion tile in shared memory
  bool a_split, b_split;
  // tensor-core path only, splits == 1: when set, the epilogue also writes the column sums of every 32-row block of the final C
  // to colsum_part[ceil(M/32)][round4(N)] (the bias gradient of the layer below a dgrad, without re-reading C from HBM)
  float* colsum_part;
  // alternative (tensor-core path, splits == 1): the epilogue ADDS the column sums of C[:, :colsum_n] into colsum_out[0, colsum_n)
  // with L2 reductions (red.global.add.f32) - the caller zeroes colsum_out beforehand; no partial buffer, no reduce kernel
  float* colsum_out; int colsum_n;
  bool c_zeroed;  // split-K through L2 reductions: C is already zero (skip the launcher's memset)
};
// out[n] = sum_b part[b*ld + n], b < nblk (deterministic order)
int dtc_colsum_part_launch(const float* part, int nblk, int ld, int ncols, float* out, cudaStream_t st);

// GEMM engine: 0 = FP32 SIMT everywhere, 1 = tcgen05 3xTF32 where the shape is tile-worthy (default; env DTC_GEMM=simt|tc)
int dtc_gemm_mode();
void dtc_gemm_set_mode(int mode);
bool dtc_gemm_tc_eligible(const GemmArgs& a);
int dtc_gemm_tc_launch(GemmArgs a, cudaStream_t st);
// launches on `st`; returns 0 or a negative dtc_status with the message in dtc_last_error()
int dtc_gemm_launch(GemmArgs a, cudaStream_t st);
// splits the launcher would pick for a reduction of length K producing an MxN output (workspace sizing)
int dtc_gemm_pick_splits(int M, int N, int K);
// column sums: out[n] = sum_m X[m*ld + n]  (bias gradients); ws >= COLSUM_CHUNKS*N floats
#define COLSUM_CHUNKS 64
int dtc_colsum_launch(const float* X, int ld, int M, int N, float* out, float* ws, cudaStream_t st);
