// P14 / SURVEY 8f N2: Memory (rsl_rl/rsl_rl/modules/actor_critic_decoder.py:584-614) = nn.GRU(input_size, hidden_size, num_layers)
// forward for N rows over T time steps plus the per-row hidden reset on `dones`.  The reference declares this module but never
// instantiates it on the training path; it is built here because the task names it, with torch.nn.GRU as oracle.
//   r = sigmoid(W_ir x + b_ir + W_hr h + b_hr),  z = sigmoid(W_iz x + b_iz + W_hz h + b_hz),
//   n = tanh(W_in x + b_in + r * (W_hn h + b_hn)),  h' = (1 - z) * n + z * h            (PyTorch gate order r | z | n)
// One launch per (time step, layer): a CTA owns GRU_ROWS rows of the batch, stages [x | h] for them and streams the 3H gate rows
// of W_ih / W_hh through registers (the weights, <= 3H x (in + H) floats, stay in L1/L2: 150 x 103 floats for the reference's
// AC_Args rnn_hidden_size = 50 on a 53-wide observation).  Thread (g, r): gate row g of batch row r - 3H x GRU_ROWS dot products
// of length in + H per CTA, then the H x GRU_ROWS gate combinations from shared memory.
#include "dtc_common.cuh"

#define GRU_ROWS 8
#define GRU_MAX_IN 1024
#define GRU_MAX_H 128

__global__ void __launch_bounds__(256) k_gru_step(int N, int in, int H, const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                                  const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                  const float* __restrict__ x, int ldx, float* __restrict__ h, float* __restrict__ out) {
  extern __shared__ float gru_sh[];
  float* xs = gru_sh;                       // [GRU_ROWS][in]
  float* hs = xs + GRU_ROWS * in;           // [GRU_ROWS][H]
  float* gi = hs + GRU_ROWS * H;            // [GRU_ROWS][3H]  W_i* x + b_i*
  float* gh = gi + GRU_ROWS * 3 * H;        // [GRU_ROWS][3H]  W_h* h + b_h*
  const int row0 = blockIdx.x * GRU_ROWS;
  const int rows = min(GRU_ROWS, N - row0);
  for (int e = threadIdx.x; e < rows * in; e += blockDim.x) { const int r = e / in, c = e - r * in; xs[r * in + c] = x[(size_t)(row0 + r) * ldx + c]; }
  for (int e = threadIdx.x; e < rows * H; e += blockDim.x) { const int r = e / H, c = e - r * H; hs[r * H + c] = h[(size_t)(row0 + r) * H + c]; }
  __syncthreads();
  // gate pre-activations: GRU_ROWS rows share every weight row read
  for (int g = threadIdx.x; g < 3 * H; g += blockDim.x) {
    float ai[GRU_ROWS], ah[GRU_ROWS];
#pragma unroll
    for (int r = 0; r < GRU_ROWS; ++r) { ai[r] = 0.f; ah[r] = 0.f; }
    const float* wi = w_ih + (size_t)g * in;
    for (int c = 0; c < in; ++c) {
      const float w = __ldg(wi + c);
#pragma unroll
      for (int r = 0; r < GRU_ROWS; ++r) ai[r] = fmaf(w, xs[r * in + c], ai[r]);
    }
    const float* wh = w_hh + (size_t)g * H;
    for (int c = 0; c < H; ++c) {
      const float w = __ldg(wh + c);
#pragma unroll
      for (int r = 0; r < GRU_ROWS; ++r) ah[r] = fmaf(w, hs[r * H + c], ah[r]);
    }
    const float bi = __ldg(b_ih + g), bh = __ldg(b_hh + g);
#pragma unroll
    for (int r = 0; r < GRU_ROWS; ++r) { gi[r * 3 * H + g] = ai[r] + bi; gh[r * 3 * H + g] = ah[r] + bh; }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < rows * H; e += blockDim.x) {
    const int r = e / H, j = e - r * H;
    const float* a = gi + r * 3 * H;
    const float* b = gh + r * 3 * H;
    const float rg = 1.0f / (1.0f + expf(-(a[j] + b[j])));
    const float zg = 1.0f / (1.0f + expf(-(a[H + j] + b[H + j])));
    const float ng = tanhf(a[2 * H + j] + rg * b[2 * H + j]);
    const float hn = (1.0f - zg) * ng + zg * hs[r * H + j];
    h[(size_t)(row0 + r) * H + j] = hn;
    if (out) out[(size_t)(row0 + r) * H + j] = hn;
  }
}

__global__ void __launch_bounds__(256) k_gru_reset(int N, int H, int L, float* __restrict__ h, const uint8_t* __restrict__ dones) {
  const int64_t total = (int64_t)L * N * H;
  for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    const int n = (int)((e / H) % N);
    if (dones[n]) h[e] = 0.f;
  }
}

// floats of one layer's parameters in nn.GRU order: weight_ih_l | weight_hh_l | bias_ih_l | bias_hh_l
static int64_t gru_layer_floats(int in_l, int H) { return (int64_t)3 * H * in_l + (int64_t)3 * H * H + 6 * H; }

extern "C" int64_t dtc_gru_param_floats(int32_t input_size, int32_t hidden_size, int32_t num_layers) {
  int64_t tot = 0;
  for (int l = 0; l < num_layers; ++l) tot += gru_layer_floats(l == 0 ? input_size : hidden_size, hidden_size);
  return tot;
}

extern "C" int dtc_gru_forward(int32_t T, int32_t N, int32_t input_size, int32_t hidden_size, int32_t num_layers, const float* weights,
                               const float* x, float* h, float* out, void* stream) {
  DTC_NVTX("dtc_gru_forward");
  if (!weights || !x || !h) DTC_FAIL(DTC_ERR_ARG, "dtc_gru_forward: null argument");
  if (T <= 0 || N <= 0 || num_layers <= 0) DTC_FAIL(DTC_ERR_ARG, "dtc_gru_forward: T, N, num_layers must be positive");
  if (input_size <= 0 || input_size > GRU_MAX_IN || hidden_size <= 0 || hidden_size > GRU_MAX_H)
    DTC_FAIL(DTC_ERR_ARG, "dtc_gru_forward: input_size <= %d and hidden_size <= %d are supported", GRU_MAX_IN, GRU_MAX_H);
  if (num_layers > 1 && !out) DTC_FAIL(DTC_ERR_ARG, "dtc_gru_forward: stacked layers need the [T,N,H] output buffer as inter-layer storage");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = hidden_size;
  static int attr_bytes = 0;
  const int smem_max = (int)sizeof(float) * GRU_ROWS * (GRU_MAX_IN + GRU_MAX_H + 6 * GRU_MAX_H);
  if (!attr_bytes) { DTC_CUDA(cudaFuncSetAttribute(k_gru_step, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)); attr_bytes = smem_max; }
  for (int t = 0; t < T; ++t) {
    const float* w = weights;
    for (int l = 0; l < num_layers; ++l) {
      const int in_l = l == 0 ? input_size : H;
      const float* w_ih = w;
      const float* w_hh = w_ih + (size_t)3 * H * in_l;
      const float* b_ih = w_hh + (size_t)3 * H * H;
      const float* b_hh = b_ih + 3 * H;
      // layer l > 0 reads the layer below's output of this time step from `out` and overwrites it in place (each CTA stages
      // its rows' inputs in shared memory before any of its threads writes)
      const float* xin = l == 0 ? x + (size_t)t * N * input_size : out + (size_t)t * N * H;
      float* o = out ? out + (size_t)t * N * H : nullptr;
      const int smem = (int)sizeof(float) * GRU_ROWS * (in_l + H + 6 * H);
      k_gru_step<<<ceil_div(N, GRU_ROWS), 256, smem, st>>>(N, in_l, H, w_ih, w_hh, b_ih, b_hh, xin, in_l, h + (size_t)l * N * H, o);
      DTC_CHECK_LAUNCH("k_gru_step");
      w += gru_layer_floats(in_l, H);
    }
  }
  return DTC_OK;
}

extern "C" int dtc_gru_reset(int32_t N, int32_t hidden_size, int32_t num_layers, float* h, const uint8_t* dones, void* stream) {
  if (!h || !dones || N <= 0 || hidden_size <= 0 || num_layers <= 0) DTC_FAIL(DTC_ERR_ARG, "dtc_gru_reset: bad arguments");
  const int64_t total = (int64_t)num_layers * N * hidden_size;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_gru_reset<<<blocks, 256, 0, (cudaStream_t)stream>>>(N, hidden_size, num_layers, h, dones);
  DTC_CHECK_LAUNCH("k_gru_reset");
  return DTC_OK;
}
