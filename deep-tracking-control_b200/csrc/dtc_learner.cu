// Learner half of the hot path: ActorCriticDecoder forward, the VAE and PPO optimizer steps with
// hand-written backward, GAE, minibatch gather, transition store, grad-clip + Adam.
//   rsl_rl/rsl_rl/modules/actor_critic_decoder.py:91-302 (Vae), :305-551 (ActorCriticDecoder)
//   rsl_rl/rsl_rl/algorithms/ppo.py:137-357 (PPO.act / process_env_step / compute_returns / update)
//   rsl_rl/rsl_rl/storage/rollout_storage.py:99-214 (add_transitions / compute_returns / mini_batch_generator)
// All Linear layers run on the GEMM family of dtc_gemm.cu; everything else is fused into a few row kernels below.
// No host synchronisation anywhere: the learning rate, KL mean, gradient norm and loss sums live in device memory.
#include <math.h>
#include <string>
#include <vector>

#include "dtc_gemm.cuh"

// ------------------------------------------------------------------ layer / parameter table
enum LayerId {
  CD0, CD2, CD4, TD0, TD2, TD4,            // VAE-step only: cenet_decoder, terrain_decoder
  CE0, CE2, LAT, TE0, TE2, TE4,            // shared: cenet_encoder, latent_mu|latent_var (fused rows), terrain_encoder
  AB0, AB2, AB4, AB6, CB0, CB2, CB4, CB6,  // policy-step only: actor_body, critic_body (+ std)
  MM0, MM2, MM4, GB0, GB2,                 // never trained: memory_mlp (act_teacher), gb_encoder
  NLAYERS
};
struct Layer { const char* name; int out, in, in_ref, ld; int64_t w, b; };

// internal column layouts of the concatenated first-layer inputs
#define LD_HIST 268
#define LD_PRIVA 696
#define LD_XC 752
#define XC_OBS 696   // xc = [priv[693:1389] | obs 53 | base_vel 3]
#define XC_BV 749
#define LD_XA 588    // xa = [l_t 512 | obs 53 | 3 zero | z 16 | mu[:3] | 1 zero]
#define XA_OBS 512
#define XA_Z 568
#define XA_MU 584
#define LD_XD 532    // xd = [l_t 512 | z 16 | mu[:3] | 1 zero]
#define XD_Z 512
#define XD_MU 528
#define LD_XM 780    // xm = [l_t 512 | obs_history 265 | 3 zero]   (memory_mlp input of act_teacher)
#define LD_ML 36     // ml = [mu 19 | logvar 16 | 1 zero]
#define ML_LV 19
#define NPIGGY 4
#define DTC_NEV 32

static Layer g_layers[NLAYERS];
static std::vector<dtc_param_info> g_params;
static int64_t g_total = 0, g_vae_end = 0, g_pol_begin = 0, g_pol_end = 0, g_off_std = 0, g_off_piggy = 0;
static bool g_table_built = false;

static int round4(int x) { return (x + 3) & ~3; }

static void push_param(const std::string& name, int64_t off, int rows, int cols, int ld, int nseg = 1, const int* src = nullptr,
                       const int* dst = nullptr, const int* len = nullptr) {
  dtc_param_info p;
  memset(&p, 0, sizeof(p));
  snprintf(p.name, sizeof(p.name), "%s", name.c_str());
  p.offset = off; p.rows = rows; p.cols = cols; p.ld = ld; p.nseg = nseg;
  if (src) {
    for (int i = 0; i < nseg; ++i) { p.seg_src[i] = src[i]; p.seg_dst[i] = dst[i]; p.seg_len[i] = len[i]; }
  } else {
    p.seg_src[0] = 0; p.seg_dst[0] = 0; p.seg_len[0] = cols;
  }
  g_params.push_back(p);
}

static void build_table() {
  if (g_table_built) return;
  g_table_built = true;
  int64_t cur = 0;
  // `in` is the GEMM reduction length over the INTERNAL column layout (pads inside it are zero columns)
  auto add = [&](LayerId id, const char* name, int out, int in_ref, int in_internal = -1) {
    Layer& l = g_layers[id];
    const int in = in_internal > 0 ? in_internal : in_ref;
    l.name = name; l.out = out; l.in = in; l.in_ref = in_ref; l.ld = round4(in);
    l.w = cur; cur += (int64_t)out * l.ld;
    l.b = cur; cur += round4(out);
  };
  auto expose = [&](LayerId id, int nseg = 1, const int* src = nullptr, const int* dst = nullptr, const int* len = nullptr) {
    const Layer& l = g_layers[id];
    push_param(std::string(l.name) + ".weight", l.w, l.out, l.in_ref, l.ld, nseg, src, dst, len);
    push_param(std::string(l.name) + ".bias", l.b, 1, l.out, l.out);
  };
  // --- VAE-step only
  add(CD0, "vae.cenet_decoder.0", 64, 531);
  { int s[3] = {0, 16, 19}, d[3] = {XD_Z, XD_MU, 0}, n[3] = {16, 3, 512}; expose(CD0, 3, s, d, n); }
  add(CD2, "vae.cenet_decoder.2", 128, 64); expose(CD2);
  add(CD4, "vae.cenet_decoder.4", 53, 128); expose(CD4);
  add(TD0, "vae.terrain_decoder.0", 512, 512); expose(TD0);
  add(TD2, "vae.terrain_decoder.2", 512, 512); expose(TD2);
  add(TD4, "vae.terrain_decoder.4", 693, 512); expose(TD4);
  // --- shared by both optimizer steps
  g_pol_begin = cur;
  add(CE0, "vae.cenet_encoder.0", 128, 265); expose(CE0);
  add(CE2, "vae.cenet_encoder.2", 64, 128); expose(CE2);
  add(LAT, "vae.latent", 35, 64);  // rows 0..18 latent_mu, 19..34 latent_var: one GEMM
  push_param("vae.latent_mu.weight", g_layers[LAT].w, 19, 64, 64);
  push_param("vae.latent_mu.bias", g_layers[LAT].b, 1, 19, 19);
  push_param("vae.latent_var.weight", g_layers[LAT].w + 19 * 64, 16, 64, 64);
  push_param("vae.latent_var.bias", g_layers[LAT].b + 19, 1, 16, 16);
  add(TE0, "vae.terrain_encoder.0", 512, 693); expose(TE0);
  add(TE2, "vae.terrain_encoder.2", 512, 512); expose(TE2);
  add(TE4, "vae.terrain_encoder.4", 512, 512); expose(TE4);
  g_vae_end = cur;
  // --- policy-step only
  add(AB0, "actor_body.0", 512, 584, LD_XA);
  { int s[4] = {0, 53, 69, 72}, d[4] = {XA_OBS, XA_Z, XA_MU, 0}, n[4] = {53, 16, 3, 512}; expose(AB0, 4, s, d, n); }
  add(AB2, "actor_body.2", 256, 512); expose(AB2);
  add(AB4, "actor_body.4", 128, 256); expose(AB4);
  add(AB6, "actor_body.6", 12, 128); expose(AB6);
  add(CB0, "critic_body.0", 512, 752);
  { int s[3] = {0, 53, 56}, d[3] = {XC_OBS, XC_BV, 0}, n[3] = {53, 3, 696}; expose(CB0, 3, s, d, n); }
  add(CB2, "critic_body.2", 256, 512); expose(CB2);
  add(CB4, "critic_body.4", 128, 256); expose(CB4);
  add(CB6, "critic_body.6", 1, 128); expose(CB6);
  g_off_std = cur;
  push_param("std", cur, 1, 12, 12);
  cur += 12;
  g_pol_end = cur;
  g_off_piggy = cur;
  cur += NPIGGY;
  // --- parameters that never receive a gradient on the training path
  add(MM0, "vae.memory_mlp.0", 256, 777);
  { int s[2] = {0, 265}, d[2] = {512, 0}, n[2] = {265, 512}; expose(MM0, 2, s, d, n); }
  add(MM2, "vae.memory_mlp.2", 128, 256); expose(MM2);
  add(MM4, "vae.memory_mlp.4", 512, 128); expose(MM4);
  add(GB0, "vae.gb_encoder.0", 128, 128); expose(GB0);
  add(GB2, "vae.gb_encoder.2", 64, 128); expose(GB2);
  g_total = cur;
}

extern "C" int dtc_param_count(void) { build_table(); return (int)g_params.size(); }
extern "C" int dtc_param_get(int i, dtc_param_info* out) {
  build_table();
  if (i < 0 || i >= (int)g_params.size() || !out) DTC_FAIL(DTC_ERR_ARG, "dtc_param_get: bad index %d", i);
  *out = g_params[i];
  return DTC_OK;
}
extern "C" int64_t dtc_param_total_floats(void) { build_table(); return g_total; }
extern "C" int dtc_param_range(int which, int64_t* begin, int64_t* end) {
  build_table();
  if (!begin || !end) DTC_FAIL(DTC_ERR_ARG, "dtc_param_range: null output");
  switch (which) {
    case 0: *begin = 0; *end = g_vae_end; break;
    case 1: *begin = g_pol_begin; *end = g_pol_end; break;
    case 2: *begin = g_pol_begin; *end = g_off_piggy + NPIGGY; break;
    case 3: *begin = 0; *end = g_total; break;
    default: DTC_FAIL(DTC_ERR_ARG, "dtc_param_range: unknown range %d", which);
  }
  return DTC_OK;
}

// ------------------------------------------------------------------ device-side state
struct LvStat {  // latent_var outlier repair (actor_critic_decoder.py:293-299)
  double sum, sumsq, out_grad_sum;
  unsigned long long n_in;
  float mean, lo, hi, median;
  unsigned int cnt_eq, key;
  unsigned int hist1[2048], hist2[2048], hist3[1024];
};
enum { ST_VALUE = 0, ST_SURR, ST_RECONS, ST_VEL, ST_KLD, ST_HEIGHT, ST_ENTROPY, ST_KL_MEAN, ST_LR, ST_GN_VAE, ST_GN_POL,
       ST_NVAE, ST_NPOL, ST_KL_SUM, ST_GN_SUMSQ, ST_COUNT = 16 };

#define WS_BUFFERS(X)                                                                                                     \
  X(XH, LD_HIST) X(XP, LD_PRIVA) X(XC, LD_XC) X(XM, LD_XM)                                                                 \
  X(H1, 128) X(E, 64) X(ML, LD_ML) X(EPS, 16) X(XA, LD_XA) X(XD, LD_XD) X(T1, 512) X(T2, 512)                              \
  X(D1, 64) X(D2, 128) X(REC, 56) X(U1, 512) X(U2, 512) X(HR, 696)                                                         \
  X(A1, 512) X(A2, 256) X(A3, 128) X(MEAN, 12) X(C1, 512) X(C2, 256) X(C3, 128) X(V, 4) X(EPSA, 12)                        \
  X(dML, LD_ML) X(dE, 64) X(dH1, 128) X(dT2, 512) X(dT1, 512) X(dX, LD_XA) X(dD1, 64) X(dD2, 128) X(dREC, 56)              \
  X(dU1, 512) X(dU2, 512) X(dHR, 696) X(dA1, 512) X(dA2, 256) X(dA3, 128) X(dMEAN, 12) X(dC1, 512) X(dC2, 256)             \
  X(dC3, 128) X(dV, 4) X(OUTM, 4)

struct dtc_learner {
  int R;
  float *params, *grads, *m_main, *v_main, *m_vae, *v_vae;
#define X(name, ld) float* name;
  WS_BUFFERS(X)
#undef X
  LvStat* lvstat;
  double* stats;
  float* skws;  // split-K partials
  float* csws;  // column-sum partials
  // bias gradients fused into the dgrad that produces their dY: per-32-row-block column sums from the tensor-core epilogue, one
  // buffer per stream that runs dgrads (0: caller's stream, 1: side stream c); bias_done[layer] tells wgrad() to skip its own colsum
  float* cspart[2];
  bool bias_done[NLAYERS];
  // 3xTF32 companions (dtc_gemm_tc.cu): every activation / gradient buffer has a twin at +lo_shift floats, the parameters
  // have params_lo, the packed batch rows have the *_lo arrays of dtc_storage (registered per call in ext[])
  float* ws_val_begin; float* ws_val_end; ptrdiff_t lo_shift;
  float* params_lo;
  struct { const float* base; size_t n; const float* lo; } ext[3];
  int64_t vae_steps, main_steps;
  int last_M;
  // intra-step concurrency: two non-blocking side streams forked from / joined into the caller's stream inside one call
  bool side_ready;
  cudaStream_t side[2];
  cudaEvent_t ev[DTC_NEV];
  int ev_next;
  // data parallel: gradient buckets of the last step, [which][bucket] (see dtc_learner_wait_bucket)
  bool bucket_ready;
  cudaEvent_t bucket_ev[2][2];
  const uint64_t* act_counter_base;  // optional device-side counter added to dtc_policy_act's Philox counter (CUDA-graph replays)
};

// Where the TF32 companions of ACTIVATIONS come from: 1 (default) = computed tile by tile in shared memory by the GEMM kernels'
// splitter warps (dtc_gemm_tc.cu: tc_split_tile) - no activation *_lo array is written or read; 0 (env DTC_TC_SPLIT=0) = round 1's
// layout, every activation buffer has a companion array in HBM filled by its producer.  Parameters keep their companion array
// (refreshed by the optimizer kernel) either way.
static int g_split_sm = -1;
static bool split_sm() {
  if (g_split_sm < 0) { const char* e = getenv("DTC_TC_SPLIT"); g_split_sm = (e && e[0] == '1') ? 1 : 0; }
  return g_split_sm != 0;
}
static const float* lo_of(const dtc_learner* l, const float* p) {
  if (!p) return nullptr;
  if (p >= l->params && p < l->params + g_total) return l->params_lo + (p - l->params);
  if (split_sm()) return nullptr;
  if (p >= l->ws_val_begin && p < l->ws_val_end) return p + l->lo_shift;
  if (p >= l->params && p < l->params + g_total) return l->params_lo + (p - l->params);
  for (int i = 0; i < 3; ++i)
    if (l->ext[i].base && p >= l->ext[i].base && p < l->ext[i].base + l->ext[i].n) return l->ext[i].lo ? l->ext[i].lo + (p - l->ext[i].base) : nullptr;
  return nullptr;
}
static float* lo_of(const dtc_learner* l, float* p) { return const_cast<float*>(lo_of(l, (const float*)p)); }

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t splitk_ws_floats(int R) {
  build_table();
  size_t mx = 0;
  for (int i = 0; i < NLAYERS; ++i) {
    const Layer& l = g_layers[i];
    size_t s = (size_t)dtc_gemm_pick_splits(l.out, l.ld, R) * l.out * round4(l.ld);
    if (s > mx) mx = s;
  }
  return mx;
}

static size_t ws_value_bytes(size_t R) {
  size_t tot = 0;
#define X(name, ld) tot += align256(R * (ld) * sizeof(float));
  WS_BUFFERS(X)
#undef X
  return tot;
}
extern "C" int64_t dtc_learner_workspace_bytes(int32_t max_rows) {
  build_table();
  size_t R = max_rows, tot = 0;
  tot += 2 * ws_value_bytes(R) + align256((size_t)g_total * sizeof(float));
  tot += align256(sizeof(LvStat)) + align256(ST_COUNT * sizeof(double));
  tot += align256(splitk_ws_floats(max_rows) * sizeof(float));
  tot += align256((size_t)COLSUM_CHUNKS * 768 * sizeof(float));
  tot += 2 * align256((size_t)((R + 31) / 32) * 768 * sizeof(float));
  return (int64_t)tot;
}

extern "C" int dtc_learner_create(int32_t max_rows, float* params, float* grads, float* adam_main_m, float* adam_main_v,
                                  float* adam_vae_m, float* adam_vae_v, void* workspace, int64_t workspace_bytes,
                                  void* stream, dtc_learner** out) {
  build_table();
  if (!out || !params || !workspace || max_rows <= 0) DTC_FAIL(DTC_ERR_ARG, "dtc_learner_create: bad arguments");
  if (workspace_bytes < dtc_learner_workspace_bytes(max_rows)) DTC_FAIL(DTC_ERR_ARG, "dtc_learner_create: workspace too small");
  if ((uintptr_t)workspace & 255) DTC_FAIL(DTC_ERR_ARG, "dtc_learner_create: workspace must be 256-byte aligned");
  dtc_learner* l = new dtc_learner();
  l->R = max_rows;
  l->params = params; l->grads = grads;
  l->m_main = adam_main_m; l->v_main = adam_main_v; l->m_vae = adam_vae_m; l->v_vae = adam_vae_v;
  char* p = (char*)workspace;
  size_t R = max_rows;
  l->ws_val_begin = (float*)p;
#define X(name, ld) l->name = (float*)p; p += align256(R * (ld) * sizeof(float));
  WS_BUFFERS(X)
#undef X
  l->ws_val_end = (float*)p;
  l->lo_shift = (ptrdiff_t)(ws_value_bytes(R) / sizeof(float));
  p += ws_value_bytes(R);
  l->params_lo = (float*)p; p += align256((size_t)g_total * sizeof(float));
  memset(l->ext, 0, sizeof(l->ext));
  l->lvstat = (LvStat*)p; p += align256(sizeof(LvStat));
  l->stats = (double*)p; p += align256(ST_COUNT * sizeof(double));
  l->skws = (float*)p; p += align256(splitk_ws_floats(max_rows) * sizeof(float));
  l->csws = (float*)p; p += align256((size_t)COLSUM_CHUNKS * 768 * sizeof(float));
  for (int i = 0; i < 2; ++i) { l->cspart[i] = (float*)p; p += align256((size_t)((R + 31) / 32) * 768 * sizeof(float)); }
  memset(l->bias_done, 0, sizeof(l->bias_done));
  l->vae_steps = l->main_steps = 0;
  l->last_M = 0;
  l->side_ready = false;
  l->bucket_ready = false;
  l->act_counter_base = nullptr;
  l->ev_next = 0;
  // ordered with the caller's own work (the parameter upload ahead of this call, the first step after it): no host sync
  DTC_CUDA(cudaMemsetAsync(l->stats, 0, ST_COUNT * sizeof(double), (cudaStream_t)stream));
  DTC_CUDA(cudaMemsetAsync(l->ws_val_begin, 0, 2 * ws_value_bytes(R), (cudaStream_t)stream));
  *out = l;
  return dtc_learner_refresh_params(l, stream);
}
extern "C" void dtc_learner_destroy(dtc_learner* l) {
  if (!l) return;
  if (l->side_ready) {
    for (int i = 0; i < 2; ++i) cudaStreamDestroy(l->side[i]);
    for (int i = 0; i < DTC_NEV; ++i) cudaEventDestroy(l->ev[i]);
  }
  if (l->bucket_ready)
    for (int i = 0; i < 4; ++i) cudaEventDestroy(l->bucket_ev[i >> 1][i & 1]);
  delete l;
}

// ------------------------------------------------------------------ intra-step concurrency
// One optimizer step is ~100 launches on three independent chains: the CENet encoder/decoder (GEMMs 12..128 wide, a
// handful of CTAs each), the weight gradients (a split-K GEMM + reduce + column sum per layer, needed only by the
// optimizer), and the 512-wide activation-gradient chain that is the critical path.  The step forks the first two onto side
// streams (events, no host synchronisation) so that they run in the shadow of the big tensor-core GEMMs and fill the tails of
// their persistent grids; everything is joined back into the caller's stream before the call returns.
static int g_overlap = -1;  // -1: not read yet (env DTC_OVERLAP=0 disables)
static int overlap_on() {
  if (g_overlap < 0) { const char* e = getenv("DTC_OVERLAP"); g_overlap = (e && e[0] == '0') ? 0 : 1; }
  return g_overlap;
}
extern "C" void dtc_set_overlap(int on) { g_overlap = on ? 1 : 0; }
extern "C" int dtc_get_overlap(void) { return overlap_on(); }

struct StepStreams { cudaStream_t main, w, c; };  // w: weight gradients + critic forward; c: CENet chains + critic backward

// work queued on `to` after this call waits for everything queued on `from` so far
static int chain(dtc_learner* l, cudaStream_t from, cudaStream_t to) {
  if (from == to) return DTC_OK;
  cudaEvent_t e = l->ev[l->ev_next];
  l->ev_next = (l->ev_next + 1) % DTC_NEV;
  DTC_CUDA(cudaEventRecord(e, from));
  DTC_CUDA(cudaStreamWaitEvent(to, e, 0));
  return DTC_OK;
}
static int streams_begin(dtc_learner* l, cudaStream_t st, StepStreams* S) {
  S->main = S->w = S->c = st;
  if (!overlap_on() || g_dtc_prof) return DTC_OK;  // per-launch profiling wants every kernel alone on the device
  if (!l->side_ready) {
    for (int i = 0; i < 2; ++i) DTC_CUDA(cudaStreamCreateWithFlags(&l->side[i], cudaStreamNonBlocking));
    for (int i = 0; i < DTC_NEV; ++i) DTC_CUDA(cudaEventCreateWithFlags(&l->ev[i], cudaEventDisableTiming));
    l->side_ready = true;
  }
  S->w = l->side[0];
  S->c = l->side[1];
  RETURN_IF_ERR(chain(l, st, S->w));
  return chain(l, st, S->c);
}
static int streams_end(dtc_learner* l, const StepStreams& S) {
  RETURN_IF_ERR(chain(l, S.w, S.main));
  return chain(l, S.c, S.main);
}

__global__ void __launch_bounds__(256) k_split_lo(const float* __restrict__ x, float* __restrict__ lo, int64_t n) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) lo[i] = tf32_lo(x[i]);
}
static int split_lo(const float* x, float* lo, int64_t n, cudaStream_t st) {
  if (!lo || n <= 0) return DTC_OK;
  int64_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  k_split_lo<<<(int)b, 256, 0, st>>>(x, lo, n);
  DTC_CHECK_LAUNCH("k_split_lo");
  return DTC_OK;
}
// recomputes the TF32 companions of ALL parameters; call after writing the flat parameter buffer from outside
extern "C" int dtc_learner_refresh_params(dtc_learner* l, void* stream) {
  if (!l) DTC_FAIL(DTC_ERR_ARG, "null learner");
  return split_lo(l->params, l->params_lo, g_total, (cudaStream_t)stream);
}
extern "C" double* dtc_learner_stats(dtc_learner* l) { return l ? l->stats : nullptr; }

__global__ void k_set_double(double* p, double v) { *p = v; }
extern "C" int dtc_learner_set_lr(dtc_learner* l, double lr, void* stream) {
  if (!l) DTC_FAIL(DTC_ERR_ARG, "null learner");
  k_set_double<<<1, 1, 0, (cudaStream_t)stream>>>(l->stats + ST_LR, lr);
  DTC_CHECK_LAUNCH("k_set_double");
  return DTC_OK;
}
extern "C" int dtc_learner_reset_stats(dtc_learner* l, void* stream) {
  if (!l) DTC_FAIL(DTC_ERR_ARG, "null learner");
  DTC_CUDA(cudaMemsetAsync(l->stats, 0, ST_KL_MEAN * sizeof(double), (cudaStream_t)stream));
  DTC_CUDA(cudaMemsetAsync(l->stats + ST_NVAE, 0, 2 * sizeof(double), (cudaStream_t)stream));
  return DTC_OK;
}
extern "C" int dtc_learner_set_act_counter_base(dtc_learner* l, const uint64_t* device_counter) {
  if (!l) DTC_FAIL(DTC_ERR_ARG, "dtc_learner_set_act_counter_base: null learner");
  l->act_counter_base = device_counter;
  return DTC_OK;
}
extern "C" int dtc_learner_set_adam_steps(dtc_learner* l, int64_t vae_steps, int64_t main_steps) {
  if (!l) DTC_FAIL(DTC_ERR_ARG, "null learner");
  l->vae_steps = vae_steps; l->main_steps = main_steps;
  return DTC_OK;
}
extern "C" int dtc_learner_get_adam_steps(dtc_learner* l, int64_t* vae_steps, int64_t* main_steps) {
  if (!l) DTC_FAIL(DTC_ERR_ARG, "null learner");
  if (vae_steps) *vae_steps = l->vae_steps;
  if (main_steps) *main_steps = l->main_steps;
  return DTC_OK;
}
extern "C" int dtc_learner_debug_buffer(dtc_learner* l, const char* name, float** ptr, int32_t* rows, int32_t* cols, int32_t* ld) {
  if (!l || !name) DTC_FAIL(DTC_ERR_ARG, "null argument");
#define X(nm, ldv) if (!strcmp(name, #nm)) { *ptr = l->nm; *rows = l->last_M; *cols = (ldv); *ld = (ldv); return DTC_OK; }
  WS_BUFFERS(X)
#undef X
  DTC_FAIL(DTC_ERR_ARG, "unknown debug buffer %s", name);
}

// ------------------------------------------------------------------ GEMM helpers
#define RET_IF(x) RETURN_IF_ERR(x)

// Gradient buffer zeroed once at the start of every optimizer step (default; env DTC_GRAD_PREZERO=0 restores per-GEMM handling): the
// split-K weight gradients and the bias column sums are then ADDED into it by L2 reductions straight from the GEMM epilogues - no
// per-GEMM memset, no partial-sum buffers, no reduce kernels.
static int g_prezero = -1;
static bool grads_prezeroed() {
  if (g_prezero < 0) { const char* e = getenv("DTC_GRAD_PREZERO"); g_prezero = (e && e[0] == '0') ? 0 : 1; }
  return g_prezero != 0;
}
// need_lo = false for outputs no GEMM reads (reconstructions, heads, latent statistics): their TF32 companion is never used
static int fwd(dtc_learner* l, int id, const float* A, int lda, float* C, int ldc, int act, int M, cudaStream_t st, bool need_lo = true) {
  const Layer& L = g_layers[id];
  GemmArgs g{};
  g.A = A; g.lda = lda; g.a_kc = true;
  g.B = l->params + L.w; g.ldb = L.ld; g.b_kc = true;
  g.C = C; g.ldc = ldc; g.M = M; g.N = L.out; g.K = L.in;
  g.A_lo = lo_of(l, A); g.B_lo = lo_of(l, g.B); g.C_lo = need_lo ? lo_of(l, C) : nullptr;
  g.a_split = split_sm();
  g.bias = l->params + L.b;
  g.epi = act == 1 ? EPI_BIAS_RELU : act == 2 ? EPI_BIAS_ELU : EPI_BIAS;
  g.splits = 1;
  return dtc_gemm_launch(g, st);
}
// dW[out,in] = dY^T X, db = colsum(dY)
static int wgrad(dtc_learner* l, int id, const float* dY, int ldy, const float* X, int ldx, int M, cudaStream_t st) {
  const Layer& L = g_layers[id];
  GemmArgs g{};
  g.A = dY; g.lda = ldy; g.a_kc = false;
  g.B = X; g.ldb = ldx; g.b_kc = false;
  g.A_lo = lo_of(l, dY); g.B_lo = lo_of(l, X);
  g.a_split = g.b_split = split_sm();
  g.c_zeroed = grads_prezeroed();
  g.C = l->grads + L.w; g.ldc = L.ld; g.M = L.out; g.N = L.in; g.K = M;
  g.epi = EPI_STORE;
  g.splits = dtc_gemm_pick_splits(L.out, L.in, M);
  g.ws = l->skws;
  RET_IF(dtc_gemm_launch(g, st));
  if (l->bias_done[id]) { l->bias_done[id] = false; return DTC_OK; }  // db came out of the dgrad that produced dY
  return dtc_colsum_launch(dY, ldy, M, L.out, l->grads + L.b, l->csws, st);
}
// dX[:, :ncols] (+)= (dY W[:, :ncols]) * act'(act_src)
// bias_layer >= 0: dX[:, :out(bias_layer)] is that layer's dY, so its bias gradient = column sums of dX; on the tensor-core path
// they come out of this GEMM's epilogue (slot: 0 when st is the caller's stream, 1 for side stream c)
static int dgrad(dtc_learner* l, int id, const float* dY, int ldy, float* dX, int lddx, int ncols, const float* act_src,
                 int ld_act, int epi, bool accumulate, int M, cudaStream_t st, int bias_layer = -1, int slot = 0) {
  const Layer& L = g_layers[id];
  GemmArgs g{};
  g.A = dY; g.lda = ldy; g.a_kc = true;
  g.B = l->params + L.w; g.ldb = L.ld; g.b_kc = false;
  g.A_lo = lo_of(l, dY); g.B_lo = lo_of(l, g.B); g.C_lo = lo_of(l, dX);
  g.a_split = split_sm();
  g.C = dX; g.ldc = lddx; g.M = M; g.N = ncols; g.K = L.out;
  g.act_src = act_src; g.ld_act = ld_act; g.epi = epi; g.accumulate = accumulate;
  g.splits = 1;
  const bool fuse = bias_layer >= 0 && dtc_gemm_mode() == 1 && dtc_gemm_tc_eligible(g) && g_layers[bias_layer].out <= ncols && ncols <= 768;
  if (fuse) {
    if (grads_prezeroed()) { g.colsum_out = l->grads + g_layers[bias_layer].b; g.colsum_n = g_layers[bias_layer].out; }  // added in L2
    else g.colsum_part = l->cspart[slot];
  }
  RET_IF(dtc_gemm_launch(g, st));
  if (fuse) {
    const Layer& Lb = g_layers[bias_layer];
    if (!grads_prezeroed()) RET_IF(dtc_colsum_part_launch(l->cspart[slot], (M + 31) / 32, round4(ncols), Lb.out, l->grads + Lb.b, st));
    l->bias_done[bias_layer] = true;
  }
  return DTC_OK;
}

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ unsigned int float_key(float x) {  // order-preserving map float -> uint32
  unsigned int b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
  unsigned int b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}
// standard normal draw number j of row `row` from Philox(seed, counter)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t counter, uint32_t row, int j) {
  Philox ph(seed, counter, ((uint64_t)row << 3) | (uint64_t)(j >> 2));
  uint4 r = ph.next();
  float2 n = (j & 2) ? box_muller(r.z, r.w) : box_muller(r.x, r.y);
  return (j & 1) ? n.y : n.x;
}

// block-wide: finds the bin holding rank `rank` (0-based) in hist[nbins]; returns bin and rank inside the bin.
// 256 threads; sh must hold 256 unsigned ints + 2.
__device__ void find_bin(const unsigned int* __restrict__ hist, int nbins, unsigned long long rank, unsigned int* sh,
                         unsigned int& bin, unsigned long long& rank_in) {
  const int per = nbins / 256;
  const int t = threadIdx.x;
  unsigned int loc = 0;
  for (int i = 0; i < per; ++i) loc += hist[t * per + i];
  sh[t] = loc;
  __syncthreads();
  if (t == 0) {
    unsigned long long run = 0;
    int sel = 255;
    for (int i = 0; i < 256; ++i) {
      if (rank < run + sh[i]) { sel = i; break; }
      run += sh[i];
    }
    unsigned int b = sel * per;
    for (int i = 0; i < per; ++i) {
      unsigned int h = hist[sel * per + i];
      if (rank < run + h || i == per - 1) { b = sel * per + i; break; }
      run += h;
    }
    sh[256] = b;
    sh[257] = (unsigned int)(rank - run);
  }
  __syncthreads();
  bin = sh[256];
  rank_in = sh[257];
  __syncthreads();
}

__device__ __forceinline__ void lv_bounds(const LvStat* s, long long n, float& mean, float& lo, float& hi) {
  double mean_d = s->sum / (double)n;
  double var = (s->sumsq - (double)n * mean_d * mean_d) / (double)(n - 1);
  var = var > 0.0 ? var : 0.0;
  mean = (float)mean_d;
  float thr = __fmul_rn(2.0f, (float)sqrt(var));
  lo = __fsub_rn(mean, thr);
  hi = __fadd_rn(mean, thr);
}

__global__ void __launch_bounds__(256) k_lv_moments(const float* __restrict__ ML, int M, LvStat* st) {
  __shared__ double sh[32];
  double s = 0.0, q = 0.0;
  const long long n = (long long)M * 16;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < n; e += gridDim.x * 256ll) {
    double x = (double)ML[(e >> 4) * LD_ML + ML_LV + (e & 15)];
    s += x; q += x * x;
  }
  s = block_sum(s, sh);
  q = block_sum(q, sh);
  if (threadIdx.x == 0) { atomicAdd(&st->sum, s); atomicAdd(&st->sumsq, q); }
}

template <int PASS>
__global__ void __launch_bounds__(256) k_lv_hist(const float* __restrict__ ML, int M, LvStat* st) {
  __shared__ unsigned int h[2048];
  __shared__ unsigned int sh[258];
  __shared__ unsigned int cnt_s;
  const long long n = (long long)M * 16;
  float mean, lo, hi;
  lv_bounds(st, n, mean, lo, hi);
  unsigned int b1 = 0, b2 = 0;
  unsigned long long r = 0;
  if (PASS >= 2) {
    unsigned long long k = (st->n_in - 1) / 2;
    find_bin(st->hist1, 2048, k, sh, b1, r);
    if (PASS >= 3) find_bin(st->hist2, 2048, r, sh, b2, r);
  }
  for (int i = threadIdx.x; i < 2048; i += 256) h[i] = 0;
  if (threadIdx.x == 0) cnt_s = 0;
  __syncthreads();
  unsigned int cnt = 0;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < n; e += gridDim.x * 256ll) {
    float x = ML[(e >> 4) * LD_ML + ML_LV + (e & 15)];
    if (x < lo || x > hi) continue;
    unsigned int key = float_key(x);
    if (PASS == 1) { atomicAdd(&h[key >> 21], 1u); ++cnt; }
    else if (PASS == 2) { if ((key >> 21) == b1) atomicAdd(&h[(key >> 10) & 2047u], 1u); }
    else { if ((key >> 10) == ((b1 << 11) | b2)) atomicAdd(&h[key & 1023u], 1u); }
  }
  if (PASS == 1) atomicAdd(&cnt_s, cnt);
  __syncthreads();
  unsigned int* gh = PASS == 1 ? st->hist1 : PASS == 2 ? st->hist2 : st->hist3;
  const int nb = PASS == 3 ? 1024 : 2048;
  for (int i = threadIdx.x; i < nb; i += 256)
    if (h[i]) atomicAdd(&gh[i], h[i]);
  if (PASS == 1 && threadIdx.x == 0) atomicAdd(&st->n_in, (unsigned long long)cnt_s);
}

__global__ void __launch_bounds__(256) k_lv_final(int M, LvStat* st) {
  __shared__ unsigned int sh[258];
  const long long n = (long long)M * 16;
  unsigned int b1, b2, b3;
  unsigned long long r;
  find_bin(st->hist1, 2048, (st->n_in - 1) / 2, sh, b1, r);
  find_bin(st->hist2, 2048, r, sh, b2, r);
  find_bin(st->hist3, 1024, r, sh, b3, r);
  if (threadIdx.x == 0) {
    float mean, lo, hi;
    lv_bounds(st, n, mean, lo, hi);
    unsigned int key = (b1 << 21) | (b2 << 10) | b3;
    st->mean = mean; st->lo = lo; st->hi = hi;
    st->key = key;
    st->median = key_float(key);
    st->cnt_eq = st->hist3[b3];
  }
}

// Packs caller tensors into the GEMM-ready row layouts (see include/dtc_b200.h: dtc_storage).  Any destination may be
// NULL.  One thread per destination float.
__global__ void __launch_bounds__(256) k_pack_inputs(int M, const float* __restrict__ obs, int obs_ld, const float* __restrict__ hist,
                                                     int hist_ld, const float* __restrict__ priv, int priv_ld,
                                                     const float* __restrict__ bv, int bv_ld, float* __restrict__ xh,
                                                     float* __restrict__ xp, float* __restrict__ xc, float* __restrict__ xh_lo,
                                                     float* __restrict__ xp_lo, float* __restrict__ xc_lo) {
  // one thread per 16-byte piece of a packed row (67 + 174 + 188 pieces): 32-bit index arithmetic, 16-byte stores; the sources are
  // read with scalar loads (the caller's row pitches and the 693-column offset of the critic part are not 16-byte multiples)
  constexpr int W4 = (LD_HIST + LD_PRIVA + LD_XC) / 4, H4 = LD_HIST / 4, P4 = LD_PRIVA / 4;
  const unsigned total = (unsigned)M * W4;
  for (unsigned e = blockIdx.x * 256u + threadIdx.x; e < total; e += gridDim.x * 256u) {
    const unsigned m = e / W4;
    int c = (int)(e - m * W4);
    float v[4];
    float *dst, *dst_lo;
    if (c < H4) {
      if (!xh) continue;
      const float* src = hist + (size_t)m * hist_ld + 4 * c;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = 4 * c + k < 265 ? src[k] : 0.f;
      dst = xh + (size_t)m * LD_HIST + 4 * c; dst_lo = xh_lo ? xh_lo + (size_t)m * LD_HIST + 4 * c : nullptr;
    } else if (c < H4 + P4) {
      c -= H4;
      if (!xp) continue;
      const float* src = priv + (size_t)m * priv_ld + 4 * c;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = 4 * c + k < 693 ? src[k] : 0.f;
      dst = xp + (size_t)m * LD_PRIVA + 4 * c; dst_lo = xp_lo ? xp_lo + (size_t)m * LD_PRIVA + 4 * c : nullptr;
    } else {
      c -= H4 + P4;
      if (!xc) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int col = 4 * c + k;
        v[k] = col < XC_OBS ? priv[(size_t)m * priv_ld + 693 + col] : col < XC_BV ? obs[(size_t)m * obs_ld + (col - XC_OBS)] : bv[(size_t)m * bv_ld + (col - XC_BV)];
      }
      dst = xc + (size_t)m * LD_XC + 4 * c; dst_lo = xc_lo ? xc_lo + (size_t)m * LD_XC + 4 * c : nullptr;
    }
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    if (dst_lo) *reinterpret_cast<float4*>(dst_lo) = make_float4(tf32_lo(v[0]), tf32_lo(v[1]), tf32_lo(v[2]), tf32_lo(v[3]));
  }
}

// Outlier repair of logvar in place, reparameterisation z = mu[3:] + eps * exp(0.5 logvar), and assembly of the
// non-GEMM columns of the actor input xa (mode 0) or the decoder input xd (mode 1).
// (actor_critic_decoder.py:293-301, :431, ppo.py:201).  8 rows per 128-thread block.
__global__ void __launch_bounds__(128) k_latent_fwd(int M, int mode, float* __restrict__ ML, const LvStat* __restrict__ st,
                                                    const float* __restrict__ eps_in, uint64_t seed, uint64_t counter,
                                                    float* __restrict__ EPS, float* __restrict__ OUTM, const float* __restrict__ xc,
                                                    float* __restrict__ X, float* __restrict__ X_lo, const uint64_t* __restrict__ cbase) {
  if (cbase) counter += 2ull * *cbase;  // dtc_policy_act under a CUDA graph: the launch carries a counter relative to a device-side one
  const int r8 = threadIdx.x >> 4, j = threadIdx.x & 15;
  const int row = blockIdx.x * 8 + r8;
  const int ldx = mode == 0 ? LD_XA : LD_XD, zoff = mode == 0 ? XA_Z : XD_Z, muoff = mode == 0 ? XA_MU : XD_MU;
  if (row < M) {
    float* ml = ML + (size_t)row * LD_ML;
    float lv = ml[ML_LV + j];
    bool outl = (lv < st->lo) || (lv > st->hi);
    if (outl) { lv = st->median; ml[ML_LV + j] = lv; }
    reinterpret_cast<unsigned char*>(OUTM)[(size_t)row * 16 + j] = outl ? 1 : 0;
    float e = eps_in ? eps_in[(size_t)row * 16 + j] : philox_normal(seed, counter, (uint32_t)row, j);
    EPS[(size_t)row * 16 + j] = e;
    float sd = expf(__fmul_rn(0.5f, lv));
    float z = __fadd_rn(__fmul_rn(e, sd), ml[3 + j]);
    float* x = X + (size_t)row * ldx;
    float* xl = X_lo ? X_lo + (size_t)row * ldx : nullptr;  // the l_t columns' companions come from the TE4 GEMM epilogue
    x[zoff + j] = z;
    if (xl) xl[zoff + j] = tf32_lo(z);
    if (j < 4) {
      const float mu = j < 3 ? ml[j] : 0.f;
      x[muoff + j] = mu;
      if (xl) xl[muoff + j] = tf32_lo(mu);
    }
    if (j == 0) ml[LD_ML - 1] = 0.f;
  }
  if (mode == 0) {
    for (int e = threadIdx.x; e < 8 * 56; e += 128) {
      int rr = blockIdx.x * 8 + e / 56, c = e % 56;
      if (rr < M) {
        const float v = c < 53 ? xc[(size_t)rr * LD_XC + XC_OBS + c] : 0.f;
        X[(size_t)rr * LD_XA + XA_OBS + c] = v;
        if (X_lo) X_lo[(size_t)rr * LD_XA + XA_OBS + c] = tf32_lo(v);
      }
    }
  }
}

// Backward of k_latent_fwd: gathers d z / d mu[:3] from dX, adds the direct loss terms already in dML (VAE step), and
// routes the gradient of repaired outliers to the median element (index_put_ + median() autograd semantics).
__global__ void __launch_bounds__(128) k_latent_bwd(int M, int mode, int has_direct, const float* __restrict__ ML,
                                                    const float* __restrict__ EPS, const float* __restrict__ OUTM,
                                                    const float* __restrict__ dX, float* __restrict__ dML, LvStat* st) {
  __shared__ double sh[32];
  const int r8 = threadIdx.x >> 4, j = threadIdx.x & 15;
  const int row = blockIdx.x * 8 + r8;
  const int ldx = mode == 0 ? LD_XA : LD_XD, zoff = mode == 0 ? XA_Z : XD_Z, muoff = mode == 0 ? XA_MU : XD_MU;
  double routed = 0.0;
  if (row < M) {
    const float* dx = dX + (size_t)row * ldx;
    float* d = dML + (size_t)row * LD_ML;
    float lv = ML[(size_t)row * LD_ML + ML_LV + j];
    float dz = dx[zoff + j];
    float sd = expf(__fmul_rn(0.5f, lv));
    float dlv = dz * EPS[(size_t)row * 16 + j] * sd * 0.5f;
    float dmu = dz;
    if (has_direct) { dlv += d[ML_LV + j]; dmu += d[3 + j]; }
    bool outl = reinterpret_cast<const unsigned char*>(OUTM)[(size_t)row * 16 + j] != 0;
    if (outl) { routed = (double)dlv; dlv = 0.f; }
    d[ML_LV + j] = dlv;
    d[3 + j] = dmu;
    if (j < 3) d[j] = (has_direct ? d[j] : 0.f) + dx[muoff + j];
    if (j == 3) d[LD_ML - 1] = 0.f;
  }
  routed = block_sum(routed, sh);
  if (threadIdx.x == 0 && routed != 0.0) atomicAdd(&st->out_grad_sum, routed);
}
__global__ void __launch_bounds__(256) k_lv_median_grad(int M, const float* __restrict__ ML, const float* __restrict__ OUTM,
                                                        float* __restrict__ dML, const LvStat* __restrict__ st) {
  const double S = st->out_grad_sum;
  if (S == 0.0 || st->cnt_eq == 0) return;
  const float add = (float)(S / (double)st->cnt_eq);
  const float med = st->median;
  const long long n = (long long)M * 16;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < n; e += gridDim.x * 256ll) {
    size_t row = (size_t)(e >> 4);
    int j = (int)(e & 15);
    if (reinterpret_cast<const unsigned char*>(OUTM)[e] == 0 && ML[row * LD_ML + ML_LV + j] == med) dML[row * LD_ML + ML_LV + j] += add;
  }
}

#define LOG_SQRT_2PI 0.91893853320467274178f

// PPO.act head (ppo.py:141-154): sample, log-prob, copies.  One thread per row.
__global__ void __launch_bounds__(128) k_act_head(int M, const float* __restrict__ MEAN, const float* __restrict__ V,
                                                  const float* __restrict__ stdp, const float* __restrict__ eps_in, uint64_t seed,
                                                  uint64_t counter, float* __restrict__ o_act, float* __restrict__ o_val,
                                                  float* __restrict__ o_logp, float* __restrict__ o_mu, float* __restrict__ o_sig,
                                                  float* __restrict__ x_act, float* __restrict__ x_val, float* __restrict__ x_logp,
                                                  float* __restrict__ x_mu, float* __restrict__ x_sig, const uint64_t* __restrict__ cbase) {
  if (cbase) counter += 2ull * *cbase;
  const int row = blockIdx.x * 128 + threadIdx.x;
  if (row >= M) return;
  float logp = 0.f;
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    float mu = MEAN[(size_t)row * 12 + j];
    float sg = __fadd_rn(__fmul_rn(mu, 0.0f), stdp[j]);
    float e = eps_in ? eps_in[(size_t)row * 12 + j] : philox_normal(seed, counter, (uint32_t)row, j);
    float a = __fadd_rn(mu, __fmul_rn(sg, e));
    float d = __fsub_rn(a, mu);
    float var = __fmul_rn(sg, sg);
    float lp = __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.0f, var)), logf(sg)), LOG_SQRT_2PI);
    logp = __fadd_rn(logp, lp);
    if (o_act) o_act[(size_t)row * 12 + j] = a;
    if (o_mu) o_mu[(size_t)row * 12 + j] = mu;
    if (o_sig) o_sig[(size_t)row * 12 + j] = sg;
    if (x_act) x_act[(size_t)row * 12 + j] = a;
    if (x_mu) x_mu[(size_t)row * 12 + j] = mu;
    if (x_sig) x_sig[(size_t)row * 12 + j] = sg;
  }
  float v = V[(size_t)row * 4];
  if (o_val) o_val[row] = v;
  if (o_logp) o_logp[row] = logp;
  if (x_val) x_val[row] = v;
  if (x_logp) x_logp[row] = logp;
}

__global__ void k_copy_col0(int M, const float* __restrict__ V, int ld, float* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < M) out[r] = V[(size_t)r * ld];
}

// PPO losses and their gradients w.r.t. the actor mean, std and the value (ppo.py:288-327).  One thread per row.
__global__ void __launch_bounds__(128) k_ppo_loss(int M, float inv_rows, const float* __restrict__ MEAN, const float* __restrict__ V,
                                                  const float* __restrict__ stdp, const float* __restrict__ actions,
                                                  const float* __restrict__ old_val, const float* __restrict__ adv,
                                                  const float* __restrict__ ret, const float* __restrict__ old_logp,
                                                  const float* __restrict__ old_mu, const float* __restrict__ old_sig,
                                                  dtc_ppo_hparams hp, float* __restrict__ dMEAN, float* __restrict__ dV,
                                                  float* __restrict__ dstd, double* __restrict__ stats) {
  __shared__ double sh[32];
  const int row = blockIdx.x * 128 + threadIdx.x;
  double s_surr = 0.0, s_val = 0.0, s_ent = 0.0, s_kl = 0.0;
  float ds[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) ds[j] = 0.f;
  if (row < M) {
    float mu[12], sg[12], a[12];
    float logp = 0.f, ent = 0.f, kl = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      mu[j] = MEAN[(size_t)row * 12 + j];
      sg[j] = __fadd_rn(__fmul_rn(mu[j], 0.0f), stdp[j]);
      a[j] = actions[(size_t)row * 12 + j];
      float d = __fsub_rn(a[j], mu[j]);
      float var = __fmul_rn(sg[j], sg[j]);
      float ls = logf(sg[j]);
      logp = __fadd_rn(logp, __fsub_rn(__fsub_rn(__fdiv_rn(-__fmul_rn(d, d), __fmul_rn(2.0f, var)), ls), LOG_SQRT_2PI));
      ent = __fadd_rn(ent, __fadd_rn(0.5f + LOG_SQRT_2PI, ls));
      float os = old_sig[(size_t)row * 12 + j], om = old_mu[(size_t)row * 12 + j];
      float dm = __fsub_rn(om, mu[j]);
      float t = __fdiv_rn(__fadd_rn(__fmul_rn(os, os), __fmul_rn(dm, dm)), __fmul_rn(2.0f, var));
      kl = __fadd_rn(kl, __fsub_rn(__fadd_rn(logf(__fadd_rn(__fdiv_rn(sg[j], os), 1.e-5f)), t), 0.5f));
    }
    const float A = adv[row];
    const float ratio = expf(__fsub_rn(logp, old_logp[row]));
    const float lo = 1.0f - hp.clip_param, hi = 1.0f + hp.clip_param;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float s1 = __fmul_rn(-A, ratio), s2 = __fmul_rn(-A, rc);
    const bool inr = ratio >= lo && ratio <= hi;
    float w = s1 > s2 ? 1.0f : (s1 == s2 ? 0.5f + (inr ? 0.5f : 0.0f) : (inr ? 1.0f : 0.0f));
    const float dlogp = -A * w * ratio * inv_rows;
    // value loss
    const float v = V[(size_t)row * 4], vt = old_val[row], R = ret[row];
    float dv, vl;
    if (hp.use_clipped_value_loss) {
      const float diff = __fsub_rn(v, vt);
      const float vc = __fadd_rn(vt, fminf(fmaxf(diff, -hp.clip_param), hp.clip_param));
      const bool inv = diff >= -hp.clip_param && diff <= hp.clip_param;
      const float e1 = __fsub_rn(v, R), e2 = __fsub_rn(vc, R);
      const float l1 = __fmul_rn(e1, e1), l2 = __fmul_rn(e2, e2);
      vl = fmaxf(l1, l2);
      if (l1 > l2) dv = 2.0f * e1;
      else if (l1 == l2) dv = e1 + (inv ? e2 : 0.0f);
      else dv = inv ? 2.0f * e2 : 0.0f;
    } else {
      const float e1 = __fsub_rn(R, v);
      vl = __fmul_rn(e1, e1);
      dv = -2.0f * e1;
    }
    dV[(size_t)row * 4] = dv * hp.value_loss_coef * inv_rows;
    dV[(size_t)row * 4 + 1] = 0.f; dV[(size_t)row * 4 + 2] = 0.f; dV[(size_t)row * 4 + 3] = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      float d = a[j] - mu[j];
      float var = sg[j] * sg[j];
      dMEAN[(size_t)row * 12 + j] = dlogp * d / var;
      ds[j] = dlogp * (d * d / (var * sg[j]) - 1.0f / sg[j]) - hp.entropy_coef * inv_rows / sg[j];
    }
    s_surr = (double)fmaxf(s1, s2);
    s_val = (double)vl;
    s_ent = (double)ent;
    s_kl = (double)kl;
  }
  s_surr = block_sum(s_surr, sh);
  s_val = block_sum(s_val, sh);
  s_ent = block_sum(s_ent, sh);
  s_kl = block_sum(s_kl, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&stats[ST_SURR], s_surr * (double)inv_rows);
    atomicAdd(&stats[ST_VALUE], s_val * (double)inv_rows);
    atomicAdd(&stats[ST_ENTROPY], s_ent * (double)inv_rows);
    atomicAdd(&stats[ST_KL_SUM], s_kl);
  }
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    double t = block_sum((double)ds[j], sh);
    if (threadIdx.x == 0) atomicAdd(&dstd[j], (float)t);
  }
}

// VAE losses (ppo.py:213-247): reconstruction, velocity, KL.  Writes dREC and the direct terms of dML.  Element-wise over the
// [M, 56] and [M, 36] arrays with coalesced accesses (one thread per row walked 224-byte strides: 0.6 ms under the GEMMs); every
// term of the three sums belongs to exactly one element, fp32 products summed in fp64.
__global__ void __launch_bounds__(256) k_vae_loss_rows(int M, float inv_rows, const float* __restrict__ REC,
                                                       const float* __restrict__ next_obs, const float* __restrict__ ML,
                                                       const float* __restrict__ xc, float* __restrict__ dREC,
                                                       float* __restrict__ dML, double* __restrict__ stats) {
  __shared__ double sh[32];
  double s_rec = 0.0, s_vel = 0.0, s_kld = 0.0;
  const float cr = 2.0f * inv_rows / 53.0f, cv = 2.0f * inv_rows / 3.0f;
  const long long stride = (long long)gridDim.x * 256, t0 = blockIdx.x * 256ll + threadIdx.x;
  for (long long e4 = t0; e4 < (long long)M * 14; e4 += stride) {  // 56 = 14 float4 per row, rows 16-byte aligned
    const int j = (int)(e4 % 14) * 4;
    const float4 r = *reinterpret_cast<const float4*>(REC + e4 * 4), x = *reinterpret_cast<const float4*>(next_obs + e4 * 4);
    float d[4] = {r.x - x.x, r.y - x.y, r.z - x.z, r.w - x.w};
    if (j + 3 >= 53) d[3] = 0.f;
    if (j + 2 >= 53) d[2] = 0.f;
    if (j + 1 >= 53) d[1] = 0.f;
    s_rec += (double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3]);
    *reinterpret_cast<float4*>(dREC + e4 * 4) = make_float4(cr * d[0], cr * d[1], cr * d[2], cr * d[3]);
  }
  for (long long e = t0; e < (long long)M * LD_ML; e += stride) {
    const int row = (int)(e / LD_ML), j = (int)(e - (long long)row * LD_ML);
    if (j >= ML_LV + 16) continue;  // the pad column
    const float v = ML[e];
    if (j < 3) {
      const float dv = v - xc[(size_t)row * LD_XC + XC_BV + j];
      s_vel += (double)(dv * dv);
      dML[e] = cv * dv;
    } else if (j < ML_LV) {
      s_kld += (double)(-v * v);
      dML[e] = 4.0f * inv_rows * v;
    } else {
      const float ex = expf(v);
      s_kld += (double)(1.0f + v - ex);
      dML[e] = -2.0f * inv_rows * (1.0f - ex);
    }
  }
  s_rec = block_sum(s_rec, sh);
  s_vel = block_sum(s_vel, sh);
  s_kld = block_sum(s_kld, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&stats[ST_RECONS], s_rec / 53.0 * (double)inv_rows);
    atomicAdd(&stats[ST_VEL], s_vel / 3.0 * (double)inv_rows);
    atomicAdd(&stats[ST_KLD], -0.5 * s_kld * (double)inv_rows);
  }
}
// height reconstruction loss (ppo.py:218-223): mse(terrain_decoder(l_t), priv[:, 696:]) and its gradient
__global__ void __launch_bounds__(256) k_vae_loss_height(int M, float inv_count, const float* __restrict__ HR,
                                                         const float* __restrict__ xc, float* __restrict__ dHR,
                                                         float* __restrict__ dHR_lo, double* __restrict__ stats) {
  __shared__ double sh[32];
  const long long total4 = (long long)M * 174;  // 696 = 174 float4 per row; HR / dHR rows are 16-byte aligned
  double s = 0.0;
  const float c = 2.0f * inv_count;
  for (long long e4 = blockIdx.x * 256ll + threadIdx.x; e4 < total4; e4 += gridDim.x * 256ll) {
    const int m = (int)(e4 / 174), j = (int)(e4 - (long long)m * 174) * 4;
    const float4 h = *reinterpret_cast<const float4*>(HR + e4 * 4);
    const float* x = xc + (size_t)m * LD_XC + 3 + j;  // the target starts at column 3 of the packed critic row: scalar loads
    float d[4] = {h.x - x[0], h.y - x[1], h.z - x[2], 0.f};
    if (j + 3 < 693) d[3] = h.w - x[3];
    if (j + 2 >= 693) d[2] = 0.f;
    if (j + 1 >= 693) d[1] = 0.f;
    float part = 0.f;  // four squares in fp32, the running sum in fp64 (as before to ~1e-7 relative)
#pragma unroll
    for (int k = 0; k < 4; ++k) part = fmaf(d[k], d[k], part);
    s += (double)part;
    const float4 g = make_float4(c * d[0], c * d[1], c * d[2], c * d[3]);
    *reinterpret_cast<float4*>(dHR + e4 * 4) = g;
    if (dHR_lo) *reinterpret_cast<float4*>(dHR_lo + e4 * 4) = make_float4(tf32_lo(g.x), tf32_lo(g.y), tf32_lo(g.z), tf32_lo(g.w));
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(&stats[ST_HEIGHT], s * (double)inv_count);
}

// sum of squares of grads[begin,end) -> stats[ST_GN_SUMSQ]
__global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    double x = (double)g[i];
    s += x * x;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
// publishes the local KL sum as a float next to the gradients so that it rides in the gradient all-reduce
__global__ void k_publish_kl(double* stats, float* piggy) { piggy[0] = (float)stats[ST_KL_SUM]; stats[ST_KL_SUM] = 0.0; }
// adaptive learning rate (ppo.py:295-307) from the (all-reduced) KL sum; finalises the gradient norm
__global__ void k_step_scalars(double* stats, const float* piggy, int which, float grad_scale, int rows_global, dtc_ppo_hparams hp) {
  double norm = sqrt(stats[ST_GN_SUMSQ]) * (double)grad_scale;
  stats[which == 0 ? ST_GN_VAE : ST_GN_POL] = norm;
  stats[which == 0 ? ST_NVAE : ST_NPOL] += 1.0;
  if (which == 1) {
    const float kl_mean = piggy[0] / (float)rows_global;
    stats[ST_KL_MEAN] = (double)kl_mean;
    if (hp.adaptive_lr) {
      double lr = stats[ST_LR];
      if (kl_mean > hp.desired_kl * 2.0f) lr = fmax(1e-5, lr / 1.5);
      else if (kl_mean < hp.desired_kl / 2.0f && kl_mean > 0.0f) lr = fmin(1e-2, lr * 1.5);
      stats[ST_LR] = lr;
    }
  }
}
// clip_grad_norm_ + Adam (torch.optim.Adam single-tensor formulas) over a flat range
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, float* __restrict__ p_lo, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, int64_t n, const double* __restrict__ stats, int which,
                                              float grad_scale, float max_norm, double lr_fixed, double bc1, double bc2_sqrt) {
  const double norm = stats[which == 0 ? ST_GN_VAE : ST_GN_POL];
  float coef = (float)((double)max_norm / (norm + 1e-6));
  coef = coef > 1.0f ? 1.0f : coef;
  const float gs = grad_scale;
  const double lr = which == 0 ? lr_fixed : stats[ST_LR];
  const float step_size = (float)(lr / bc1);
  const float bcs = (float)bc2_sqrt;
  const float b1 = 0.9f, b2 = 0.999f, w1 = (float)(1.0 - 0.9), w2 = (float)(1.0 - 0.999);
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    float gi = __fmul_rn(__fmul_rn(g[i], gs), coef);
    float mi = m[i], vi = v[i];
    mi = __fadd_rn(mi, __fmul_rn(w1, __fsub_rn(gi, mi)));
    vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(__fmul_rn(w2, gi), gi));
    m[i] = mi; v[i] = vi;
    float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bcs), 1e-8f);
    const float pn = __fsub_rn(p[i], __fmul_rn(step_size, __fdiv_rn(mi, denom)));
    p[i] = pn;
    p_lo[i] = tf32_lo(pn);
    (void)b1;
  }
}

// ------------------------------------------------------------------ storage kernels
__global__ void __launch_bounds__(256) k_store_env(dtc_storage s, int step, const float* __restrict__ rewards,
                                                   const uint8_t* __restrict__ dones, const uint8_t* __restrict__ time_outs,
                                                   const float* __restrict__ next_obs, int ld, float gamma) {
  const int N = s.N;
  const size_t base = (size_t)step * N;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < N * 57; e += gridDim.x * 256) {
    int n = e / 57, c = e - n * 57;
    if (c < 56) {
      s.next_obs[(base + n) * 56 + c] = c < 53 ? next_obs[(size_t)n * ld + c] : 0.f;
    } else {
      float to = time_outs && time_outs[n] ? 1.0f : 0.0f;
      float r = __fadd_rn(rewards[n], __fmul_rn(gamma, __fmul_rn(s.values[base + n], to)));
      s.rewards[base + n] = r;
      s.dones[base + n] = dones[n] ? 1 : 0;
    }
  }
}

// GAE reverse scan (rollout_storage.py:138-150), one thread per environment; raw advantages + their moments
__global__ void __launch_bounds__(128) k_gae(dtc_storage s, const float* __restrict__ last_values, float gamma, float lam,
                                             double* __restrict__ scratch) {
  __shared__ double sh[32];
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int N = s.N, T = s.T;
  double sum = 0.0, sq = 0.0;
  if (n < N) {
    float adv = 0.f;
    float nextv = last_values[n];
    for (int t = T - 1; t >= 0; --t) {
      const size_t i = (size_t)t * N + n;
      const float v = s.values[i];
      const float nt = 1.0f - (s.dones[i] ? 1.0f : 0.0f);
      const float delta = __fsub_rn(__fadd_rn(s.rewards[i], __fmul_rn(__fmul_rn(nt, gamma), nextv)), v);
      adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(nt, gamma), lam), adv));
      const float ret = __fadd_rn(adv, v);
      s.returns[i] = ret;
      const float raw = __fsub_rn(ret, v);
      s.advantages[i] = raw;
      sum += (double)raw;
      sq += (double)raw * (double)raw;
      nextv = v;
    }
  }
  sum = block_sum(sum, sh);
  sq = block_sum(sq, sh);
  if (threadIdx.x == 0) {
    atomicAdd(&scratch[0], sum);
    atomicAdd(&scratch[1], sq);
    if (blockIdx.x == 0) scratch[2] = (double)N * (double)T;
  }
}
__global__ void __launch_bounds__(256) k_gae_normalize(dtc_storage s, const double* __restrict__ st3) {
  const double cnt = st3[2];
  const double mean_d = st3[0] / cnt;
  double var = (st3[1] - cnt * mean_d * mean_d) / (cnt - 1.0);
  var = var > 0.0 ? var : 0.0;
  const float mean = (float)mean_d;
  const float den = __fadd_rn((float)sqrt(var), 1e-8f);
  const int64_t total = (int64_t)s.T * s.N;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll)
    s.advantages[i] = __fdiv_rn(__fsub_rn(s.advantages[i], mean), den);
}

// minibatch gather: one warp per destination row, float4 copies
__device__ __forceinline__ void warp_copy4(float* __restrict__ dst, const float* __restrict__ src, int n4, int lane) {
  const float4* s = reinterpret_cast<const float4*>(src);
  float4* d = reinterpret_cast<float4*>(dst);
  for (int i = lane; i < n4; i += 32) d[i] = __ldg(s + i);
}
__global__ void __launch_bounds__(256) k_gather(dtc_storage src, dtc_storage dst, const int64_t* __restrict__ perm, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (blockIdx.x * 256ll + threadIdx.x) >> 5, nw = (gridDim.x * 256ll) >> 5;
  for (int64_t r = wid; r < rows; r += nw) {
    const int64_t q = perm[r];
    warp_copy4(dst.hist + r * LD_HIST, src.hist + q * LD_HIST, LD_HIST / 4, lane);
    warp_copy4(dst.priv_a + r * LD_PRIVA, src.priv_a + q * LD_PRIVA, LD_PRIVA / 4, lane);
    warp_copy4(dst.xc + r * LD_XC, src.xc + q * LD_XC, LD_XC / 4, lane);
    warp_copy4(dst.next_obs + r * 56, src.next_obs + q * 56, 14, lane);
    warp_copy4(dst.actions + r * 12, src.actions + q * 12, 3, lane);
    warp_copy4(dst.mu + r * 12, src.mu + q * 12, 3, lane);
    warp_copy4(dst.sigma + r * 12, src.sigma + q * 12, 3, lane);
    if (lane == 0) dst.rewards[r] = src.rewards[q];
    if (lane == 1) dst.values[r] = src.values[q];
    if (lane == 2) dst.returns[r] = src.returns[q];
    if (lane == 3) dst.advantages[r] = src.advantages[q];
    if (lane == 4) dst.logp[r] = src.logp[q];
    if (lane == 5) dst.dones[r] = src.dones[q];
  }
}

// ------------------------------------------------------------------ host-side orchestration
static int grid1d(long long n, int threads, int cap = 148 * 8) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

// cenet encoder + latent heads + outlier statistics (stream c) | terrain encoder (main) -> latent kernel; X receives l_t | z | mu[:3]
// ------------------------------------------------------------------ data-parallel gradient buckets
// With sync_grads the caller all-reduces the step's flat gradient range before dtc_optimizer_apply.  Backward produces that
// range back to front: in the VAE step the decoders' weight gradients are complete while the encoders' backward still runs, in
// the policy step the actor / critic ones.  Each step therefore publishes TWO buckets with an event each, so the caller can
// start the all-reduce of the early one on a communication stream underneath the rest of the backward pass:
//   which 0 (VAE step)   : bucket 0 = decoders (ready on stream w after the last decoder wgrad), bucket 1 = shared encoders (end of call)
//   which 1 (policy step): bucket 0 = actor, critic, std, piggy-back scalars,                   bucket 1 = shared encoders (end of call)
// Ranges are contiguous in the flat layout, so a bucket is one all-reduce.
static int record_bucket(dtc_learner* l, int which, int bucket, cudaStream_t st) {
  if (!l->bucket_ready) {
    for (int i = 0; i < 4; ++i) DTC_CUDA(cudaEventCreateWithFlags(&l->bucket_ev[i >> 1][i & 1], cudaEventDisableTiming));
    l->bucket_ready = true;
  }
  DTC_CUDA(cudaEventRecord(l->bucket_ev[which][bucket], st));
  return DTC_OK;
}
extern "C" int dtc_learner_grad_bucket(int which, int bucket, int64_t* begin, int64_t* end) {
  build_table();
  if ((which != 0 && which != 1) || (bucket != 0 && bucket != 1) || !begin || !end) DTC_FAIL(DTC_ERR_ARG, "dtc_learner_grad_bucket: bad arguments");
  // VAE step: [0, g_vae_end) = decoders | shared encoders; the shared encoders finish last -> bucket 1
  // policy step: [g_pol_begin, piggy end) = shared encoders | actor, critic, std, piggy; the shared encoders finish last -> bucket 1
  if (which == 0) { *begin = bucket == 0 ? 0 : g_pol_begin; *end = bucket == 0 ? g_pol_begin : g_vae_end; }
  else { *begin = bucket == 0 ? g_vae_end : g_pol_begin; *end = bucket == 0 ? g_off_piggy + NPIGGY : g_vae_end; }
  return DTC_OK;
}
extern "C" int dtc_learner_wait_bucket(dtc_learner* l, int which, int bucket, void* stream) {
  if (!l || (which != 0 && which != 1) || (bucket != 0 && bucket != 1)) DTC_FAIL(DTC_ERR_ARG, "dtc_learner_wait_bucket: bad arguments");
  if (!l->bucket_ready) DTC_FAIL(DTC_ERR_STATE, "dtc_learner_wait_bucket: no step with sync_grads has run yet");
  DTC_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, l->bucket_ev[which][bucket], 0));
  return DTC_OK;
}

static int encode(dtc_learner* l, int M, const float* hist, const float* priv_a, const float* xc, int mode, const float* eps,
                  uint64_t seed, uint64_t counter, const StepStreams& S, const uint64_t* cbase = nullptr) {
  float* X = mode == 0 ? l->XA : l->XD;
  const int ldx = mode == 0 ? LD_XA : LD_XD;
  cudaStream_t st = S.main, sc = S.c;
  RET_IF(fwd(l, CE0, hist, LD_HIST, l->H1, 128, 1, M, sc));
  RET_IF(fwd(l, CE2, l->H1, 128, l->E, 64, 0, M, sc));
  RET_IF(fwd(l, LAT, l->E, 64, l->ML, LD_ML, 0, M, sc, false));
  DTC_CUDA(cudaMemsetAsync(l->lvstat, 0, sizeof(LvStat), sc));
  const int gb = grid1d((long long)M * 16, 256, 148 * 2);
  k_lv_moments<<<gb, 256, 0, sc>>>(l->ML, M, l->lvstat); DTC_CHECK_LAUNCH("k_lv_moments");
  k_lv_hist<1><<<gb, 256, 0, sc>>>(l->ML, M, l->lvstat); DTC_CHECK_LAUNCH("k_lv_hist1");
  k_lv_hist<2><<<gb, 256, 0, sc>>>(l->ML, M, l->lvstat); DTC_CHECK_LAUNCH("k_lv_hist2");
  k_lv_hist<3><<<gb, 256, 0, sc>>>(l->ML, M, l->lvstat); DTC_CHECK_LAUNCH("k_lv_hist3");
  k_lv_final<<<1, 256, 0, sc>>>(M, l->lvstat); DTC_CHECK_LAUNCH("k_lv_final");
  RET_IF(fwd(l, TE0, priv_a, LD_PRIVA, l->T1, 512, 1, M, st));
  RET_IF(fwd(l, TE2, l->T1, 512, l->T2, 512, 1, M, st));
  RET_IF(fwd(l, TE4, l->T2, 512, X, ldx, 0, M, st));
  RET_IF(chain(l, sc, st));
  k_latent_fwd<<<ceil_div(M, 8), 128, 0, st>>>(M, mode, l->ML, l->lvstat, eps, seed, counter, l->EPS, l->OUTM, xc, X, lo_of(l, X), cbase);
  DTC_CHECK_LAUNCH("k_latent_fwd");
  return DTC_OK;
}

// backward of encode(): dX holds d l_t | d z | d mu[:3] (layout of X); fills the gradients of CE0, CE2, LAT, TE0..TE4.
// Everything dX / dML depend on must already be ordered before S.main.
static int encode_bwd(dtc_learner* l, int M, const float* hist, const float* priv_a, int mode, int has_direct, const StepStreams& S) {
  const int ldx = mode == 0 ? LD_XA : LD_XD;
  cudaStream_t st = S.main, sw = S.w, sc = S.c;
  k_latent_bwd<<<ceil_div(M, 8), 128, 0, st>>>(M, mode, has_direct, l->ML, l->EPS, l->OUTM, l->dX, l->dML, l->lvstat);
  DTC_CHECK_LAUNCH("k_latent_bwd");
  k_lv_median_grad<<<grid1d((long long)M * 16, 256, 148 * 2), 256, 0, st>>>(M, l->ML, l->OUTM, l->dML, l->lvstat);
  DTC_CHECK_LAUNCH("k_lv_median_grad");
  RET_IF(split_lo(l->dML, lo_of(l, l->dML), (int64_t)M * LD_ML, st));
  RET_IF(chain(l, st, sw));
  RET_IF(chain(l, st, sc));
  RET_IF(wgrad(l, TE4, l->dX, ldx, l->T2, 512, M, sw));
  RET_IF(wgrad(l, LAT, l->dML, LD_ML, l->E, 64, M, sw));
  RET_IF(dgrad(l, LAT, l->dML, LD_ML, l->dE, 64, 64, nullptr, 0, EPI_STORE, false, M, sc, CE2, 1));
  RET_IF(chain(l, sc, sw));
  RET_IF(wgrad(l, CE2, l->dE, 64, l->H1, 128, M, sw));
  RET_IF(dgrad(l, TE4, l->dX, ldx, l->dT2, 512, 512, l->T2, 512, EPI_DRELU, false, M, st, TE2, 0));
  RET_IF(chain(l, st, sw));
  RET_IF(wgrad(l, TE2, l->dT2, 512, l->T1, 512, M, sw));
  RET_IF(dgrad(l, CE2, l->dE, 64, l->dH1, 128, 128, l->H1, 128, EPI_DRELU, false, M, sc, CE0, 1));
  RET_IF(chain(l, sc, sw));
  RET_IF(wgrad(l, CE0, l->dH1, 128, hist, LD_HIST, M, sw));
  RET_IF(dgrad(l, TE2, l->dT2, 512, l->dT1, 512, 512, l->T1, 512, EPI_DRELU, false, M, st, TE0, 0));
  RET_IF(chain(l, st, sw));
  RET_IF(wgrad(l, TE0, l->dT1, 512, priv_a, LD_PRIVA, M, sw));
  return DTC_OK;
}

static int actor_fwd(dtc_learner* l, int M, cudaStream_t st) {
  RET_IF(fwd(l, AB0, l->XA, LD_XA, l->A1, 512, 2, M, st));
  RET_IF(fwd(l, AB2, l->A1, 512, l->A2, 256, 2, M, st));
  RET_IF(fwd(l, AB4, l->A2, 256, l->A3, 128, 2, M, st));
  return fwd(l, AB6, l->A3, 128, l->MEAN, 12, 0, M, st, false);
}
static int critic_fwd(dtc_learner* l, int M, const float* xc, cudaStream_t st) {
  RET_IF(fwd(l, CB0, xc, LD_XC, l->C1, 512, 2, M, st));
  RET_IF(fwd(l, CB2, l->C1, 512, l->C2, 256, 2, M, st));
  RET_IF(fwd(l, CB4, l->C2, 256, l->C3, 128, 2, M, st));
  return fwd(l, CB6, l->C3, 128, l->V, 4, 0, M, st, false);
}

static int check_storage(const dtc_storage* s, const char* who) {
  if (!s || !s->hist || !s->priv_a || !s->xc || !s->next_obs || !s->actions || !s->mu || !s->sigma || !s->rewards ||
      !s->values || !s->returns || !s->advantages || !s->logp || !s->dones || s->T <= 0 || s->N <= 0)
    DTC_FAIL(DTC_ERR_ARG, "%s: incomplete dtc_storage", who);
  return DTC_OK;
}

extern "C" int dtc_policy_act(dtc_learner* l, int32_t M, const float* obs, int32_t obs_ld, const float* hist, int32_t hist_ld,
                              const float* priv, int32_t priv_ld, const float* base_vel, int32_t bv_ld, const float* eps_z,
                              const float* eps_a, uint64_t seed, uint64_t counter, const dtc_storage* s, int32_t step,
                              float* actions, float* values, float* logp, float* mean, float* sigma, void* stream) {
  DTC_NVTX("dtc_policy_act");
  if (!l || !obs || !hist || !priv || !base_vel) DTC_FAIL(DTC_ERR_ARG, "dtc_policy_act: null argument");
  if (M <= 0 || M > l->R) DTC_FAIL(DTC_ERR_ARG, "dtc_policy_act: M=%d outside (0,%d]", M, l->R);
  cudaStream_t st = (cudaStream_t)stream;
  float *xh = l->XH, *xp = l->XP, *xc = l->XC;
  float *o_act = nullptr, *o_val = nullptr, *o_logp = nullptr, *o_mu = nullptr, *o_sig = nullptr;
  if (s) {
    RET_IF(check_storage(s, "dtc_policy_act"));
    if (M != s->N) DTC_FAIL(DTC_ERR_ARG, "dtc_policy_act: M must equal storage N");
    if (step < 0 || step >= s->T) DTC_FAIL(DTC_ERR_STATE, "Rollout buffer overflow");
    const size_t b = (size_t)step * s->N;
    xh = s->hist + b * LD_HIST; xp = s->priv_a + b * LD_PRIVA; xc = s->xc + b * LD_XC;
    o_act = s->actions + b * 12; o_mu = s->mu + b * 12; o_sig = s->sigma + b * 12;
    o_val = s->values + b; o_logp = s->logp + b;
  }
  l->last_M = M;
  if (s) {
    const size_t b = (size_t)step * s->N;
    l->ext[0] = {xh, (size_t)M * LD_HIST, s->hist_lo ? s->hist_lo + b * LD_HIST : nullptr};
    l->ext[1] = {xp, (size_t)M * LD_PRIVA, s->priv_a_lo ? s->priv_a_lo + b * LD_PRIVA : nullptr};
    l->ext[2] = {xc, (size_t)M * LD_XC, s->xc_lo ? s->xc_lo + b * LD_XC : nullptr};
  } else {
    memset(l->ext, 0, sizeof(l->ext));
  }
  k_pack_inputs<<<grid1d((long long)M * ((LD_HIST + LD_PRIVA + LD_XC) / 4), 256), 256, 0, st>>>(M, obs, obs_ld, hist, hist_ld, priv, priv_ld,
                                                                                           base_vel, bv_ld, xh, xp, xc, lo_of(l, xh),
                                                                                           lo_of(l, xp), lo_of(l, xc));
  DTC_CHECK_LAUNCH("k_pack_inputs");
  StepStreams S;
  RET_IF(streams_begin(l, st, &S));
  RET_IF(critic_fwd(l, M, xc, S.w));
  RET_IF(encode(l, M, xh, xp, xc, 0, eps_z, seed, counter * 2, S, l->act_counter_base));
  RET_IF(actor_fwd(l, M, st));
  RET_IF(streams_end(l, S));
  k_act_head<<<ceil_div(M, 128), 128, 0, st>>>(M, l->MEAN, l->V, l->params + g_off_std, eps_a, seed, counter * 2 + 1, o_act, o_val,
                                              o_logp, o_mu, o_sig, actions, values, logp, mean, sigma, l->act_counter_base);
  DTC_CHECK_LAUNCH("k_act_head");
  return DTC_OK;
}

extern "C" int dtc_policy_evaluate(dtc_learner* l, int32_t M, const float* obs, int32_t obs_ld, const float* priv, int32_t priv_ld,
                                   const float* base_vel, int32_t bv_ld, float* values, void* stream) {
  DTC_NVTX("dtc_policy_evaluate");
  if (!l || !obs || !priv || !base_vel || !values) DTC_FAIL(DTC_ERR_ARG, "dtc_policy_evaluate: null argument");
  if (M <= 0 || M > l->R) DTC_FAIL(DTC_ERR_ARG, "dtc_policy_evaluate: M=%d outside (0,%d]", M, l->R);
  cudaStream_t st = (cudaStream_t)stream;
  k_pack_inputs<<<grid1d((long long)M * ((LD_HIST + LD_PRIVA + LD_XC) / 4), 256), 256, 0, st>>>(M, obs, obs_ld, nullptr, 0, priv, priv_ld,
                                                                                           base_vel, bv_ld, nullptr, nullptr, l->XC, nullptr,
                                                                                           nullptr, lo_of(l, l->XC));
  DTC_CHECK_LAUNCH("k_pack_inputs");
  RET_IF(critic_fwd(l, M, l->XC, st));
  k_copy_col0<<<ceil_div(M, 256), 256, 0, st>>>(M, l->V, 4, values);
  DTC_CHECK_LAUNCH("k_copy_col0");
  return DTC_OK;
}

// xm = [l_t | hist | 0], then b_t = b_t1 + l_t * b_t1 written over xa's l_t columns; latent -> xa z / mu columns
__global__ void __launch_bounds__(256) k_teacher_pack(int M, const float* __restrict__ XA, const float* __restrict__ XH,
                                                      float* __restrict__ XM) {
  const long long total = (long long)M * LD_XM;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    int m = (int)(e / LD_XM), c = (int)(e - (long long)m * LD_XM);
    XM[e] = c < 512 ? XA[(size_t)m * LD_XA + c] : XH[(size_t)m * LD_HIST + (c - 512)];
  }
}
__global__ void __launch_bounds__(256) k_teacher_gate(int M, const float* __restrict__ B1, const float* __restrict__ ML,
                                                      const float* __restrict__ xc, float* __restrict__ XA) {
  const long long total = (long long)M * LD_XA;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < total; e += gridDim.x * 256ll) {
    int m = (int)(e / LD_XA), c = (int)(e - (long long)m * LD_XA);
    float v;
    if (c < 512) { float b = B1[(size_t)m * 512 + c]; v = __fadd_rn(b, __fmul_rn(XA[e], b)); }
    else if (c < XA_OBS + 53) v = xc[(size_t)m * LD_XC + XC_OBS + (c - XA_OBS)];
    else if (c < XA_Z) v = 0.f;
    else if (c < XA_MU) v = ML[(size_t)m * LD_ML + 3 + (c - XA_Z)];
    else if (c < XA_MU + 3) v = ML[(size_t)m * LD_ML + (c - XA_MU)];
    else v = 0.f;
    XA[e] = v;
  }
}
extern "C" int dtc_policy_act_teacher(dtc_learner* l, int32_t M, const float* obs, int32_t obs_ld, const float* hist,
                                      int32_t hist_ld, const float* priv, int32_t priv_ld, float* actions, void* stream) {
  DTC_NVTX("dtc_policy_act_teacher");
  if (!l || !obs || !hist || !priv || !actions) DTC_FAIL(DTC_ERR_ARG, "dtc_policy_act_teacher: null argument");
  if (M <= 0 || M > l->R) DTC_FAIL(DTC_ERR_ARG, "dtc_policy_act_teacher: M=%d outside (0,%d]", M, l->R);
  cudaStream_t st = (cudaStream_t)stream;
  l->last_M = M;
  // base_vel is not an input of this path: the obs pointer doubles as a dummy source for the 3 base_vel columns
  k_pack_inputs<<<grid1d((long long)M * ((LD_HIST + LD_PRIVA + LD_XC) / 4), 256), 256, 0, st>>>(M, obs, obs_ld, hist, hist_ld, priv, priv_ld,
                                                                                           obs, obs_ld, l->XH, l->XP, l->XC, lo_of(l, l->XH),
                                                                                           lo_of(l, l->XP), nullptr);
  DTC_CHECK_LAUNCH("k_pack_inputs");
  RET_IF(fwd(l, CE0, l->XH, LD_HIST, l->H1, 128, 1, M, st));
  RET_IF(fwd(l, CE2, l->H1, 128, l->E, 64, 0, M, st));
  RET_IF(fwd(l, LAT, l->E, 64, l->ML, LD_ML, 0, M, st));
  RET_IF(fwd(l, TE0, l->XP, LD_PRIVA, l->T1, 512, 1, M, st));
  RET_IF(fwd(l, TE2, l->T1, 512, l->T2, 512, 1, M, st));
  RET_IF(fwd(l, TE4, l->T2, 512, l->XA, LD_XA, 0, M, st));
  k_teacher_pack<<<grid1d((long long)M * LD_XM, 256), 256, 0, st>>>(M, l->XA, l->XH, l->XM);
  DTC_CHECK_LAUNCH("k_teacher_pack");
  RET_IF(split_lo(l->XM, lo_of(l, l->XM), (int64_t)M * LD_XM, st));
  RET_IF(fwd(l, MM0, l->XM, LD_XM, l->C2, 256, 1, M, st));
  RET_IF(fwd(l, MM2, l->C2, 256, l->C3, 128, 1, M, st));
  RET_IF(fwd(l, MM4, l->C3, 128, l->C1, 512, 0, M, st));
  k_teacher_gate<<<grid1d((long long)M * LD_XA, 256), 256, 0, st>>>(M, l->C1, l->ML, l->XC, l->XA);
  DTC_CHECK_LAUNCH("k_teacher_gate");
  RET_IF(split_lo(l->XA, lo_of(l, l->XA), (int64_t)M * LD_XA, st));
  RET_IF(actor_fwd(l, M, st));
  DTC_CUDA(cudaMemcpyAsync(actions, l->MEAN, (size_t)M * 12 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return DTC_OK;
}

extern "C" int dtc_store_transition(const dtc_storage* s, int32_t step, const float* rewards, const uint8_t* dones,
                                    const uint8_t* time_outs, const float* next_obs, int32_t next_obs_ld, float gamma, void* stream) {
  DTC_NVTX("dtc_store_transition");
  RET_IF(check_storage(s, "dtc_store_transition"));
  if (step < 0 || step >= s->T) DTC_FAIL(DTC_ERR_STATE, "Rollout buffer overflow");
  if (!rewards || !dones || !next_obs) DTC_FAIL(DTC_ERR_ARG, "dtc_store_transition: null argument");
  k_store_env<<<grid1d((long long)s->N * 57, 256), 256, 0, (cudaStream_t)stream>>>(*s, step, rewards, dones, time_outs, next_obs,
                                                                                 next_obs_ld, gamma);
  DTC_CHECK_LAUNCH("k_store_env");
  return DTC_OK;
}

extern "C" int dtc_gae(const dtc_storage* s, const float* last_values, float gamma, float lam, double* scratch, int defer_normalize,
                       void* stream) {
  DTC_NVTX("dtc_gae");
  RET_IF(check_storage(s, "dtc_gae"));
  if (!last_values || !scratch) DTC_FAIL(DTC_ERR_ARG, "dtc_gae: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DTC_CUDA(cudaMemsetAsync(scratch, 0, 4 * sizeof(double), st));
  k_gae<<<ceil_div(s->N, 128), 128, 0, st>>>(*s, last_values, gamma, lam, scratch);
  DTC_CHECK_LAUNCH("k_gae");
  if (!defer_normalize) return dtc_gae_normalize(s, scratch, stream);
  return DTC_OK;
}
extern "C" int dtc_gae_normalize(const dtc_storage* s, const double* stats3, void* stream) {
  DTC_NVTX("dtc_gae_normalize");
  RET_IF(check_storage(s, "dtc_gae_normalize"));
  k_gae_normalize<<<grid1d((long long)s->T * s->N, 256), 256, 0, (cudaStream_t)stream>>>(*s, stats3);
  DTC_CHECK_LAUNCH("k_gae_normalize");
  return DTC_OK;
}

extern "C" int dtc_gather_minibatch(const dtc_storage* src, const dtc_storage* dst, const int64_t* perm, int64_t rows, void* stream) {
  DTC_NVTX("dtc_gather_minibatch");
  RET_IF(check_storage(src, "dtc_gather_minibatch(src)"));
  RET_IF(check_storage(dst, "dtc_gather_minibatch(dst)"));
  if (!perm || rows < 0 || rows > (int64_t)dst->T * dst->N) DTC_FAIL(DTC_ERR_ARG, "dtc_gather_minibatch: bad rows");
  if (rows == 0) return DTC_OK;
  k_gather<<<grid1d(rows * 32, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(*src, *dst, perm, rows);
  DTC_CHECK_LAUNCH("k_gather");
  cudaStream_t st = (cudaStream_t)stream;
  if (split_sm()) return DTC_OK;
  RET_IF(split_lo(dst->hist, dst->hist_lo, rows * LD_HIST, st));
  RET_IF(split_lo(dst->priv_a, dst->priv_a_lo, rows * LD_PRIVA, st));
  return split_lo(dst->xc, dst->xc_lo, rows * LD_XC, st);
}

static int optimizer_apply(dtc_learner* l, int which, const dtc_ppo_hparams* hp, float grad_scale, int rows_global, cudaStream_t st) {
  int64_t b, e;
  dtc_param_range(which == 0 ? 0 : 1, &b, &e);
  const int64_t n = e - b;
  DTC_CUDA(cudaMemsetAsync(l->stats + ST_GN_SUMSQ, 0, sizeof(double), st));
  k_sumsq<<<grid1d(n, 256, 148 * 4), 256, 0, st>>>(l->grads + b, n, l->stats + ST_GN_SUMSQ);
  DTC_CHECK_LAUNCH("k_sumsq");
  k_step_scalars<<<1, 1, 0, st>>>(l->stats, l->grads + g_off_piggy, which, grad_scale, rows_global, *hp);
  DTC_CHECK_LAUNCH("k_step_scalars");
  int64_t& steps = which == 0 ? l->vae_steps : l->main_steps;
  steps += 1;
  const double bc1 = 1.0 - pow(0.9, (double)steps), bc2 = 1.0 - pow(0.999, (double)steps);
  float* m = which == 0 ? l->m_vae : l->m_main;
  float* v = which == 0 ? l->v_vae : l->v_main;
  k_adam<<<grid1d(n, 256, 148 * 8), 256, 0, st>>>(l->params + b, l->params_lo + b, l->grads + b, m + b, v + b, n, l->stats, which, grad_scale,
                                                 hp->max_grad_norm, 5.e-4, bc1, sqrt(bc2));
  DTC_CHECK_LAUNCH("k_adam");
  return DTC_OK;
}
extern "C" int dtc_optimizer_apply(dtc_learner* l, int which, const dtc_ppo_hparams* hp, float grad_scale, int32_t rows_global,
                                   void* stream) {
  DTC_NVTX("dtc_optimizer_apply");
  if (!l || !hp || (which != 0 && which != 1)) DTC_FAIL(DTC_ERR_ARG, "dtc_optimizer_apply: bad arguments");
  if (!l->m_main || !l->v_main || !l->m_vae || !l->v_vae || !l->grads) DTC_FAIL(DTC_ERR_STATE, "learner was created without optimizer state");
  return optimizer_apply(l, which, hp, grad_scale, rows_global, (cudaStream_t)stream);
}

static int check_batch(dtc_learner* l, const dtc_storage* b, int64_t row0, int32_t M, const char* who) {
  if (!l) DTC_FAIL(DTC_ERR_ARG, "%s: null learner", who);
  RET_IF(check_storage(b, who));
  if (M <= 0 || M > l->R) DTC_FAIL(DTC_ERR_ARG, "%s: M=%d outside (0,%d]", who, M, l->R);
  if (row0 < 0 || row0 + M > (int64_t)b->T * b->N) DTC_FAIL(DTC_ERR_ARG, "%s: rows outside the batch", who);
  if (!l->grads) DTC_FAIL(DTC_ERR_STATE, "%s: learner has no gradient buffer", who);
  return DTC_OK;
}

extern "C" int dtc_vae_step(dtc_learner* l, const dtc_storage* batch, int64_t row0, int32_t M, const float* eps, uint64_t seed,
                            uint64_t counter, const dtc_ppo_hparams* hp, int sync_grads, void* stream) {
  DTC_NVTX("dtc_vae_step");
  RET_IF(check_batch(l, batch, row0, M, "dtc_vae_step"));
  if (!hp) DTC_FAIL(DTC_ERR_ARG, "dtc_vae_step: null hparams");
  cudaStream_t st = (cudaStream_t)stream;
  l->last_M = M;
  if (grads_prezeroed()) DTC_CUDA(cudaMemsetAsync(l->grads, 0, (size_t)g_vae_end * sizeof(float), st));  // [decoders | shared encoders]
  const float* hist = batch->hist + row0 * LD_HIST;
  const float* priv_a = batch->priv_a + row0 * LD_PRIVA;
  const float* xc = batch->xc + row0 * LD_XC;
  const float* next_obs = batch->next_obs + row0 * 56;
  l->ext[0] = {hist, (size_t)M * LD_HIST, batch->hist_lo ? batch->hist_lo + row0 * LD_HIST : nullptr};
  l->ext[1] = {priv_a, (size_t)M * LD_PRIVA, batch->priv_a_lo ? batch->priv_a_lo + row0 * LD_PRIVA : nullptr};
  l->ext[2] = {xc, (size_t)M * LD_XC, batch->xc_lo ? batch->xc_lo + row0 * LD_XC : nullptr};
  l->ext[0] = {hist, (size_t)M * LD_HIST, batch->hist_lo ? batch->hist_lo + row0 * LD_HIST : nullptr};
  l->ext[1] = {priv_a, (size_t)M * LD_PRIVA, batch->priv_a_lo ? batch->priv_a_lo + row0 * LD_PRIVA : nullptr};
  l->ext[2] = {xc, (size_t)M * LD_XC, batch->xc_lo ? batch->xc_lo + row0 * LD_XC : nullptr};
  const float inv_rows = 1.0f / (float)M;
  StepStreams S;
  RET_IF(streams_begin(l, st, &S));
  cudaStream_t sw = S.w, sc = S.c;
  // forward (ppo.py:197-223): CENet decoder on stream c, terrain decoder on the caller's stream
  RET_IF(encode(l, M, hist, priv_a, xc, 1, eps, seed, counter, S));
  RET_IF(chain(l, st, sc));
  RET_IF(fwd(l, CD0, l->XD, LD_XD, l->D1, 64, 1, M, sc));
  RET_IF(fwd(l, CD2, l->D1, 64, l->D2, 128, 1, M, sc));
  RET_IF(fwd(l, CD4, l->D2, 128, l->REC, 56, 0, M, sc, false));
  RET_IF(fwd(l, TD0, l->XD, LD_XD, l->U1, 512, 1, M, st));
  RET_IF(fwd(l, TD2, l->U1, 512, l->U2, 512, 1, M, st));
  RET_IF(fwd(l, TD4, l->U2, 512, l->HR, 696, 0, M, st, false));
  // losses and output gradients (ppo.py:213-247)
  k_vae_loss_rows<<<grid1d((long long)M * 14, 256, 148 * 2), 256, 0, sc>>>(M, inv_rows, l->REC, next_obs, l->ML, xc, l->dREC, l->dML, l->stats);
  DTC_CHECK_LAUNCH("k_vae_loss_rows");
  RET_IF(split_lo(l->dREC, lo_of(l, l->dREC), (int64_t)M * 56, sc));
  k_vae_loss_height<<<grid1d((long long)M * 174, 256), 256, 0, st>>>(M, 1.0f / ((float)M * 693.0f), l->HR, xc, l->dHR, lo_of(l, l->dHR),
                                                                     l->stats);
  DTC_CHECK_LAUNCH("k_vae_loss_height");
  // backward: cenet decoder (stream c), terrain decoder (caller's stream), weight gradients (stream w)
  RET_IF(chain(l, sc, sw));
  RET_IF(wgrad(l, CD4, l->dREC, 56, l->D2, 128, M, sw));
  RET_IF(chain(l, st, sw));
  RET_IF(wgrad(l, TD4, l->dHR, 696, l->U2, 512, M, sw));
  RET_IF(dgrad(l, CD4, l->dREC, 56, l->dD2, 128, 128, l->D2, 128, EPI_DRELU, false, M, sc, CD2, 1));
  RET_IF(chain(l, sc, sw));
  RET_IF(wgrad(l, CD2, l->dD2, 128, l->D1, 64, M, sw));
  RET_IF(dgrad(l, TD4, l->dHR, 696, l->dU2, 512, 512, l->U2, 512, EPI_DRELU, false, M, st, TD2, 0));
  RET_IF(chain(l, st, sw));
  RET_IF(wgrad(l, TD2, l->dU2, 512, l->U1, 512, M, sw));
  RET_IF(dgrad(l, CD2, l->dD2, 128, l->dD1, 64, 64, l->D1, 64, EPI_DRELU, false, M, sc, CD0, 1));
  RET_IF(chain(l, sc, sw));
  RET_IF(wgrad(l, CD0, l->dD1, 64, l->XD, LD_XD, M, sw));
  RET_IF(dgrad(l, TD2, l->dU2, 512, l->dU1, 512, 512, l->U1, 512, EPI_DRELU, false, M, st, TD0, 0));
  RET_IF(chain(l, st, sw));
  RET_IF(wgrad(l, TD0, l->dU1, 512, l->XD, LD_XD, M, sw));
  if (sync_grads) RET_IF(record_bucket(l, 0, 0, sw));  // every decoder gradient is complete on stream w
  RET_IF(dgrad(l, CD0, l->dD1, 64, l->dX, LD_XD, LD_XD, nullptr, 0, EPI_STORE, false, M, sc));
  // fan-in into d l_t: the terrain decoder's input gradient accumulates onto the CENet decoder's (dML comes from stream c too)
  RET_IF(chain(l, sc, st));
  RET_IF(dgrad(l, TD0, l->dU1, 512, l->dX, LD_XD, 512, nullptr, 0, EPI_STORE, true, M, st, TE4, 0));
  RET_IF(encode_bwd(l, M, hist, priv_a, 1, 1, S));
  RET_IF(streams_end(l, S));
  if (sync_grads) return record_bucket(l, 0, 1, st);
  return optimizer_apply(l, 0, hp, 1.0f, M, st);
}

extern "C" int dtc_ppo_step(dtc_learner* l, const dtc_storage* batch, int64_t row0, int32_t M, const float* eps, uint64_t seed,
                            uint64_t counter, const dtc_ppo_hparams* hp, int sync_grads, void* stream) {
  DTC_NVTX("dtc_ppo_step");
  RET_IF(check_batch(l, batch, row0, M, "dtc_ppo_step"));
  if (!hp) DTC_FAIL(DTC_ERR_ARG, "dtc_ppo_step: null hparams");
  cudaStream_t st = (cudaStream_t)stream;
  l->last_M = M;
  if (grads_prezeroed())  // [shared encoders | actor, critic, std | piggy-back scalars]
    DTC_CUDA(cudaMemsetAsync(l->grads + g_pol_begin, 0, (size_t)(g_off_piggy + NPIGGY - g_pol_begin) * sizeof(float), st));
  const float* hist = batch->hist + row0 * LD_HIST;
  const float* priv_a = batch->priv_a + row0 * LD_PRIVA;
  const float* xc = batch->xc + row0 * LD_XC;
  l->ext[0] = {hist, (size_t)M * LD_HIST, batch->hist_lo ? batch->hist_lo + row0 * LD_HIST : nullptr};
  l->ext[1] = {priv_a, (size_t)M * LD_PRIVA, batch->priv_a_lo ? batch->priv_a_lo + row0 * LD_PRIVA : nullptr};
  l->ext[2] = {xc, (size_t)M * LD_XC, batch->xc_lo ? batch->xc_lo + row0 * LD_XC : nullptr};
  const float inv_rows = 1.0f / (float)M;
  StepStreams S;
  RET_IF(streams_begin(l, st, &S));
  cudaStream_t sw = S.w, sc = S.c;
  // forward (ppo.py:265-292): the critic does not depend on the encoders and runs on stream w
  RET_IF(critic_fwd(l, M, xc, sw));
  RET_IF(encode(l, M, hist, priv_a, xc, 0, eps, seed, counter, S));
  RET_IF(actor_fwd(l, M, st));
  RET_IF(chain(l, sw, st));
  DTC_CUDA(cudaMemsetAsync(l->grads + g_off_std, 0, (12 + NPIGGY) * sizeof(float), st));
  k_ppo_loss<<<ceil_div(M, 128), 128, 0, st>>>(M, inv_rows, l->MEAN, l->V, l->params + g_off_std, batch->actions + row0 * 12,
                                              batch->values + row0, batch->advantages + row0, batch->returns + row0,
                                              batch->logp + row0, batch->mu + row0 * 12, batch->sigma + row0 * 12, *hp, l->dMEAN,
                                              l->dV, l->grads + g_off_std, l->stats);
  DTC_CHECK_LAUNCH("k_ppo_loss");
  k_publish_kl<<<1, 1, 0, st>>>(l->stats, l->grads + g_off_piggy);
  DTC_CHECK_LAUNCH("k_publish_kl");
  RET_IF(split_lo(l->dMEAN, lo_of(l, l->dMEAN), (int64_t)M * 12, st));
  RET_IF(split_lo(l->dV, lo_of(l, l->dV), (int64_t)M * 4, st));
  // backward: actor chain on the caller's stream, critic chain on stream c, weight gradients on stream w
  RET_IF(chain(l, st, sw));
  RET_IF(chain(l, st, sc));
  RET_IF(wgrad(l, AB6, l->dMEAN, 12, l->A3, 128, M, sw));
  RET_IF(wgrad(l, CB6, l->dV, 4, l->C3, 128, M, sw));
  RET_IF(dgrad(l, AB6, l->dMEAN, 12, l->dA3, 128, 128, l->A3, 128, EPI_DELU, false, M, st, AB4, 0));
  RET_IF(chain(l, st, sw));
  RET_IF(wgrad(l, AB4, l->dA3, 128, l->A2, 256, M, sw));
  RET_IF(dgrad(l, CB6, l->dV, 4, l->dC3, 128, 128, l->C3, 128, EPI_DELU, false, M, sc, CB4, 1));
  RET_IF(chain(l, sc, sw));
  RET_IF(wgrad(l, CB4, l->dC3, 128, l->C2, 256, M, sw));
  RET_IF(dgrad(l, AB4, l->dA3, 128, l->dA2, 256, 256, l->A2, 256, EPI_DELU, false, M, st, AB2, 0));
  RET_IF(chain(l, st, sw));
  RET_IF(wgrad(l, AB2, l->dA2, 256, l->A1, 512, M, sw));
  RET_IF(dgrad(l, CB4, l->dC3, 128, l->dC2, 256, 256, l->C2, 256, EPI_DELU, false, M, sc, CB2, 1));
  RET_IF(chain(l, sc, sw));
  RET_IF(wgrad(l, CB2, l->dC2, 256, l->C1, 512, M, sw));
  RET_IF(dgrad(l, AB2, l->dA2, 256, l->dA1, 512, 512, l->A1, 512, EPI_DELU, false, M, st, AB0, 0));
  RET_IF(chain(l, st, sw));
  RET_IF(wgrad(l, AB0, l->dA1, 512, l->XA, LD_XA, M, sw));
  RET_IF(dgrad(l, CB2, l->dC2, 256, l->dC1, 512, 512, l->C1, 512, EPI_DELU, false, M, sc, CB0, 1));
  RET_IF(chain(l, sc, sw));
  RET_IF(wgrad(l, CB0, l->dC1, 512, xc, LD_XC, M, sw));
  if (sync_grads) RET_IF(record_bucket(l, 1, 0, sw));  // actor, critic, std and the piggy-back scalars are complete on stream w
  RET_IF(dgrad(l, AB0, l->dA1, 512, l->dX, LD_XA, LD_XA, nullptr, 0, EPI_STORE, false, M, st, TE4, 0));
  RET_IF(encode_bwd(l, M, hist, priv_a, 0, 0, S));
  RET_IF(streams_end(l, S));
  if (sync_grads) return record_bucket(l, 1, 1, st);
  return optimizer_apply(l, 1, hp, 1.0f, M, st);
}

extern "C" int dtc_gemm_debug(int32_t M, int32_t N, int32_t K, const float* A, const float* A_lo, int32_t lda, int32_t a_kc, const float* B,
                              const float* B_lo, int32_t ldb, int32_t b_kc, float* C, float* C_lo, int32_t ldc, int32_t splits, float* ws,
                              int32_t mode, void* stream) {
  GemmArgs g{};
  g.A = A; g.A_lo = A_lo; g.lda = lda; g.a_kc = a_kc != 0;
  g.B = B; g.B_lo = B_lo; g.ldb = ldb; g.b_kc = b_kc != 0;
  g.C = C; g.C_lo = C_lo; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.epi = EPI_STORE; g.splits = splits; g.ws = ws;
  const int saved = dtc_gemm_mode();
  if (mode == 2) {  // tensor cores, operands' companions computed in shared memory (A_lo / B_lo ignored)
    g.A_lo = g.B_lo = nullptr; g.a_split = g.b_split = true;
    mode = 1;
  }
  dtc_gemm_set_mode(mode);
  int rc = dtc_gemm_launch(g, (cudaStream_t)stream);
  dtc_gemm_set_mode(saved);
  return rc;
}
