/* dtc_b200 - C ABI of the B200-native hot path of priest-yang/Deep-Tracking-Control.
 *
 * The reference has no FFI: its boundary is two duck-typed Python surfaces.  The
 * Python shims in deep-tracking-control_b200/{legged_gym,rsl_rl}/ keep those surfaces and bind the entry
 * points below through ctypes (INTEGRATION.md shows the stub).  Every function:
 *   - takes raw DEVICE pointers (tensor.data_ptr()), explicit sizes and a cudaStream_t (as void*);
 *   - returns 0 on success, a negative dtc_status otherwise (message: dtc_last_error());
 *   - never throws, never synchronises the stream, never allocates caller-visible memory
 *     (workspaces are sized by *_workspace_bytes and supplied by the caller).
 * Each entry point cites the reference code it replaces (paths relative to the reference root).
 */
#ifndef DTC_B200_H
#define DTC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum { DTC_OK = 0, DTC_ERR_ARG = -1, DTC_ERR_CUDA = -2, DTC_ERR_STATE = -3 } dtc_status;

const char* dtc_last_error(void);
int dtc_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t dtc_launch_count(void);
/* sizeof() of the ABI structs, for binding self-checks: 0 env_config, 1 env_buffers, 2 env_noise, 3 storage,
 * 4 ppo_hparams, 5 param_info */
int dtc_struct_size(int which);
/* bench.py roofline leg: CUDA events around every GEMM-family and foothold launch while enabled */
void dtc_profile_enable(int on);
int dtc_profile_read(double* gemm_flops, double* gemm_ms, int64_t* gemm_launches, double* foothold_ms, int64_t* foothold_launches);
/* per-kind totals of the last dtc_profile_read: 0 = GEMM family except the CTA-pair kernel, 1 = foothold kernel,
 * 2 = CTA-pair tcgen05 GEMM (work = 2*M*N*K fp32-equivalent flops) */
int dtc_profile_kind(int kind, double* work, double* ms, int64_t* launches);

/* ------------------------------------------------------------------ environment half */

/* Constant task description, resolved from the caller's configuration object the way LeggedRobot._parse_cfg /
 * _prepare_reward_function / _get_noise_scale_vec do (legged_robot.py:1230-1240, 929-952, 729-752); defaults: Lite3DTCCfg
 * (legged_gym/envs/lite3/lite3_dtc_config.py:3-181). */
typedef struct {
  int32_t num_envs;
  int32_t map_rows, map_cols;      /* height_samples [rows, cols] int16 (legged_gym/utils/terrain.py:26-30) */
  float horizontal_scale, vertical_scale, border_size;
  float dt;                        /* control dt = decimation * sim dt */
  int32_t max_episode_length;      /* ceil(20 s / dt) */
  int32_t resampling_steps;        /* int(10 s / dt) */
  int32_t push_interval;           /* ceil(15 s / dt) */
  float max_push_vel_xy;
  /* ranges as {lower, upper-lower}; the width is computed in double by the host, as `(upper - lower) * rand + lower`
   * does in torch_rand_float (legged_robot.py:573-578) */
  float cmd_lin_x[2], cmd_lin_y[2], cmd_heading[2];
  float motor_strength[2];
  float cmd_lin_x_max, cmd_ang_yaw_max; /* command_ranges[...][1] used by the soft tracking rewards */
  float base_height_target, tracking_sigma, max_acc;
  float terrain_length;            /* env_length of the curriculum rule */
  int32_t max_terrain_level;       /* num_rows */
  int32_t num_terrain_cols;
  float episode_length_s;
  float p_gains[12], d_gains[12];  /* control.stiffness / damping resolved per DOF name (legged_robot.py:1098-1109) */
  float action_scale, torque_limit;
  int32_t terrain_curriculum;      /* cfg.terrain.curriculum: move environments between terrain levels on reset (legged_robot.py:216-217) */
  int32_t push_robots;             /* cfg.domain_rand.push_robots (legged_robot.py:546) */
  float default_dof_pos[12];
  float dof_pos_lower[12], dof_pos_upper[12];
  float base_init_state[13];
  float grid_x[33], grid_y[21];    /* measured_points_x / _y as float32 */
  float plane_op[2 * 693];         /* rows 0,1 of (A^T A)^-1 A^T (legged_robot.py:1535-1547) */
  float reward_scale[24];          /* scale*dt in EPISODE_SUM_NAMES (alphabetical) order, termination at its slot */
  float noise_scale_vec[53];
  float obs_scale_lin_vel, obs_scale_ang_vel, obs_scale_dof_pos, obs_scale_dof_vel, obs_scale_height, obs_scale_force;
  float clip_obs, clip_actions;
} dtc_env_config;

/* Device buffers of one environment batch; all float32 unless noted.  Owned by the caller (torch tensors). */
typedef struct {
  /* simulator tensors (legged_robot.py:759-779) */
  float* root_states;        /* [N,13] */
  float* dof_state;          /* [N,12,2] */
  float* contact_forces;     /* [N,17,3] */
  float* rigid_body_state;   /* [N,17,13] */
  const int16_t* height_samples; /* [rows, cols] */
  /* per-step products */
  float* actions;            /* [N,12] clipped */
  float* torques;            /* [N,12] */
  float* lag_buffer;         /* [6,N,12] */
  float* base_lin_vel;       /* [N,3] */
  float* base_vel_scaled;    /* [N,3] base_lin_vel * obs_scales.lin_vel = get_base_vel() (legged_robot.py:1429-1431) */
  float* base_ang_vel;       /* [N,3] */
  float* projected_gravity;  /* [N,3] */
  float* commands;           /* [N,4] */
  float* cmd_buffer;         /* [10,N,4] */
  float* lin_vel_buffer;     /* [10,N,2] */
  float* ang_vel_buffer;     /* [10,N,1] */
  float* measured_heights;   /* [N,693] */
  float* pred_footholds;     /* [N,4,3] */
  int32_t* optimal_idx;      /* [N,4] */
  int32_t* nominal_idx;      /* [N,4] */
  float* foothold_obs;       /* [N,8] */
  float* optimal_footholds_world; /* [N,4,3] */
  float* center_clear_mean;  /* [N] mean(root_z - clip(heights[210:483], 0)) for check_termination */
  float* plane_ab;           /* [N,2] LS plane slopes for _reward_orientation */
  float* foot_clearance;     /* [N,4] */
  uint8_t* contact_filt;     /* [N,4] */
  uint8_t* last_contacts;    /* [N,4] */
  uint8_t* stumb_buffer;     /* [5,N,4] */
  float* feet_air_time;      /* [N,4] */
  float* pitch_est;          /* [N] */
  float* last_actions;       /* [N,12] */
  float* last_actions_2;     /* [N,12] */
  float* last_dof_vel;       /* [N,12] */
  float* last_root_vel;      /* [N,6] */
  float* last_foot_vel;      /* [N,4,3] */
  float* motor_strengths;    /* [N,12] */
  float* robot_mass;         /* [N] */
  float* height_noise_offset;/* [N] (reference stores it broadcast to [N,693]) */
  float* forces0;            /* [N,3] forces[:,0,:] */
  int64_t* episode_length_buf; /* [N] */
  int64_t* terrain_levels;   /* [N] */
  int64_t* terrain_types;    /* [N] */
  float* env_origins;        /* [N,3] */
  const float* terrain_origins; /* [rows, cols, 3] */
  uint8_t* reset_buf;        /* [N] */
  uint8_t* time_out_buf;     /* [N] */
  float* rew_buf;            /* [N] */
  float* episode_sums;       /* [24,N] */
  float* reward_terms;       /* [24,N] this step's scaled terms (debug / tests) */
  float* obs_buf;            /* [N,53] */
  float* privileged_obs_buf; /* [N, priv_ld] */
  float* obs_history;        /* [N, hist_ld] (HistoryWrapper state) */
  float* episode_stats;      /* [26]: per-key sum of episode_sums over envs reset this step (24), count, int32 bits: sum of terrain_levels */
  float* episode_stats_last; /* [26]: episode_stats of the last step that reset at least one environment - what the reference's
                                persistent `extras["episode"]` holds (legged_robot.py:253-262 runs only when len(env_ids) > 0) */
  uint8_t* time_outs_sent;   /* [N] extras["time_outs"]: time_out_buf as of the last step with a reset (legged_robot.py:263-264) */
  int32_t priv_ld, hist_ld;
} dtc_env_buffers;

/* Optional injected randomness (tests); any pointer may be NULL -> in-kernel Philox(seed, step, env). */
typedef struct {
  const float* resample_u;   /* [N,3] U[0,1): lin_vel_x, lin_vel_y, heading (legged_robot.py:573-576) */
  const float* push_u;       /* [N,2] (legged_robot.py:677) */
  const float* reset_u;      /* [N,25]: 0 curriculum randint, 1..12 dof, 13..14 xy, 15..20 vel, 21..23 cmd, 24 motor */
  const float* priv_u;       /* [N,693] rand_like(heights) (legged_robot_dtc.py:278) */
  const float* obs_u;        /* [N,53]  rand_like(obs_buf)  (legged_robot_dtc.py:287) */
} dtc_env_noise;

typedef struct dtc_env dtc_env;
int dtc_env_create(const dtc_env_config* cfg, dtc_env** out);
void dtc_env_destroy(dtc_env* e);
/* binds the caller's buffers; also tabulates the foothold kernel's single-tap min3 map from buf->height_samples
 * (library-owned, synchronous: the heightmap must already hold the terrain, as after LeggedRobot._create_heightfield,
 * legged_robot.py:1201-1228) */
int dtc_env_bind(dtc_env* e, const dtc_env_buffers* buf);
/* call after writing height_samples in place (the reference never does after start-up) */
int dtc_env_heightmap_updated(dtc_env* e);
/* CUDA-graph support: launches captured into a graph carry baked arguments, but every environment launch needs the running
 * common_step_counter (command resampling, pushes, Philox streams).  With a device-side counter set here, each environment kernel
 * adds its value to the `common_step_counter` / `step` argument of the call, so a captured rollout passes the step RELATIVE to the
 * counter and advances the counter with dtc_counter_add at the end of the graph.  NULL (default) = arguments are absolute.
 * dtc_count_launches adds kernels launched by graph replays to dtc_launch_count(). */
int dtc_env_set_step_base(dtc_env* e, const int64_t* device_counter);
int dtc_counter_add(int64_t* device_counter, int64_t inc, void* stream);
void dtc_count_launches(int64_t n);

/* E1+E2: clip actions (first_substep == 0), PD torque of decimation sub-steps [first_substep, first_substep + num_substeps) with the
 * lag buffer (legged_robot.py:92-111,595-630).  The reference recomputes the torque from the refreshed dof state in every
 * sub-step and hands it to gym.set_dof_actuation_force_tensor, so a real simulator is driven with four calls (s, 1) around
 * gym.simulate(); a stub that leaves dof_state alone inside the loop may take all four in one call (0, 4) - same arithmetic.
 * lag_choice[4]: np.random.randint(1,5) per sub-step (:608), ONE value for all environments: 1..4 = the caller's draw (tests,
 * seeded numpy parity), 0 = drawn inside the kernel from Philox(seed, step) so the step needs no host randomness (CUDA graphs);
 * entries outside the range of this call are ignored.  step: the common_step_counter this step will have in post_physics_step. */
int dtc_env_pre_physics(dtc_env* e, const float* actions_in, const int32_t lag_choice[4], int32_t first_substep,
                        int32_t num_substeps, int64_t step, uint64_t seed, void* stream);

/* E3+E4(commands part): base velocities, history buffers, command resampling, heading command
 * (legged_robot_dtc.py:66-91, legged_robot.py:534-539,567-593). */
int dtc_env_state_prep(dtc_env* e, int64_t common_step_counter, uint64_t seed, const dtc_env_noise* noise, void* stream);

/* E5 + E7..E10: height sampling, Raibert footholds, terrain score, argmin, decode
 * (legged_robot.py:1279-1317; legged_robot_dtc.py:100-201).  THE foothold-scoring kernel.
 * variant 0 = brute force with L2 gathers, 3 = patch staged with bulk row copies, 4 = lazy window scoring, 5 = persistent warps +
 * min3 map + prefetched patch, 6 = CTA-batched per-environment work + packed fp32x2 sampling (default); 1 and 2 are rejected
 * (removed).  debug_score may be NULL or [N,693,4] (then the brute-force kernel writes it first and the chosen variant follows). */
int dtc_foothold_step(dtc_env* e, int variant, float* debug_score, void* stream);

/* E4(rest), E6, E11, E12, E13: push, foot clearance, contact filter, termination, 23 rewards, reset
 * (legged_robot.py:546-564,1443-1472,274-291,200-272; legged_robot_dtc.py:229-245,522-586).
 * reset_normal: np.random.normal(0,0.02) of legged_robot.py:230 (one value per step for all environments reset in it): the
 * caller's draw, or NaN = drawn inside the kernel from Philox(seed, step). */
int dtc_env_reward_reset(dtc_env* e, int64_t common_step_counter, uint64_t seed, float reset_normal,
                         const dtc_env_noise* noise, void* stream);

/* E14+E15+E1(clip): observations, privileged observations, noise, clip, history shift
 * (legged_robot_dtc.py:254-287; legged_robot.py:118-121; history_wrapper.py:23) and the last_* roll (:215-219). */
int dtc_env_observe(dtc_env* e, int64_t common_step_counter, uint64_t seed, const dtc_env_noise* noise, void* stream);

/* ------------------------------------------------------------------ learner half */

/* Parameter table.  All parameters of ActorCriticDecoder (rsl_rl/modules/actor_critic_decoder.py:91-369) live in ONE
 * flat float32 buffer.  A weight is stored [rows, ld] (ld = in_features rounded up to 4, pad columns zero) with its
 * input columns grouped into 16-byte aligned segments so that concatenated first-layer inputs need no copy:
 * reference column seg_src[i]+j <-> internal column seg_dst[i]+j, j < seg_len[i].  Biases / std: rows = 1. */
typedef struct {
  char name[48];
  int64_t offset;
  int32_t rows, cols, ld;
  int32_t nseg, seg_src[4], seg_dst[4], seg_len[4];
} dtc_param_info;
int dtc_param_count(void);
int dtc_param_get(int i, dtc_param_info* out);
int64_t dtc_param_total_floats(void);
/* [begin,end) float ranges: 0 = VAE optimizer step (ppo.py:249-254), 1 = policy optimizer step (ppo.py:332-335),
 * 2 = policy step range plus the piggy-back scalars that ride in the same gradient all-reduce (slot 0: sum of KL),
 * 3 = everything */
int dtc_param_range(int which, int64_t* begin, int64_t* end);

/* Packed rollout storage (RolloutStorage, rsl_rl/storage/rollout_storage.py:36-116): R = T*N rows, time-major.
 * Rows are laid out GEMM-ready:  hist [R,268] = obs_history 265 + 3 zero;  priv_a [R,696] = priv[:, :693] + 3 zero;
 * xc [R,752] = priv[:, 693:1389] | obs 53 | base_vel 3  (the critic input of actor_critic_decoder.py:540-551 up to a
 * column permutation);  next_obs [R,56];  actions/mu/sigma [R,12];  rewards/values/returns/advantages/logp [R];
 * dones [R] uint8. */
typedef struct {
  float *hist, *priv_a, *xc, *next_obs, *actions, *mu, *sigma;
  float *rewards, *values, *returns, *advantages, *logp;
  uint8_t* dones;
  /* optional 3xTF32 companions of the three GEMM-input arrays (x - trunc_tf32(x), same shapes); NULL = those GEMMs run on
   * the FP32 SIMT path.  Written by dtc_policy_act / dtc_gather_minibatch. */
  float *hist_lo, *priv_a_lo, *xc_lo;
  int32_t T, N;
} dtc_storage;

typedef struct dtc_learner dtc_learner;
int64_t dtc_learner_workspace_bytes(int32_t max_rows);
/* `stream`: the stream the caller wrote `params` on; the workspace is cleared and the 3xTF32 parameter companions are derived on
 * it (asynchronously), so the first step queued on the same stream sees both */
int dtc_learner_create(int32_t max_rows, float* params, float* grads, float* adam_main_m, float* adam_main_v,
                       float* adam_vae_m, float* adam_vae_v, void* workspace, int64_t workspace_bytes, void* stream,
                       dtc_learner** out);
void dtc_learner_destroy(dtc_learner* l);
/* call after writing the flat parameter buffer from outside (load_state_dict): refreshes the 3xTF32 companions */
int dtc_learner_refresh_params(dtc_learner* l, void* stream);

/* P6: PPO.act = ActorCriticDecoder.act + evaluate + log_prob (ppo.py:137-155; actor_critic_decoder.py:409-451,540-551).
 * eps_z [M,16] / eps_a [M,12]: standard normal draws, or NULL for in-kernel Philox(seed, counter).
 * If `s` is not NULL the packed inputs and the outputs are written straight into rows [step*N, step*N+M) of the
 * storage (the act-time half of add_transitions, rollout_storage.py:99-116) and M must equal s->N.
 * actions/values/logp/mean/sigma: optional extra outputs ([M,12] / [M]). */
int dtc_policy_act(dtc_learner* l, int32_t M, const float* obs, int32_t obs_ld, const float* hist, int32_t hist_ld,
                   const float* priv, int32_t priv_ld, const float* base_vel, int32_t bv_ld,
                   const float* eps_z, const float* eps_a, uint64_t seed, uint64_t counter,
                   const dtc_storage* s, int32_t step,
                   float* actions, float* values, float* logp, float* mean, float* sigma, void* stream);
/* CUDA-graph support, as dtc_env_set_step_base: a device-side counter added to dtc_policy_act's `counter` argument (NULL = absolute). */
int dtc_learner_set_act_counter_base(dtc_learner* l, const uint64_t* device_counter);
/* P4 alone (ppo.py:170-171) */
int dtc_policy_evaluate(dtc_learner* l, int32_t M, const float* obs, int32_t obs_ld, const float* priv, int32_t priv_ld,
                        const float* base_vel, int32_t bv_ld, float* values, void* stream);
/* deployment path act_teacher (actor_critic_decoder.py:504-538): mean actions [M,12] */
int dtc_policy_act_teacher(dtc_learner* l, int32_t M, const float* obs, int32_t obs_ld, const float* hist, int32_t hist_ld,
                           const float* priv, int32_t priv_ld, float* actions, void* stream);

/* P7+P8, env-time half: rewards (+ gamma * values * time_outs, ppo.py:157-168), dones, next_obs into row block `step`. */
int dtc_store_transition(const dtc_storage* s, int32_t step, const float* rewards, const uint8_t* dones,
                         const uint8_t* time_outs, const float* next_obs, int32_t next_obs_ld, float gamma, void* stream);

/* P9: GAE reverse scan + advantage normalisation (rollout_storage.py:138-152).  scratch: >= 4 doubles.
 * With defer_normalize != 0 the {sum, sumsq, count} triple is left in scratch[0..2] (for a cross-rank all-reduce)
 * and dtc_gae_normalize finishes the job. */
int dtc_gae(const dtc_storage* s, const float* last_values, float gamma, float lam, double* scratch, int defer_normalize, void* stream);
int dtc_gae_normalize(const dtc_storage* s, const double* stats3, void* stream);

/* P10: minibatch row gather by permutation (rollout_storage.py:165-214): dst row r = src row perm[r]. */
int dtc_gather_minibatch(const dtc_storage* src, const dtc_storage* dst, const int64_t* perm, int64_t rows, void* stream);

/* P11 / P12: one VAE optimizer step / one policy optimizer step on rows [row0,row0+M) of a (gathered) storage
 * (ppo.py:197-254 / :265-338): forward + hand-written backward + clip_grad_norm_ + Adam.
 * eps: [M,16] normal draws or NULL (Philox).  Loss sums accumulate into the learner's device-side statistics.
 * If sync_grads != 0 the call stops after backward so the caller can all-reduce grads[dtc_param_range(0 or 2)] and
 * then call dtc_optimizer_apply with grad_scale = 1/world. */
typedef struct {
  float clip_param, value_loss_coef, entropy_coef, max_grad_norm, desired_kl;
  int32_t adaptive_lr, use_clipped_value_loss, reserved;
} dtc_ppo_hparams;
int dtc_vae_step(dtc_learner* l, const dtc_storage* batch, int64_t row0, int32_t M, const float* eps, uint64_t seed,
                 uint64_t counter, const dtc_ppo_hparams* hp, int sync_grads, void* stream);
int dtc_ppo_step(dtc_learner* l, const dtc_storage* batch, int64_t row0, int32_t M, const float* eps, uint64_t seed,
                 uint64_t counter, const dtc_ppo_hparams* hp, int sync_grads, void* stream);
int dtc_optimizer_apply(dtc_learner* l, int which /*0 vae, 1 policy*/, const dtc_ppo_hparams* hp, float grad_scale,
                        int32_t rows_global /* minibatch rows over all ranks (KL mean) */, void* stream);

/* Data parallel: with sync_grads != 0 a step publishes its flat gradient range in TWO contiguous buckets, each with
 * an event, in the order backward completes them - bucket 0 (VAE step: the decoders; policy step: actor, critic, std and the
 * piggy-back scalars) while the shared encoders' backward is still running, bucket 1 (the shared encoders) at the end of the call.
 * dtc_learner_grad_bucket gives a bucket's [begin, end) floats; dtc_learner_wait_bucket makes `stream` (the caller's
 * communication stream) wait for it, so the all-reduce of bucket 0 runs underneath the rest of the backward pass. */
int dtc_learner_grad_bucket(int which /*0 vae, 1 policy*/, int bucket, int64_t* begin, int64_t* end);
int dtc_learner_wait_bucket(dtc_learner* l, int which, int bucket, void* stream);

/* learner statistics (device doubles): [0] value_loss sum, [1] surrogate sum, [2] recons, [3] vel, [4] kld,
 * [5] height, [6] entropy sum, [7] last kl_mean, [8] learning_rate, [9] last grad norm (vae), [10] last grad norm
 * (policy), [11] number of vae steps accumulated, [12] number of policy steps accumulated */
double* dtc_learner_stats(dtc_learner* l);
int dtc_learner_set_lr(dtc_learner* l, double lr, void* stream);
int dtc_learner_reset_stats(dtc_learner* l, void* stream);
int dtc_learner_set_adam_steps(dtc_learner* l, int64_t vae_steps, int64_t main_steps);
int dtc_learner_get_adam_steps(dtc_learner* l, int64_t* vae_steps, int64_t* main_steps);
/* debug access to named activation / gradient buffers of the last step (tests) */
int dtc_learner_debug_buffer(dtc_learner* l, const char* name, float** ptr, int32_t* rows, int32_t* cols, int32_t* ld);

/* plain GEMM entry (tests / microbench): C[M,N] = act(A[M,K] * W[N,K]^T + bias) */
int dtc_linear_forward(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, const float* W, int32_t ldw,
                       const float* bias, int32_t act /*0 none,1 relu,2 elu*/, float* C, int32_t ldc, void* stream);
/* general form used by the learner: C = A op B with either operand k-contiguous (1) or k-strided (0); tests and
 * microbenchmarks.  mode 0 = FP32 SIMT, 1 = tcgen05 3xTF32 (A_lo / B_lo = companions x - trunc_tf32(x), may be NULL),
 * 2 = tcgen05 3xTF32 with both companions computed tile by tile in shared memory (A_lo / B_lo ignored): the learner's default. */
int dtc_gemm_debug(int32_t M, int32_t N, int32_t K, const float* A, const float* A_lo, int32_t lda, int32_t a_kc, const float* B,
                   const float* B_lo, int32_t ldb, int32_t b_kc, float* C, float* C_lo, int32_t ldc, int32_t splits, float* ws,
                   int32_t mode, void* stream);
/* GEMM engine of the learner: 0 = FP32 SIMT, 1 = tcgen05 3xTF32 for tile-worthy shapes (default; env DTC_GEMM=simt|tc) */
void dtc_set_gemm_mode(int mode);
int dtc_get_gemm_mode(void);
/* tensor-core engine only: 1 (default; env DTC_GEMM_PAIR) = shapes with enough 256-row tiles run on CTA pairs
 * (tcgen05 cta_group::2, the two SMs of a TPC share each B tile), 0 = always the single-CTA kernel.  Same results. */
void dtc_set_gemm_pair(int on);
int dtc_get_gemm_pair(void);
/* Intra-step concurrency of dtc_policy_act / dtc_vae_step / dtc_ppo_step (default on): the CENet chains and the weight
 * gradients are forked onto two library-owned non-blocking streams with events and joined back into the caller's stream
 * before the call returns; results are identical either way.  Off while dtc_profile_enable(1) is active. */
void dtc_set_overlap(int on);
int dtc_get_overlap(void);

/* ------------------------------------------------------------------ terrain rasterisation on the device
 * legged_gym/utils/terrain.py:9-243 lays num_rows x num_cols sub-terrains (pyramid stairs, discrete obstacles, stepping
 * stones: isaacgym.terrain_utils) inside a flat border and uploads the int16 heightfield.  Here the host draws only the random
 * parameters and the device evaluates every cell.  type: 0 flat; 1 stepping stones (a = stone side px, b = gap px, c = pit depth
 * in height units, table = first-stone offset of every stone column); 2 pyramid stairs (a = step width px, b = signed step
 * height in units, c = platform px); 3 discrete obstacles (a = count <= 20, table = {x, y, w, l, height} per rectangle).
 * platform_half > 0 flattens the central square of that half-width.  tables: int32 [n_rows * n_cols][256] (device).
 * terrain_origins [n_rows, n_cols, 3] (may be NULL) = sub-terrain centres with the height of their central 20 x 20 cells. */
typedef struct { int32_t type, a, b, c, platform_half; } dtc_subterrain;
int dtc_terrain_rasterize(int32_t rows, int32_t cols, int32_t border_px, int32_t sub_px, int32_t n_rows, int32_t n_cols,
                          const dtc_subterrain* subs, const int32_t* tables, double terrain_length, double vertical_scale,
                          int16_t* height_samples, float* terrain_origins, void* stream);

/* The reference's own generator, legged_gym/utils/terrain.py:9-243 (Terrain.curiculum / randomized_terrain / make_terrain /
 * add_terrain_to_map and the isaacgym.terrain_utils + gap / pit / stones_everywhere generators behind it): each sub-terrain is a
 * background height plus an ORDERED list of `height_field_raw[x0:x1, y0:y1] = h` assignments, which the host records while replaying
 * the generator's loops and numpy draws (deep-tracking-control_b200/legged_gym/utils/terrain.py); the device paints them - the last
 * rectangle covering a cell wins.  subs[s] = {type 5, a = number of rectangles, b = index of the first in `rects`, c = background};
 * rects: int32 [n][5] = {x0, x1, y0, y1, height} in sub-terrain cells, half-open.  terrain_origins [n_rows, n_cols, 3] = sub-terrain
 * centre with the maximum height over origin_window = {x1, x2, y1, y2} (terrain.py:152-160). */
int dtc_terrain_paint(int32_t rows, int32_t cols, int32_t border_px, int32_t len_px, int32_t wid_px, int32_t n_rows, int32_t n_cols,
                      const dtc_subterrain* subs, const int32_t* rects, const int32_t origin_window[4], double terrain_length,
                      double terrain_width, double vertical_scale, int16_t* height_samples, float* terrain_origins, void* stream);

/* ------------------------------------------------------------------ C1: gradient all-reduce over NVLink peer memory
 * (the reference is single-process and has no collective - this replaces the torch.distributed / NCCL all-reduce
 * a data-parallel port of rsl_rl/algorithms/ppo.py:254,338 would call between backward and optimizer.step()).
 * One process per GPU.  dtc_dp_create allocates this rank's exchange buffer and flag words with cudaMalloc; dtc_dp_handles returns
 * their two 64-byte cudaIpcMemHandle_t, which the host side passes to the other ranks by any channel (torch.distributed
 * all_gather_object in rsl_rl/utils/dp.py) and every rank maps with dtc_dp_open.  dtc_dp_allreduce(data, n) is then three launches on
 * `stream` - publish, reduce-scatter + all-gather through P2P loads / stores, collect - with device-side flag handshakes only: no host
 * synchronisation, no NCCL.  Every element is summed by one rank in rank order: all replicas get bit-identical sums.  n % 4 == 0,
 * data 16-byte aligned, n <= max_floats; all ranks must issue the same sequence of calls.  dtc_dp_error synchronises `stream` and
 * reports whether any wait ever exceeded ~20 s (then the data is not a sum and the run should stop). */
typedef struct dtc_dp dtc_dp;
int dtc_dp_create(int32_t rank, int32_t world, int64_t max_floats, dtc_dp** out);
int dtc_dp_handles(dtc_dp* d, void* buf_handle64, void* flags_handle64);
int dtc_dp_open(dtc_dp* d, int32_t peer, const void* buf_handle64, const void* flags_handle64);
 /* Registered in-place variant (what training uses): dtc_dp_register(base, nfloats) returns the IPC handle of the cudaMalloc allocation
 * that contains this rank's own buffer (e.g. the flat gradient buffer inside a block of torch's caching allocator) and base's byte
 * offset in it; after every rank has mapped every other rank's range with dtc_dp_open_registered, dtc_dp_allreduce on a sub-range of
 * the registered buffer is ONE cooperative kernel that sums straight out of / into all ranks' buffers (no staging copies). */
int dtc_dp_register(dtc_dp* d, float* base, int64_t nfloats, void* handle64, int64_t* offset_bytes);
int dtc_dp_open_registered(dtc_dp* d, int32_t peer, const void* handle64, int64_t offset_bytes);
int dtc_dp_allreduce(dtc_dp* d, float* data, int64_t n, void* stream);
int dtc_dp_error(dtc_dp* d, void* stream);
void dtc_dp_destroy(dtc_dp* d);

/* ------------------------------------------------------------------ P14 / SURVEY 8f N2: the optional GRU `Memory`
 * Memory.forward / Memory.reset (rsl_rl/rsl_rl/modules/actor_critic_decoder.py:584-614): nn.GRU(input_size, hidden_size,
 * num_layers) over T time steps for N rows.  `weights`: per layer weight_ih_l [3H,in_l] | weight_hh_l [3H,H] | bias_ih_l [3H] |
 * bias_hh_l [3H] back to back (nn.GRU's parameter and gate order r|z|n; dtc_gru_param_floats gives the total).
 * x [T,N,input_size]; h [num_layers,N,H] initial hidden state in, final hidden state out; out [T,N,H] top-layer outputs
 * (may be NULL for a single layer).  The reference never instantiates Memory on its training path. */
int64_t dtc_gru_param_floats(int32_t input_size, int32_t hidden_size, int32_t num_layers);
int dtc_gru_forward(int32_t T, int32_t N, int32_t input_size, int32_t hidden_size, int32_t num_layers, const float* weights,
                    const float* x, float* h, float* out, void* stream);
/* hidden_state[..., dones, :] = 0 (Memory.reset, actor_critic_decoder.py:609-613) */
int dtc_gru_reset(int32_t N, int32_t hidden_size, int32_t num_layers, float* h, const uint8_t* dones, void* stream);

#ifdef __cplusplus
}
#endif
#endif
