// Environment half of the hot path as sm_100a kernels.
// One launch sequence per policy step:  pre_physics -> [simulator stub] -> state_prep -> foothold -> reward_reset -> observe
#include <stdlib.h>
#include "dtc_common.cuh"
#include "dtc_env_internal.cuh"

// ------------------------------------------------------------------ RNG slots (one Philox stream per env per step)
enum { SLOT_RESAMPLE = 0, SLOT_PUSH = 1, SLOT_RESET = 2, SLOT_LAG = 10, SLOT_RESET_NORMAL = 11, SLOT_PRIV = 16, SLOT_OBS = 400 };
// draws the reference makes ONCE per step for all environments (np.random.randint lag choice, np.random.normal reset offset):
// env index 0xffffffff keeps them apart from every per-environment stream
__device__ __forceinline__ Philox step_rng(uint64_t seed, int64_t step, int slot) {
  return Philox(seed, (uint64_t)step, ((uint64_t)0xffffffffu << 32) | ((uint64_t)slot << 8));
}
__device__ __forceinline__ Philox env_rng(uint64_t seed, int64_t step, int env, int slot) {
  return Philox(seed, (uint64_t)step, ((uint64_t)(uint32_t)env << 32) | ((uint64_t)slot << 8));
}

// ================================================================== E1 + E2
// legged_robot.py:92-111 (clip, decimation loop) and :595-630 (_compute_torques): sub-steps [first, first + count) of the
// decimation loop.  A real simulator changes dof_pos / dof_vel between sub-steps, so the host launches one sub-step at a time
// around gym.simulate(); when the simulator is a stub that leaves the dof state alone inside the loop the four sub-steps run
// back to back in one launch (count = 4) - the same arithmetic either way.
__global__ void __launch_bounds__(256) k_pre_physics(const dtc_env_config* __restrict__ cfg, dtc_env_buffers b,
                                                     const float* __restrict__ actions_in, int4 choice, int first, int count,
                                                     int64_t step, uint64_t seed, const int64_t* __restrict__ step_base) {
  if (step_base) step += *step_base;  // CUDA-graph replays: the launch carries the step relative to a device-side counter
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int N = cfg->num_envs;
  if (i >= N * 12) return;
  int j = i % 12;
  float a;
  if (first == 0) {
    a = fminf(fmaxf(actions_in[i], -cfg->clip_actions), cfg->clip_actions);
    b.actions[i] = a;
  } else {
    a = b.actions[i];
  }
  float scaled = a * cfg->action_scale;
  float lag[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) lag[k] = b.lag_buffer[(size_t)k * N * 12 + i];
  float q = b.dof_state[2 * i], qd = b.dof_state[2 * i + 1];
  float ms = b.motor_strengths[i];
  int ch[4] = {choice.x, choice.y, choice.z, choice.w};
  if (ch[0] == 0 || ch[1] == 0 || ch[2] == 0 || ch[3] == 0) {
    // np.random.randint(1, 5) of legged_robot.py:608, one value per sub-step shared by all environments, drawn on the device
    const uint4 r = step_rng(seed, step, SLOT_LAG).next();
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int s = 0; s < 4; ++s)
      if (ch[s] == 0) ch[s] = 1 + (int)(w[s] >> 30);
  }
  float tq = 0.f;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (s < first || s >= first + count) continue;
#pragma unroll
    for (int k = 0; k < 5; ++k) lag[k] = lag[k + 1];
    lag[5] = scaled;
    float sel = lag[1];
#pragma unroll
    for (int k = 2; k < 5; ++k) sel = (ch[s] == k) ? lag[k] : sel;
    float goal = fminf(fmaxf(sel + cfg->default_dof_pos[j], cfg->dof_pos_lower[j]), cfg->dof_pos_upper[j]);
    // p_gains*Kp(=1) * (goal - q + offsets(=0)) - d_gains*Kd(=1) * qd
    tq = __fsub_rn(__fmul_rn(cfg->p_gains[j], __fsub_rn(goal, q)), __fmul_rn(cfg->d_gains[j], qd));
    tq = __fmul_rn(tq, ms);
    tq = fminf(fmaxf(tq, -cfg->torque_limit), cfg->torque_limit);
  }
  b.torques[i] = tq;
#pragma unroll
  for (int k = 0; k < 6; ++k) b.lag_buffer[(size_t)k * N * 12 + i] = lag[k];
}

// ================================================================== E3 + E4 (command part)
__device__ __forceinline__ float wrap_to_pi_dev(float x) {
  const float two_pi = 6.28318530717958647692f, pi = 3.14159265358979323846f;
  float m = fmodf(x, two_pi);
  if (m != 0.f && m < 0.f) m += two_pi;  // torch.remainder (sign of divisor)
  if (m > pi) m = __fsub_rn(m, two_pi);
  return m;
}

__device__ __forceinline__ void resample_commands_dev(const dtc_env_config* cfg, dtc_env_buffers& b, int n, float u0, float u1, float u2) {
  float cx = __fadd_rn(__fmul_rn(cfg->cmd_lin_x[1], u0), cfg->cmd_lin_x[0]);
  float cy = __fadd_rn(__fmul_rn(cfg->cmd_lin_y[1], u1), cfg->cmd_lin_y[0]);
  float ch = __fadd_rn(__fmul_rn(cfg->cmd_heading[1], u2), cfg->cmd_heading[0]);
  float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)));
  float keep = nrm > 0.1f ? 1.f : 0.f;
  b.commands[n * 4 + 0] = cx * keep;
  b.commands[n * 4 + 1] = cy * keep;
  b.commands[n * 4 + 3] = ch;
  b.forces0[n * 3 + 0] = 0.f; b.forces0[n * 3 + 1] = 0.f; b.forces0[n * 3 + 2] = 0.f;
}

__global__ void __launch_bounds__(128) k_state_prep(const dtc_env_config* __restrict__ cfg, dtc_env_buffers b,
                                                    int64_t step, uint64_t seed, dtc_env_noise nz, const int64_t* __restrict__ step_base) {
  if (step_base) step += *step_base;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int N = cfg->num_envs;
  if (n >= N) return;
  if (n == 0) {
#pragma unroll
    for (int k = 0; k < 26; ++k) b.episode_stats[k] = 0.f;  // refilled by k_reward_reset's atomics this step
  }
  const float* rs = b.root_states + (size_t)n * 13;
  float q[4] = {rs[3], rs[4], rs[5], rs[6]};
  float lv[3] = {rs[7], rs[8], rs[9]}, av[3] = {rs[10], rs[11], rs[12]}, g[3] = {0.f, 0.f, -1.f};
  float blv[3], bav[3], pg[3];
  quat_rotate_inverse_exact(q, lv, blv);
  quat_rotate_inverse_exact(q, av, bav);
  quat_rotate_inverse_exact(q, g, pg);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    b.base_lin_vel[n * 3 + i] = blv[i];
    b.base_vel_scaled[n * 3 + i] = __fmul_rn(blv[i], cfg->obs_scale_lin_vel);  // get_base_vel() (legged_robot.py:1429-1431)
    b.base_ang_vel[n * 3 + i] = bav[i];
    b.projected_gravity[n * 3 + i] = pg[i];
  }
  int64_t el = b.episode_length_buf[n] + 1;
  b.episode_length_buf[n] = el;
  // roll the 10-deep history buffers (legged_robot_dtc.py:76-81); cmd_buffer takes the PRE-resample command
  for (int k = 0; k < 9; ++k) {
    size_t d = (size_t)k * N + n, s = (size_t)(k + 1) * N + n;
    b.lin_vel_buffer[d * 2] = b.lin_vel_buffer[s * 2];
    b.lin_vel_buffer[d * 2 + 1] = b.lin_vel_buffer[s * 2 + 1];
    b.ang_vel_buffer[d] = b.ang_vel_buffer[s];
#pragma unroll
    for (int c = 0; c < 4; ++c) b.cmd_buffer[d * 4 + c] = b.cmd_buffer[s * 4 + c];
  }
  {
    size_t d = (size_t)9 * N + n;
    b.lin_vel_buffer[d * 2] = blv[0];
    b.lin_vel_buffer[d * 2 + 1] = blv[1];
    b.ang_vel_buffer[d] = bav[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) b.cmd_buffer[d * 4 + c] = b.commands[n * 4 + c];
  }
  // _resample_commands at the 10 s boundary (legged_robot.py:534-535)
  if (el % cfg->resampling_steps == 0) {
    float u0, u1, u2;
    if (nz.resample_u) { u0 = nz.resample_u[n * 3]; u1 = nz.resample_u[n * 3 + 1]; u2 = nz.resample_u[n * 3 + 2]; }
    else { Philox p = env_rng(seed, step, n, SLOT_RESAMPLE); uint4 r = p.next(); u0 = u01(r.x); u1 = u01(r.y); u2 = u01(r.z); }
    resample_commands_dev(cfg, b, n, u0, u1, u2);
  }
  // heading -> yaw-rate command (legged_robot.py:536-539)
  float fwd[3], e1[3] = {1.f, 0.f, 0.f};
  quat_apply_exact(q, e1, fwd);
  float heading = atan2f(fwd[1], fwd[0]);
  float w = wrap_to_pi_dev(__fsub_rn(b.commands[n * 4 + 3], heading));
  b.commands[n * 4 + 2] = fminf(fmaxf(__fmul_rn(0.5f, w), -1.5f), 1.5f);
}


// ================================================================== E4(rest) + E6 + E11 + E12 + E13
// One thread per environment: ~2 KB of state in, 24 reward terms, termination and the in-place episode reset.
// Reward terms are evaluated in the reference's (alphabetical) order and summed term by term in float so the
// accumulation order of legged_robot.py:279-291 is kept.
__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

__global__ void __launch_bounds__(128) k_reward_reset(const dtc_env_config* __restrict__ cfg, dtc_env_buffers b,
                                                      int64_t step, uint64_t seed, float reset_normal, dtc_env_noise nz,
                                                      const int64_t* __restrict__ step_base) {
  if (step_base) step += *step_base;
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = cfg->num_envs;
  if (n >= N) return;
  const float dt = cfg->dt;
  float* rs = b.root_states + (size_t)n * 13;
  const float* rb = b.rigid_body_state + (size_t)n * 17 * 13;
  const float* cf = b.contact_forces + (size_t)n * 17 * 3;
  const int16_t* __restrict__ hs = b.height_samples;
  const int rows = cfg->map_rows, cols = cfg->map_cols;

  // ---- push (legged_robot.py:546-556): counter % interval in {0,1} -> random xy base velocity
  {
    int64_t m = step % cfg->push_interval;
    if (cfg->push_robots && (m == 0 || m == 1)) {
      float u0, u1;
      if (nz.push_u) { u0 = nz.push_u[n * 2]; u1 = nz.push_u[n * 2 + 1]; }
      else { Philox p = env_rng(seed, step, n, SLOT_PUSH); uint4 r = p.next(); u0 = u01(r.x); u1 = u01(r.y); }
      float w = 2.0f * cfg->max_push_vel_xy;
      rs[7] = __fadd_rn(__fmul_rn(w, u0), -cfg->max_push_vel_xy);
      rs[8] = __fadd_rn(__fmul_rn(w, u1), -cfg->max_push_vel_xy);
      if (m == 0) { b.forces0[n * 3] = 0.f; b.forces0[n * 3 + 1] = 0.f; }  // max_push_force_xy = 0
    } else {
      b.forces0[n * 3] = 0.f; b.forces0[n * 3 + 1] = 0.f; b.forces0[n * 3 + 2] = 0.f;
    }
  }
  // ---- feet: positions, velocities, clearance (E6, legged_robot.py:1443-1472), contacts (:562-564)
  float fpos[4][3], fvel[4][3], clr[4], fF[4][3];
  bool contact[4], cfilt[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const float* body = rb + (4 + 4 * f) * 13;
#pragma unroll
    for (int i = 0; i < 3; ++i) { fpos[f][i] = body[i]; fvel[f][i] = body[7 + i]; fF[f][i] = cf[(4 + 4 * f) * 3 + i]; }
    int px = (int)__fdiv_rn(__fadd_rn(fpos[f][0], cfg->border_size), cfg->horizontal_scale);
    int py = (int)__fdiv_rn(__fadd_rn(fpos[f][1], cfg->border_size), cfg->horizontal_scale);
    px = min(max(px, 1), rows - 3);
    py = min(max(py, 1), cols - 3);
    const int16_t* g = hs + (size_t)px * cols + py;
    // px - 2 / py - 2 reach -1 for a foot at the low map edge (px, py are clipped to >= 1): torch indexing wraps a negative
    // index to the far edge (height_samples[-1, py] is the LAST row), legged_robot.py:1456-1470
    const int16_t* gxm2 = hs + (size_t)(px - 2 < 0 ? px - 2 + rows : px - 2) * cols + py;
    const int16_t* gym2 = hs + (size_t)px * cols + (py - 2 < 0 ? py - 2 + cols : py - 2);
    int h = __ldg(g);
    h = max(h, (int)__ldg(g + cols)); h = max(h, (int)__ldg(g + 1)); h = max(h, (int)__ldg(g + 2 * cols));
    h = max(h, (int)__ldg(g + 2)); h = max(h, (int)__ldg(g + cols + 1)); h = max(h, (int)__ldg(g - cols));
    h = max(h, (int)__ldg(g - 1)); h = max(h, (int)__ldg(gxm2)); h = max(h, (int)__ldg(gym2));
    clr[f] = __fsub_rn(fpos[f][2], __fmul_rn((float)h, cfg->vertical_scale));
    b.foot_clearance[n * 4 + f] = clr[f];
    contact[f] = fF[f][2] > 1.0f;
    cfilt[f] = contact[f] || (b.last_contacts[n * 4 + f] != 0);
    b.contact_filt[n * 4 + f] = cfilt[f];
    b.last_contacts[n * 4 + f] = contact[f];
  }
  // ---- E11 check_termination (legged_robot_dtc.py:229-245; termination-contact set is empty for Lite3 DTC)
  const int64_t el = b.episode_length_buf[n];
  const bool time_out = el > (int64_t)cfg->max_episode_length;
  const float pgx = b.projected_gravity[n * 3], pgz = b.projected_gravity[n * 3 + 2];
  bool reset = time_out || (pgz > 0.2f) || (b.center_clear_mean[n] < 0.15f);
  b.time_out_buf[n] = time_out;

  // ---- E12: 23 reward terms (alphabetical) + termination
  float q[12], qd[12], tq[12], act[12], la[12], la2[12], ldv[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    q[j] = b.dof_state[(n * 12 + j) * 2];
    qd[j] = b.dof_state[(n * 12 + j) * 2 + 1];
    tq[j] = b.torques[n * 12 + j];
    act[j] = b.actions[n * 12 + j];
    la[j] = b.last_actions[n * 12 + j];
    la2[j] = b.last_actions_2[n * 12 + j];
    ldv[j] = b.last_dof_vel[n * 12 + j];
  }
  const float blv[3] = {b.base_lin_vel[n * 3], b.base_lin_vel[n * 3 + 1], b.base_lin_vel[n * 3 + 2]};
  const float bav[3] = {b.base_ang_vel[n * 3], b.base_ang_vel[n * 3 + 1], b.base_ang_vel[n * 3 + 2]};
  const float cmdx = b.commands[n * 4], cmdy = b.commands[n * 4 + 1];
  const float cmd_norm = sqrtf(cmdx * cmdx + cmdy * cmdy);
  float term[24];
  // 0 action_rate (:1620)
  { float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) { float d = la[j] - act[j]; s += d * d; } term[0] = s; }
  // 1 ang_vel_xy (:1325)
  term[1] = bav[0] * bav[0] + bav[1] * bav[1];
  // 2 base_height (dtc:531)
  { float mz = (fpos[0][2] + fpos[1][2] + fpos[2][2] + fpos[3][2]) / 4.0f; float d = (rs[2] - mz) - cfg->base_height_target; term[2] = d * d; }
  // 3 collision (:1350): TORSO, THIGH x4, SHANK x4
  { const int pen[9] = {0, 2, 6, 10, 14, 3, 7, 11, 15}; float s = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) s += norm3(cf[pen[k] * 3], cf[pen[k] * 3 + 1], cf[pen[k] * 3 + 2]) > 0.1f ? 1.f : 0.f;
    term[3] = s; }
  // 4 dof_acc (:1342)
  { float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) { float d = (ldv[j] - qd[j]) / dt; s += d * d; } term[4] = s; }
  // 5 dof_pos_limits (:1358)
  { float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) s += -fminf(q[j] - cfg->dof_pos_lower[j], 0.f) + fmaxf(q[j] - cfg->dof_pos_upper[j], 0.f);
    term[5] = s; }
  // 6 feet_air_time (:1386) - stateful.  last_contacts already equals `contact` here (callback ran first), so
  //   the function's local contact_filt is just `contact`.
  { float s = 0.f;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      float air = b.feet_air_time[n * 4 + f];
      bool first = (air > 0.f) && contact[f];
      air += dt;
      s += (air - 0.5f) * (first ? 1.f : 0.f);
      air *= contact[f] ? 0.f : 1.f;
      b.feet_air_time[n * 4 + f] = air;
    }
    term[6] = s * (cmd_norm > 0.1f ? 1.f : 0.f); }
  // 7 feet_slip (:1494)
  { float s = 0.f;
#pragma unroll
    for (int f = 0; f < 4; ++f) { float v = sqrtf(fvel[f][0] * fvel[f][0] + fvel[f][1] * fvel[f][1]); s += (contact[f] ? 1.f : 0.f) * (v * v); }
    term[7] = s; }
  // 8 foot_acc (:1525)
  { float mask = b.terrain_levels[n] > 5 ? 0.2f : 1.0f, s = 0.f;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      float ax = (b.last_foot_vel[(n * 4 + f) * 3] - fvel[f][0]) / dt, ay = (b.last_foot_vel[(n * 4 + f) * 3 + 1] - fvel[f][1]) / dt,
            az = (b.last_foot_vel[(n * 4 + f) * 3 + 2] - fvel[f][2]) / dt;
      s += fmaxf(mask * (norm3(ax, ay, az) - cfg->max_acc), 0.f);
    }
    term[8] = s; }
  // 9 foot_clearance (:1474) - stateful 5-deep stumble buffer
  { float s = 0.f;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      bool stumb = sqrtf(fF[f][0] * fF[f][0] + fF[f][1] * fF[f][1]) > 4.0f * fabsf(fF[f][2]);
      bool flag = stumb;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint8_t v = b.stumb_buffer[((size_t)(k + 1) * N + n) * 4 + f];
        b.stumb_buffer[((size_t)k * N + n) * 4 + f] = v;
        flag = flag || (v != 0);
      }
      b.stumb_buffer[((size_t)4 * N + n) * 4 + f] = stumb;
      s += (!flag && clr[f] > 0.18f) ? 1.f : 0.f;
    }
    term[9] = s; }
  // 10 foothold_miss (dtc:536)
  term[10] = fminf(fminf(fpos[0][2], fpos[1][2]), fminf(fpos[2][2], fpos[3][2])) < 0.f ? 1.f : 0.f;
  // 11 hip_pos (:1504)
  term[11] = q[0] * q[0] + q[3] * q[3] + q[6] * q[6] + q[9] * q[9];
  // 12 lin_vel_z (:1321)
  term[12] = blv[2] * blv[2];
  // 13 orientation (:1559, get_plane_norm :1535) - stateful pitch_est
  { float ax = b.plane_ab[n * 2], by = b.plane_ab[n * 2 + 1];
    float inv = sqrtf(ax * ax + by * by + 1.0f);
    float pnx = -(ax / inv), pny = -(by / inv);
    float pitch = atanf(pnx), roll = -atanf(pny);
    float pitch_c = (pitch >= -0.1f && pitch <= 0.1f) ? 0.f : pitch;
    float roll_c = (roll >= -0.1f && roll <= 0.1f) ? 0.f : roll;
    float pe = b.pitch_est[n] * 0.2f + 0.8f * pitch_c;
    b.pitch_est[n] = pe;
    float cr = cosf(roll_c * 0.5f), sr = sinf(roll_c * 0.5f), cp = cosf(pe * 0.5f), sp = sinf(pe * 0.5f);
    float qq[4] = {sr * cp, cr * sp, -(sr * sp), cr * cp};  // quat_from_euler_xyz(roll, pitch, 0): cy=1, sy=0
    float g[3] = {0.f, 0.f, -1.f}, loc[3];
    quat_rotate_inverse_exact(qq, g, loc);
    float d = pgx - loc[0];
    term[13] = d * d; }
  // 14 pos_acc (:1600): 8 box corners (+-0.15, +-0.1, +-0.075)
  { float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float px = (k & 4) ? 0.15f : -0.15f, py = (k & 2) ? 0.1f : -0.1f, pz = (k & 1) ? 0.075f : -0.075f;
      float vx = blv[0] + (bav[1] * pz - bav[2] * py), vy = blv[1] + (bav[2] * px - bav[0] * pz), vz = blv[2] + (bav[0] * py - bav[1] * px);
      float nn = norm3(vx, vy, vz);
      s += nn * nn;
    }
    term[14] = s; }
  // 15 power (:1435), 16 powerchange (:1613)
  { float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) s += fmaxf(tq[j] * qd[j], 0.f);
    term[15] = s;
    float co = fmaxf(cmdx, 1.0f);
    float r = s / (b.robot_mass[n] * 9.815f * co);
    term[16] = r * r; }
  // 17 smooth (:1440)
  { float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) { float d = act[j] - 2.0f * la[j] + la2[j]; s += d * d; } term[17] = s; }
  // 18 soft_tracking_ang_vel (dtc:555; lookback 4, tolerance 0.15), 19 soft_tracking_lin_vel (dtc:542; lookback 3)
  { float s = 0.f;
#pragma unroll
    for (int k = 6; k < 10; ++k) {
      float d = (b.cmd_buffer[((size_t)k * N + n) * 4 + 2] - b.ang_vel_buffer[(size_t)k * N + n]) / cfg->cmd_ang_yaw_max;
      d = d * d;
      d = d <= 0.0225f ? 0.f : 1.f;
      s += expf(-d / cfg->tracking_sigma);
    }
    term[18] = s / 4.0f;
    float lx = b.lin_vel_buffer[((size_t)7 * N + n) * 2], ly = b.lin_vel_buffer[((size_t)7 * N + n) * 2 + 1];
    float range = cfg->cmd_lin_x_max;
    s = 0.f;
#pragma unroll
    for (int k = 7; k < 10; ++k) {
      float dx = (b.cmd_buffer[((size_t)k * N + n) * 4] - lx) / range, dy = (b.cmd_buffer[((size_t)k * N + n) * 4 + 1] - ly) / range;
      s += expf(-(dx * dx + dy * dy) / cfg->tracking_sigma);
    }
    term[19] = s / 3.0f; }
  // 20 stand_still (:1422)
  { float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) s += fabsf(q[j] - cfg->default_dof_pos[j]);
    term[20] = s * (cmd_norm < 0.1f ? 1.f : 0.f); }
  // 21 termination (:1354) - applied after the loop
  term[21] = (reset && !time_out) ? 1.f : 0.f;
  // 22 torques (:1334)
  { float s = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) s += tq[j] * tq[j];
    term[22] = s; }
  // 23 tracking_optimal_footholds (dtc:577) - uses the callback's contact_filt
  { float s = 0.f;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      float dx = fpos[f][0] - b.optimal_footholds_world[(n * 4 + f) * 3], dy = fpos[f][1] - b.optimal_footholds_world[(n * 4 + f) * 3 + 1];
      float r = -logf(0.8f + sqrtf(dx * dx + dy * dy));
      s += cfilt[f] ? r : 0.f;
    }
    term[23] = s; }
  float rew = 0.f;
#pragma unroll
  for (int k = 0; k < 24; ++k) {
    if (k == 21) continue;
    float r = term[k] * cfg->reward_scale[k];
    rew += r;
    b.episode_sums[(size_t)k * N + n] += r;
    b.reward_terms[(size_t)k * N + n] = r;
  }
  { float r = term[21] * cfg->reward_scale[21];
    rew += r;
    b.episode_sums[(size_t)21 * N + n] += r;
    b.reward_terms[(size_t)21 * N + n] = r; }
  b.rew_buf[n] = rew;
  b.reset_buf[n] = reset;
  if (!reset) {
    atomicAdd(reinterpret_cast<int*>(b.episode_stats) + 25, (int)b.terrain_levels[n]);  // mean(terrain_levels) of extras["episode"]
    return;
  }

  // ---- E13 reset_idx (legged_robot.py:200-272), per environment
  float u[25];
  if (nz.reset_u) {
#pragma unroll
    for (int k = 0; k < 25; ++k) u[k] = nz.reset_u[n * 25 + k];
  } else {
    Philox p = env_rng(seed, step, n, SLOT_RESET);
#pragma unroll
    for (int k = 0; k < 28; k += 4) {
      uint4 r = p.next();
      if (k < 25) u[k] = u01(r.x);
      if (k + 1 < 25) u[k + 1] = u01(r.y);
      if (k + 2 < 25) u[k + 2] = u01(r.z);
      if (k + 3 < 25) u[k + 3] = u01(r.w);
    }
  }
  // terrain curriculum (:690-711), only with cfg.terrain.curriculum (:216-217)
  if (!cfg->terrain_curriculum) {
    atomicAdd(reinterpret_cast<int*>(b.episode_stats) + 25, (int)b.terrain_levels[n]);
  } else {
    float dx = rs[0] - b.env_origins[n * 3], dy = rs[1] - b.env_origins[n * 3 + 1];
    float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    bool up = dist > cfg->terrain_length * 0.6f;
    float cn = __fsqrt_rn(__fadd_rn(__fmul_rn(b.commands[n * 4], b.commands[n * 4]), __fmul_rn(b.commands[n * 4 + 1], b.commands[n * 4 + 1])));
    bool down = (dist < __fmul_rn(__fmul_rn(cn, cfg->episode_length_s), 0.5f)) && !up;
    int64_t lv = b.terrain_levels[n] + (up ? 1 : 0) - (down ? 1 : 0);
    if (lv >= cfg->max_terrain_level) {
      int r = (int)(u[0] * (float)cfg->max_terrain_level);
      lv = min(r, cfg->max_terrain_level - 1);
    } else if (lv < 0) lv = 0;
    b.terrain_levels[n] = lv;
    atomicAdd(reinterpret_cast<int*>(b.episode_stats) + 25, (int)lv);
    const float* to = b.terrain_origins + ((size_t)lv * cfg->num_terrain_cols + b.terrain_types[n]) * 3;
    b.env_origins[n * 3] = to[0]; b.env_origins[n * 3 + 1] = to[1]; b.env_origins[n * 3 + 2] = to[2];
  }
  // _reset_dofs (:640), _reset_root_states (dtc:299-311)
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    b.dof_state[(n * 12 + j) * 2] = __fmul_rn(cfg->default_dof_pos[j], __fadd_rn(__fmul_rn(1.0f, u[1 + j]), 0.5f));
    b.dof_state[(n * 12 + j) * 2 + 1] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < 13; ++k) rs[k] = cfg->base_init_state[k];
  rs[0] = __fadd_rn(rs[0], b.env_origins[n * 3]); rs[1] = __fadd_rn(rs[1], b.env_origins[n * 3 + 1]); rs[2] = __fadd_rn(rs[2], b.env_origins[n * 3 + 2]);
  rs[0] = __fadd_rn(rs[0], __fadd_rn(__fmul_rn(1.0f, u[13]), -0.5f));
  rs[1] = __fadd_rn(rs[1], __fadd_rn(__fmul_rn(1.0f, u[14]), -0.5f));
#pragma unroll
  for (int k = 0; k < 6; ++k) rs[7 + k] = __fadd_rn(__fmul_rn(1.0f, u[15 + k]), -0.5f);
  resample_commands_dev(cfg, b, n, u[21], u[22], u[23]);
  { float ms = __fadd_rn(__fmul_rn(u[24], cfg->motor_strength[1]), cfg->motor_strength[0]);
#pragma unroll
    for (int j = 0; j < 12; ++j) b.motor_strengths[n * 12 + j] = ms; }
  if (reset_normal != reset_normal) {
    // np.random.normal(0, 0.02) of legged_robot.py:230: one value per reset_idx() call, shared by the environments it resets
    const uint4 r = step_rng(seed, step, SLOT_RESET_NORMAL).next();
    reset_normal = 0.02f * box_muller(r.x, r.y).x;
  }
  b.height_noise_offset[n] = __fadd_rn(__fmul_rn(b.height_noise_offset[n], 0.0f), reset_normal);
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    b.last_actions[n * 12 + j] = 0.f; b.last_actions_2[n * 12 + j] = 0.f; b.last_dof_vel[n * 12 + j] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) b.lag_buffer[(size_t)k * N * 12 + n * 12 + j] = 0.f;
  }
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    b.feet_air_time[n * 4 + f] = 0.f; b.contact_filt[n * 4 + f] = 0; b.last_contacts[n * 4 + f] = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) b.stumb_buffer[((size_t)k * N + n) * 4 + f] = 0;
  }
  b.episode_length_buf[n] = 0;
  b.pitch_est[n] = 0.f;
  for (int k = 0; k < 24; ++k) {
    atomicAdd(&b.episode_stats[k], b.episode_sums[(size_t)k * N + n]);
    b.episode_sums[(size_t)k * N + n] = 0.f;
  }
  atomicAdd(&b.episode_stats[24], 1.0f);
  for (int k = 0; k < 10; ++k) {
    size_t d = (size_t)k * N + n;
    b.lin_vel_buffer[d * 2] = 0.f; b.lin_vel_buffer[d * 2 + 1] = 0.f; b.ang_vel_buffer[d] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) b.cmd_buffer[d * 4 + c] = 0.f;
  }
}

// ================================================================== E14 + E15 + clip + last_* roll
// One warp per environment: 53 obs + 1389 privileged + 265 history floats are written as coalesced rows.
__global__ void __launch_bounds__(128) k_observe(const dtc_env_config* __restrict__ cfg, dtc_env_buffers b, int64_t step,
                                                 uint64_t seed, dtc_env_noise nz, const int64_t* __restrict__ step_base) {
  if (step_base) step += *step_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = cfg->num_envs;
  const int n = blockIdx.x * 4 + warp;
  if (n >= N) return;
  const float clipv = cfg->clip_obs;
  const float root_z = b.root_states[(size_t)n * 13 + 2];
  // the reference's `extras` dict persists between steps and is rewritten only by a reset_idx() call with a non-empty id list
  // (legged_robot.py:210,253-264): publish this step's episode statistics / time-out flags only if some environment was reset
  if (b.episode_stats[24] > 0.f) {
    if (lane == 0) b.time_outs_sent[n] = b.time_out_buf[n];
    if (n == 0 && lane < 26) b.episode_stats_last[lane] = b.episode_stats[lane];
  }
  // ---- obs (legged_robot_dtc.py:259-272, :287) + history shift (history_wrapper.py:23)
  float* hist = b.obs_history + (size_t)n * b.hist_ld;
  float* obs = b.obs_buf + (size_t)n * 53;
  // shift first (reads 53..264, writes 0..211): stage through registers, 7 per lane
  float hreg[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { int c = lane + 32 * k; hreg[k] = c < 212 ? hist[53 + c] : 0.f; }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 7; ++k) { int c = lane + 32 * k; if (c < 212) hist[c] = hreg[k]; }
  for (int c = lane; c < 53; c += 32) {
    float v;
    if (c < 3) v = b.base_ang_vel[n * 3 + c] * cfg->obs_scale_ang_vel;
    else if (c < 6) v = b.projected_gravity[n * 3 + c - 3];
    else if (c < 9) v = b.commands[n * 4 + c - 6] * (c < 8 ? cfg->obs_scale_lin_vel : cfg->obs_scale_ang_vel);
    else if (c < 21) v = (b.dof_state[(n * 12 + c - 9) * 2] - cfg->default_dof_pos[c - 9]) * cfg->obs_scale_dof_pos;
    else if (c < 33) v = b.dof_state[(n * 12 + c - 21) * 2 + 1] * cfg->obs_scale_dof_vel;
    else if (c < 45) v = b.actions[n * 12 + c - 33];
    else v = b.foothold_obs[n * 8 + c - 45];
    float u;
    if (nz.obs_u) u = nz.obs_u[(size_t)n * 53 + c];
    else { Philox p = env_rng(seed, step, n, SLOT_OBS + c); u = u01(p.next().x); }
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(2.0f, u), 1.0f), cfg->noise_scale_vec[c]));
    v = fminf(fmaxf(v, -clipv), clipv);
    obs[c] = v;
    hist[212 + c] = v;
  }
  // ---- privileged obs (legged_robot_dtc.py:274-281)
  float* priv = b.privileged_obs_buf + (size_t)n * b.priv_ld;
  const float* mh = b.measured_heights + (size_t)n * NP;
  const float hno = b.height_noise_offset[n];
  const float base = __fsub_rn(root_z, cfg->base_height_target);
  for (int c = lane; c < NP; c += 32) {
    float h = __fmul_rn(fminf(fmaxf(__fsub_rn(base, mh[c]), -1.0f), 1.0f), cfg->obs_scale_height);
    float u;
    if (nz.priv_u) u = nz.priv_u[(size_t)n * NP + c];
    else {
      Philox p = env_rng(seed, step, n, SLOT_PRIV + (c >> 2));
      uint4 r = p.next();
      uint32_t w = (c & 3) == 0 ? r.x : (c & 3) == 1 ? r.y : (c & 3) == 2 ? r.z : r.w;
      u = u01(w);
    }
    float noisy = __fadd_rn(__fadd_rn(h, __fmul_rn(__fsub_rn(__fmul_rn(2.0f, u), 1.0f), 0.1f)), hno);
    priv[c] = fminf(fmaxf(noisy, -clipv), clipv);
    priv[NP + 3 + c] = fminf(fmaxf(h, -clipv), clipv);
  }
  if (lane < 3) priv[NP + lane] = fminf(fmaxf(b.forces0[n * 3 + lane] * cfg->obs_scale_force, -clipv), clipv);
  // ---- last_* roll (legged_robot_dtc.py:215-219)
  if (lane < 12) {
    int i = n * 12 + lane;
    b.last_actions_2[i] = b.last_actions[i];
    b.last_actions[i] = b.actions[i];
    b.last_dof_vel[i] = b.dof_state[i * 2 + 1];
    int f = lane / 3, c = lane - f * 3;
    b.last_foot_vel[i] = b.rigid_body_state[((size_t)n * 17 + 4 + 4 * f) * 13 + 7 + c];
  } else if (lane < 18) {
    b.last_root_vel[n * 6 + lane - 12] = b.root_states[(size_t)n * 13 + 7 + lane - 12];
  }
}

// ================================================================== C ABI
thread_local char g_dtc_err[512] = "";
int64_t g_dtc_launches = 0;

extern "C" const char* dtc_last_error(void) { return g_dtc_err; }
extern "C" int dtc_version(void) { return 100; }
extern "C" int64_t dtc_launch_count(void) { return g_dtc_launches; }

extern "C" int dtc_env_create(const dtc_env_config* cfg, dtc_env** out) {
  if (!cfg || !out) DTC_FAIL(DTC_ERR_ARG, "dtc_env_create: null argument");
  if (cfg->num_envs <= 0) DTC_FAIL(DTC_ERR_ARG, "dtc_env_create: num_envs must be positive");
  dtc_env* e = new dtc_env();
  e->cfg = *cfg;
  e->bound = false;
  e->min3 = nullptr;
  e->gtab = nullptr;
  e->min3_bytes = 0;
  e->min3_map_ok = false;
  e->step_base = nullptr;
  cudaError_t ce = cudaMalloc(&e->d_cfg, sizeof(dtc_env_config));
  if (ce == cudaSuccess) ce = cudaMemcpy(e->d_cfg, cfg, sizeof(dtc_env_config), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) { delete e; DTC_FAIL(DTC_ERR_CUDA, "dtc_env_create: %s", cudaGetErrorString(ce)); }
  *out = e;
  return DTC_OK;
}
extern "C" void dtc_env_destroy(dtc_env* e) {
  if (!e) return;
  cudaFree(e->d_cfg);
  if (e->min3) cudaFree(e->min3);
  if (e->gtab) cudaFree(e->gtab);
  delete e;
}
extern "C" int dtc_env_bind(dtc_env* e, const dtc_env_buffers* buf) {
  if (!e || !buf) DTC_FAIL(DTC_ERR_ARG, "dtc_env_bind: null argument");
  const void* const* p = (const void* const*)buf;
  size_t nptr = offsetof(dtc_env_buffers, priv_ld) / sizeof(void*);
  for (size_t i = 0; i < nptr; ++i)
    if (!p[i]) DTC_FAIL(DTC_ERR_ARG, "dtc_env_bind: buffer #%zu is null", i);
  if (buf->priv_ld < 2 * NP + 3 || buf->hist_ld < 265) DTC_FAIL(DTC_ERR_ARG, "dtc_env_bind: row strides too small");
  e->buf = *buf;
  e->bound = true;
  return dtc_env_build_min3(e);
}
extern "C" int dtc_env_set_step_base(dtc_env* e, const int64_t* device_counter) {
  if (!e) DTC_FAIL(DTC_ERR_ARG, "dtc_env_set_step_base: null env");
  e->step_base = device_counter;
  return DTC_OK;
}
__global__ void k_counter_add(int64_t* p, int64_t inc) { *p += inc; }
extern "C" int dtc_counter_add(int64_t* device_counter, int64_t inc, void* stream) {
  if (!device_counter) DTC_FAIL(DTC_ERR_ARG, "dtc_counter_add: null pointer");
  k_counter_add<<<1, 1, 0, (cudaStream_t)stream>>>(device_counter, inc);
  DTC_CHECK_LAUNCH("k_counter_add");
  return DTC_OK;
}
extern "C" void dtc_count_launches(int64_t n) { g_dtc_launches += n; }
extern "C" int dtc_env_heightmap_updated(dtc_env* e) {
  if (!e || !e->bound) DTC_FAIL(DTC_ERR_STATE, "dtc_env_heightmap_updated: env not bound");
  return dtc_env_build_min3(e);
}
extern "C" int dtc_env_pre_physics(dtc_env* e, const float* actions_in, const int32_t lag_choice[4], int32_t first_substep,
                                   int32_t num_substeps, int64_t step, uint64_t seed, void* stream) {
  DTC_NVTX("dtc_env_pre_physics");
  if (!e || !e->bound) DTC_FAIL(DTC_ERR_STATE, "dtc_env_pre_physics: env not bound");
  if (first_substep < 0 || num_substeps < 1 || first_substep + num_substeps > 4)
    DTC_FAIL(DTC_ERR_ARG, "dtc_env_pre_physics: sub-steps [%d, %d) outside the decimation loop [0, 4)", first_substep, first_substep + num_substeps);
  for (int i = first_substep; i < first_substep + num_substeps; ++i)
    if (lag_choice[i] < 0 || lag_choice[i] > 4) DTC_FAIL(DTC_ERR_ARG, "lag choice must be in 1..4 (or 0: drawn on the device)");
  if (first_substep == 0 && !actions_in) DTC_FAIL(DTC_ERR_ARG, "dtc_env_pre_physics: actions_in is null");
  int total = e->cfg.num_envs * 12;
  k_pre_physics<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      e->d_cfg, e->buf, actions_in, make_int4(lag_choice[0], lag_choice[1], lag_choice[2], lag_choice[3]), first_substep, num_substeps,
      step, seed, e->step_base);
  DTC_CHECK_LAUNCH("k_pre_physics");
  return DTC_OK;
}
static dtc_env_noise noise_or_null(const dtc_env_noise* nz) {
  dtc_env_noise z;
  memset(&z, 0, sizeof(z));
  return nz ? *nz : z;
}
extern "C" int dtc_env_state_prep(dtc_env* e, int64_t step, uint64_t seed, const dtc_env_noise* noise, void* stream) {
  DTC_NVTX("dtc_env_state_prep");
  if (!e || !e->bound) DTC_FAIL(DTC_ERR_STATE, "dtc_env_state_prep: env not bound");
  k_state_prep<<<ceil_div(e->cfg.num_envs, 128), 128, 0, (cudaStream_t)stream>>>(e->d_cfg, e->buf, step, seed, noise_or_null(noise), e->step_base);
  DTC_CHECK_LAUNCH("k_state_prep");
  return DTC_OK;
}
extern "C" int dtc_env_reward_reset(dtc_env* e, int64_t step, uint64_t seed, float reset_normal, const dtc_env_noise* noise, void* stream) {
  DTC_NVTX("dtc_env_reward_reset");
  if (!e || !e->bound) DTC_FAIL(DTC_ERR_STATE, "dtc_env_reward_reset: env not bound");
  k_reward_reset<<<ceil_div(e->cfg.num_envs, 128), 128, 0, (cudaStream_t)stream>>>(e->d_cfg, e->buf, step, seed, reset_normal,
                                                                                 noise_or_null(noise), e->step_base);
  DTC_CHECK_LAUNCH("k_reward_reset");
  return DTC_OK;
}
extern "C" int dtc_env_observe(dtc_env* e, int64_t step, uint64_t seed, const dtc_env_noise* noise, void* stream) {
  DTC_NVTX("dtc_env_observe");
  if (!e || !e->bound) DTC_FAIL(DTC_ERR_STATE, "dtc_env_observe: env not bound");
  k_observe<<<ceil_div(e->cfg.num_envs, 4), 128, 0, (cudaStream_t)stream>>>(e->d_cfg, e->buf, step, seed, noise_or_null(noise), e->step_base);
  DTC_CHECK_LAUNCH("k_observe");
  return DTC_OK;
}

// ------------------------------------------------------------------ per-launch event timing
#include <vector>
int g_dtc_prof = 0;
struct ProfRec { cudaEvent_t a, b; int kind; double work; int m, n, k, layout; };
static std::vector<ProfRec> g_prof_recs;
void dtc_prof_begin(cudaStream_t st, int kind, double work) {
  if (!g_dtc_prof) return;
  ProfRec r;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  r.kind = kind; r.work = work;
  r.m = r.n = r.k = r.layout = 0;
  cudaEventRecord(r.a, st);
  g_prof_recs.push_back(r);
}
void dtc_prof_end(cudaStream_t st) {
  if (!g_dtc_prof || g_prof_recs.empty()) return;
  cudaEventRecord(g_prof_recs.back().b, st);
}
void dtc_prof_tag(int m, int n, int k, int layout) {
  if (!g_dtc_prof || g_prof_recs.empty()) return;
  ProfRec& r = g_prof_recs.back();
  r.m = m; r.n = n; r.k = k; r.layout = layout;
}
extern "C" void dtc_profile_enable(int on) { g_dtc_prof = on; }
static double g_prof_last_work[4], g_prof_last_ms[4];
static int64_t g_prof_last_n[4];
// kinds: 0 = GEMM family except the CTA-pair kernel, 1 = foothold kernel, 2 = CTA-pair tensor-core GEMM (the dominant kernel)
extern "C" int dtc_profile_read(double* gemm_flops, double* gemm_ms, int64_t* gemm_launches, double* foothold_ms, int64_t* foothold_launches) {
  for (int k = 0; k < 4; ++k) { g_prof_last_work[k] = g_prof_last_ms[k] = 0.0; g_prof_last_n[k] = 0; }
  // env DTC_PROF_DUMP=<file>: one line per launch (kind m n k layout ms), for per-shape tables (tools/gemm_shapes.py)
  FILE* dump = nullptr;
  if (const char* path = getenv("DTC_PROF_DUMP")) dump = fopen(path, "a");
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    cudaEventSynchronize(r.b);
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (dump) fprintf(dump, "%d %d %d %d %d %.6f\n", r.kind, r.m, r.n, r.k, r.layout, ms);
    const int k = r.kind >= 0 && r.kind < 4 ? r.kind : 3;
    g_prof_last_work[k] += r.work; g_prof_last_ms[k] += ms; ++g_prof_last_n[k];
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof_recs.clear();
  if (dump) fclose(dump);
  if (gemm_flops) *gemm_flops = g_prof_last_work[0] + g_prof_last_work[2];
  if (gemm_ms) *gemm_ms = g_prof_last_ms[0] + g_prof_last_ms[2];
  if (gemm_launches) *gemm_launches = g_prof_last_n[0] + g_prof_last_n[2];
  if (foothold_ms) *foothold_ms = g_prof_last_ms[1];
  if (foothold_launches) *foothold_launches = g_prof_last_n[1];
  return DTC_OK;
}
extern "C" int dtc_profile_kind(int kind, double* work, double* ms, int64_t* launches) {
  if (kind < 0 || kind >= 4) DTC_FAIL(DTC_ERR_ARG, "dtc_profile_kind: unknown kind %d", kind);
  if (work) *work = g_prof_last_work[kind];
  if (ms) *ms = g_prof_last_ms[kind];
  if (launches) *launches = g_prof_last_n[kind];
  return DTC_OK;
}

extern "C" int dtc_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(dtc_env_config);
    case 1: return (int)sizeof(dtc_env_buffers);
    case 2: return (int)sizeof(dtc_env_noise);
    case 3: return (int)sizeof(dtc_storage);
    case 4: return (int)sizeof(dtc_ppo_hparams);
    case 5: return (int)sizeof(dtc_param_info);
    default: return -1;
  }
}
