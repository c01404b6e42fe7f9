"""TEST INFRASTRUCTURE ONLY - CPU restatement (oracle) of the environment half of the hot path.

Restates, function by function, what `LeggedRobotDTC.step()` computes outside the physics call
(reference: legged_gym/envs/base/legged_robot_dtc.py:56-288,522-586 and
legged_gym/envs/base/legged_robot.py:92-122,200-291,529-630,690-711,1279-1317,1321-1622), plus
`HistoryWrapper` (rsl_rl/rsl_rl/env/wrappers/history_wrapper.py:18-49).  Third-party arithmetic that
is absent from /root/reference (`isaacgym.torch_utils`, Isaac Gym Preview, version unpinned) is
restated from its published formulas - SURVEY.md 8c: "parity unpinned" at that boundary.

PINNING: tests/test_oracle_golden.py checks this file against tests/golden/*.pt, which were produced
by running the unmodified reference in the build container (tests/golden/make_golden.py).

Arithmetic policy.  Everything that feeds an *index* (height-sample cells, argmin) is written as
IEEE-754 single operations in a fixed order so the CUDA kernels can match it bit for bit:
  * `torch.cross` on CPU contracts each component to one FMA, fma(a1,b2,-(a2*b1)) [measured];
  * `x.norm(dim=-1)` sums squares in float left to right and takes a correctly rounded sqrt [measured];
  * the 3-element `bmm` of quat_rotate_inverse is ((a0b0+a1b1)+a2b2) with rounded products [measured];
  * per-env mean/variance over the 693 grid heights are evaluated in float64 and rounded once (the
    reference's float cascade sum differs from that by <=1 ulp; the vectorised CPU sqrt is also
    1 ulp off in 0.6 % of elements) - tests therefore compare floats at 1e-6 and indices exactly,
    tolerating an index difference only where the two candidates' scores differ by < 1e-6.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may import this.
"""
import itertools
import math

import numpy as np
import torch

F32 = torch.float32


# ----------------------------------------------------------------------------- exact-op helpers
def _sqrt(x):
    return x.double().sqrt().to(F32)  # correctly rounded float sqrt


def _fma(a, b, c):
    return (a.double() * b.double() + c.double()).to(F32)


def _cross(a, b):
    """ATen CPU cross: component i = fma(a[i+1], b[i+2], -(a[i+2]*b[i+1]))."""
    a0, a1, a2 = a[..., 0], a[..., 1], a[..., 2]
    b0, b1, b2 = b[..., 0], b[..., 1], b[..., 2]
    return torch.stack([_fma(a1, b2, -(a2 * b1)), _fma(a2, b0, -(a0 * b2)), _fma(a0, b1, -(a1 * b0))], dim=-1)


def quat_apply(a, b):
    xyz = a[..., :3]
    t = _cross(xyz, b) * 2
    return b + a[..., 3:4] * t + _cross(xyz, t)


def quat_rotate_inverse(q, v):
    q_w = q[:, 3]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = _cross(q_vec, v) * q_w.unsqueeze(-1) * 2.0
    dot = (q_vec[:, 0] * v[:, 0] + q_vec[:, 1] * v[:, 1]) + q_vec[:, 2] * v[:, 2]
    c = q_vec * dot.unsqueeze(-1) * 2.0
    return a - b + c


def yaw_quat(q):
    """quat_apply_yaw's normalised yaw-only quaternion (legged_gym/utils/math.py:8-12)."""
    qz, qw = q[..., 2], q[..., 3]
    n = _sqrt(qz * qz + qw * qw).clamp(min=1e-9)
    z = torch.zeros_like(qz)
    return torch.stack([z, z, qz / n, qw / n], dim=-1)


def quat_from_euler_xyz(roll, pitch, yaw):
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qx, qy, qz, qw], dim=-1)


def wrap_to_pi(angles):
    angles = angles % (2 * np.pi)
    angles = angles - 2 * np.pi * (angles > np.pi)
    return angles


# ----------------------------------------------------------------------------- E5, E7-E10
def grid_points(K):
    """[693,3] base-frame sampling grid, x-major (legged_robot.py:1263-1277)."""
    x = torch.tensor(K.MEASURED_POINTS_X, dtype=F32)
    y = torch.tensor(K.MEASURED_POINTS_Y, dtype=F32)
    gx, gy = torch.meshgrid(x, y, indexing="ij")
    p = torch.zeros(gx.numel(), 3, dtype=F32)
    p[:, 0] = gx.flatten()
    p[:, 1] = gy.flatten()
    return p


def sample_points_world(root_states, grid):
    """R_yaw(q) * grid + root_pos, [N,693,3] (legged_robot.py:1300)."""
    qy = yaw_quat(root_states[:, 3:7])  # [N,4]
    N, P = root_states.shape[0], grid.shape[0]
    rot = quat_apply(qy[:, None, :].expand(N, P, 4), grid[None].expand(N, P, 3))
    return rot + root_states[:, None, 0:3]


def get_heights(root_states, grid, height_samples, K):
    """E5: legged_robot.py:1279-1317."""
    pts = sample_points_world(root_states, grid)
    pts = pts + K.BORDER_SIZE
    idx = (pts / K.HORIZONTAL_SCALE).long()
    px = idx[:, :, 0].clip(0, height_samples.shape[0] - 2)
    py = idx[:, :, 1].clip(0, height_samples.shape[1] - 2)
    h = torch.min(torch.min(height_samples[px, py], height_samples[px + 1, py]), height_samples[px, py + 1])
    return h * K.VERTICAL_SCALE


def foot_clearance(foot_positions, height_samples, K):
    """E6: legged_robot.py:1443-1472."""
    pts = foot_positions + K.BORDER_SIZE
    idx = (pts / K.HORIZONTAL_SCALE).long()
    px = idx[:, :, 0].clip(1, height_samples.shape[0] - 3)
    py = idx[:, :, 1].clip(1, height_samples.shape[1] - 3)
    hs = height_samples
    stack = torch.stack([hs[px, py], hs[px + 1, py], hs[px, py + 1], hs[px + 2, py], hs[px, py + 2],
                         hs[px + 1, py + 1], hs[px - 1, py], hs[px, py - 1], hs[px - 2, py], hs[px, py - 2]])
    return foot_positions[:, :, 2] - stack.max(dim=0)[0] * K.VERTICAL_SCALE


def raibert(root_states, thigh_pos, commands, base_lin_vel, K):
    """E7: legged_robot_dtc.py:100-115 (quirks kept: yaw-RATE command used as an angle; body-frame
    velocity added to world-frame positions)."""
    base_pos = root_states[:, None, 0:3]
    h2b = thigh_pos - base_pos
    th = commands[:, 2]
    c, s = torch.cos(th)[:, None], torch.sin(th)[:, None]
    # bmm(Rz, h2b^T): row0 = (c*x + (-s)*y) + 0*z ; row1 = (s*x + c*y) + 0*z ; row2 = (0*x + 0*y) + 1*z
    x, y, z = h2b[..., 0], h2b[..., 1], h2b[..., 2]
    rot = torch.stack([(c * x + (-s) * y) + 0 * z, (s * x + c * y) + 0 * z, (0 * x + 0 * y) + z], dim=-1)
    p_shoulder = base_pos + rot
    t_stance = K.SIM_DT * K.DECIMATION
    cmd = torch.cat([commands[:, :2], torch.zeros_like(commands[:, :1])], dim=1)
    v = base_lin_vel[:, None, :]
    p_sym = t_stance / 2 * v + 0.03 * (v - cmd[:, None, :])
    return p_shoulder + p_sym  # [N,4,3]


def foothold_select(root_states, measured_heights, pred_footholds, grid, K, debug=False):
    """E8-E10: legged_robot_dtc.py:127-201.  Returns dict."""
    N, P = measured_heights.shape
    GX, GY = K.GRID_X, K.GRID_Y
    g = (measured_heights - root_states[:, 2:3]).view(N, GX, GY)
    exc = (g > 1) | (g < -1)
    g = g.clamp(-0.5, 0.5)
    sp = torch.tensor(0.05, dtype=F32)
    dx = torch.empty_like(g)
    dy = torch.empty_like(g)
    dx[:, 1:-1] = (g[:, 2:] - g[:, :-2]) / sp / 2
    dx[:, 0] = (g[:, 1] - g[:, 0]) / sp
    dx[:, -1] = (g[:, -1] - g[:, -2]) / sp
    dy[:, :, 1:-1] = (g[:, :, 2:] - g[:, :, :-2]) / sp / 2
    dy[:, :, 0] = (g[:, :, 1] - g[:, :, 0]) / sp
    dy[:, :, -1] = (g[:, :, -1] - g[:, :, -2]) / sp
    slope = _sqrt(dx * dx + dy * dy)
    gd = g.double().view(N, -1)
    mean = gd.mean(1).to(F32)
    var = gd.var(1, unbiased=True).to(F32)
    rough = (g - mean[:, None, None]).abs()
    edge = _sqrt(var).clamp(0.0, 0.3)
    s = (0.2 * edge[:, None, None] + 1 * slope) + 0.3 * rough
    s = s.view(N, P)
    s = torch.where(s < 0.1, s, torch.tensor(10.0))

    hw = sample_points_world(root_states, grid).clone()
    hw[:, :, 2] = measured_heights
    diff = pred_footholds[:, None, :, :2] - hw[:, :, None, :2]  # [N,P,4,2]
    d = _sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])
    d = torch.where(d < 0.16, d, torch.tensor(10.0))
    nominal_idx = d.argmin(dim=1)  # first occurrence == CPU min(dim)[1] on ties
    score = s[:, :, None] * 0.2 + d * 0.8
    score = torch.where(exc.view(N, P, 1), torch.tensor(10.0), score)
    idx = score.argmin(dim=1)  # lowest-index argmin == torch.topk(k=1, largest=False) on CPU
    xi = torch.remainder(idx, GY)
    yi = torch.div(idx, GY, rounding_mode="trunc")
    mpx = torch.tensor(K.MEASURED_POINTS_X, dtype=F32)
    mpy = torch.tensor(K.MEASURED_POINTS_Y, dtype=F32)
    # quirk (legged_robot_dtc.py:184-192): x list indexed by idx%21, y list (tiled x4 = 84 long) by idx//21
    fobs = torch.cat([mpx[xi], mpy.repeat(4)[yi]], dim=1)
    opt_world = hw[torch.arange(N)[:, None], idx]
    out = dict(optimal_idx=idx, nominal_idx=nominal_idx, foothold_obs=fobs, optimal_footholds_world=opt_world,
               heights_world=hw)
    if debug:
        out.update(score=score, slope=slope, s=s, d=d, exc=exc)
    return out


# ----------------------------------------------------------------------------- the environment
class OracleEnv:
    """State + step() of LeggedRobotDTC for Lite3DTCCfg with the simulator stubbed by `gym` (FakeGym)."""

    def __init__(self, K, num_envs, height_samples, layout, gym, rng, robot_mass=12.0):
        self.K = K
        N = self.num_envs = num_envs
        self.gym = gym
        self.rng = rng
        self.device = "cpu"
        self.num_obs, self.num_privileged_obs, self.num_actions = K.NUM_OBS, K.NUM_PRIV, K.NUM_ACTIONS
        self.max_episode_length = float(K.MAX_EPISODE_LENGTH)
        self.dt = K.DT
        self.height_samples = torch.as_tensor(np.asarray(height_samples)).view(K.MAP_ROWS, K.MAP_COLS)
        self.terrain_levels, self.terrain_types, self.env_origins, self.terrain_origins = (
            t.clone() for t in layout)
        self.max_terrain_level = K.NUM_ROWS
        self.grid = grid_points(K)
        z = lambda *s, **k: torch.zeros(*s, dtype=k.get("dtype", F32))
        self.root_states = gym.root_states
        self.dof_state = gym.dof_state
        self.rigid_body_state = gym.rigid_body_state
        self.contact_forces = gym.net_contact_force.view(N, -1, 3)
        self.dof_pos = self.dof_state.view(N, K.NUM_DOF, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, K.NUM_DOF, 2)[..., 1]
        self.base_quat = self.root_states[:, 3:7]
        self.base_pos = self.root_states[:, :3]
        self.obs_buf = z(N, K.NUM_OBS)
        self.privileged_obs_buf = z(N, K.NUM_PRIV)
        self.rew_buf = z(N)
        self.reset_buf = torch.ones(N, dtype=torch.long)
        self.episode_length_buf = z(N, dtype=torch.long)
        self.time_out_buf = z(N, dtype=torch.bool)
        self.extras = {}
        self.common_step_counter = 0
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0]).repeat(N, 1)
        self.forward_vec = torch.tensor([1.0, 0.0, 0.0]).repeat(N, 1)
        self.torques = z(N, 12)
        self.actions = z(N, 12)
        self.last_actions = z(N, 12)
        self.last_actions_2 = z(N, 12)
        self.last_dof_vel = z(N, 12)
        self.last_root_vel = z(N, 6)
        self.last_foot_velocities = z(N, 4, 3)
        self.commands = z(N, 4)
        self.feet_air_time = z(N, 4)
        self.last_contacts = z(N, 4, dtype=torch.bool)
        self.contact_filt = z(N, 4, dtype=torch.bool)
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 7:10])
        self.base_ang_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 10:13])
        self.projected_gravity = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        self.pitch_est = z(N)
        self.lag_buffer = [z(N, 12) for _ in range(6)]
        self.stumb_buffer = [z(N, 4, dtype=torch.bool) for _ in range(5)]
        self.measured_heights = z(N, K.NUM_POINTS)
        self.measured_foot_clearance = z(N, 4)
        self.forces = z(N, K.NUM_BODIES, 3)
        self.height_noise_offset = z(N, K.NUM_POINTS)
        self.lin_vel_buffer = z(10, N, 2)
        self.ang_vel_buffer = z(10, N, 1)
        self.cmd_buffer = z(10, N, 4)
        self.default_dof_pos = torch.tensor(K.DEFAULT_DOF_POS, dtype=F32).unsqueeze(0)
        self.dof_pos_limits = torch.tensor(K.soft_dof_pos_limits(), dtype=F32)
        self.torque_limits = torch.full((12,), K.TORQUE_LIMIT)
        self.p_gains = torch.full((12,), K.P_GAIN)
        self.d_gains = torch.full((12,), K.D_GAIN)
        self.Kp_factors = torch.ones(N, 12)
        self.Kd_factors = torch.ones(N, 12)
        self.motor_strengths = torch.ones(N, 12)
        self.motor_offsets = z(N, 12)
        self.robot_mass = torch.full((N,), float(robot_mass))
        self.base_init_state = torch.tensor(K.BASE_INIT_STATE, dtype=F32)
        self.foothold_obs = z(N, 8)
        self.commands_scale = torch.tensor([K.OBS_SCALES["lin_vel"], K.OBS_SCALES["lin_vel"], K.OBS_SCALES["ang_vel"]])
        nv = z(K.NUM_OBS)
        ns, os_ = K.NOISE_SCALES, K.OBS_SCALES
        nv[:3] = ns["ang_vel"] * 1.0 * os_["ang_vel"]
        nv[3:6] = ns["gravity"] * 1.0
        nv[9:21] = ns["dof_pos"] * 1.0 * os_["dof_pos"]
        nv[21:33] = ns["dof_vel"] * 1.0 * os_["dof_vel"]
        self.noise_scale_vec = nv
        self.reward_scales = {k: v * self.dt for k, v in K.REWARD_SCALES.items()}
        self.reward_names = list(K.REWARD_NAMES)
        self.episode_sums = {k: z(N) for k in K.EPISODE_SUM_NAMES}
        acc = np.array([list(i) for i in itertools.product([-1, 1], repeat=3)]) * [0.3, 0.2, 0.15] / 2.0
        self.acc_point = torch.tensor(acc, dtype=F32).view(1, 8, 3)
        # constant LS plane-fit operator (A^T A)^-1 A^T of get_plane_norm (legged_robot.py:1535-1547)
        A = self.grid.clone()
        A[:, 2] = 1
        A = A[None]
        self.plane_op = torch.bmm(torch.linalg.inv(torch.bmm(A.transpose(1, 2), A)), A.transpose(1, 2))[0]  # [3,693]

    # ------------------------------------------------------------------ API surface (vec_env.py:36-59)
    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def get_reward_buf(self):
        return self.rew_buf

    def get_base_vel(self):
        return self.base_lin_vel * self.K.OBS_SCALES["lin_vel"]

    def reset(self):
        self.reset_idx(torch.arange(self.num_envs))
        obs, priv, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions))
        return obs, priv

    # ------------------------------------------------------------------ E1, E2
    def step(self, actions):
        K = self.K
        self.actions = torch.clip(actions, -K.CLIP_ACTIONS, K.CLIP_ACTIONS)
        for _ in range(K.DECIMATION):
            self.torques = self._compute_torques(self.actions)
            self.gym.simulate(None)
            self.gym.refresh_dof_state_tensor(None)
        self.post_physics_step()
        self.obs_buf = torch.clip(self.obs_buf, -K.CLIP_OBS, K.CLIP_OBS)
        self.privileged_obs_buf = torch.clip(self.privileged_obs_buf, -K.CLIP_OBS, K.CLIP_OBS)
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _compute_torques(self, actions):
        K = self.K
        choice = self.rng.np_randint(1, 5)
        scaled = actions * K.ACTION_SCALE
        self.lag_buffer = self.lag_buffer[1:] + [scaled.clone()]
        goal = torch.clip(self.lag_buffer[choice] + self.default_dof_pos, self.dof_pos_limits[:, 0], self.dof_pos_limits[:, 1])
        tq = self.p_gains * self.Kp_factors * (goal - self.dof_pos + self.motor_offsets) - self.d_gains * self.Kd_factors * self.dof_vel
        tq = tq * self.motor_strengths
        return torch.clip(tq, -self.torque_limits, self.torque_limits)

    # ------------------------------------------------------------------ E3..E14
    def post_physics_step(self):
        K = self.K
        N = self.num_envs
        self.gym.refresh_actor_root_state_tensor(None)
        self.gym.refresh_net_contact_force_tensor(None)
        self.gym.refresh_rigid_body_state_tensor(None)
        self.episode_length_buf += 1
        self.common_step_counter += 1
        self.base_lin_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 7:10])
        self.base_ang_vel = quat_rotate_inverse(self.base_quat, self.root_states[:, 10:13])
        self.lin_vel_buffer[:-1] = self.lin_vel_buffer[1:].clone()
        self.lin_vel_buffer[-1] = self.base_lin_vel[:, :2]
        self.ang_vel_buffer[:-1] = self.ang_vel_buffer[1:].clone()
        self.ang_vel_buffer[-1] = self.base_ang_vel[:, 2].unsqueeze(1)
        self.cmd_buffer[:-1] = self.cmd_buffer[1:].clone()
        self.cmd_buffer[-1] = self.commands
        self.projected_gravity = quat_rotate_inverse(self.base_quat, self.gravity_vec)
        rb = self.rigid_body_state.view(N, K.NUM_BODIES, 13)
        self.foot_velocities = rb[:, K.FEET_INDICES, 7:10]
        self.foot_positions = rb[:, K.FEET_INDICES, 0:3]

        self._post_physics_step_callback()

        self.hip_positions = rb[:, K.THIGH_INDICES, 0:3]
        self.pred_footholds = raibert(self.root_states, self.hip_positions, self.commands, self.base_lin_vel, K)
        sel = foothold_select(self.root_states, self.measured_heights, self.pred_footholds, self.grid, K)
        self.optimal_foothold_indice = sel["optimal_idx"].unsqueeze(1)
        self.nominal_footholds_indice = sel["nominal_idx"]
        self.foothold_obs = sel["foothold_obs"]
        self.optimal_footholds_world = sel["optimal_footholds_world"]
        self.heights_world = sel["heights_world"]

        self.check_termination()
        self.compute_reward()
        env_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        self.reset_idx(env_ids)
        self.compute_observations()
        self.last_actions_2 = self.last_actions.clone()
        self.last_actions = self.actions.clone()
        self.last_dof_vel = self.dof_vel.clone()
        self.last_root_vel = self.root_states[:, 7:13].clone()
        self.last_foot_velocities = self.foot_velocities.clone()

    def _post_physics_step_callback(self):
        K = self.K
        N = self.num_envs
        env_ids = (self.episode_length_buf % K.RESAMPLING_STEPS == 0).nonzero(as_tuple=False).flatten()
        self._resample_commands(env_ids)
        forward = quat_apply(self.base_quat, self.forward_vec)
        heading = torch.atan2(forward[:, 1], forward[:, 0])
        self.commands[:, 2] = torch.clip(0.5 * wrap_to_pi(self.commands[:, 3] - heading), -1.5, 1.5)
        self.measured_heights = get_heights(self.root_states, self.grid, self.height_samples, K)
        if self.common_step_counter % K.PUSH_INTERVAL in range(2):
            if self.common_step_counter % K.PUSH_INTERVAL == 0:
                # max_push_force_xy = 0 -> forces stay 0; draws are still consumed (legged_robot.py:550-551)
                self.forces[:, 0, 0:2] = (0.0 - -0.0) * self.rng.rand(N, 2) + -0.0
                _ = self.rng.rand(N, 3)
            mv = K.MAX_PUSH_VEL_XY
            self.root_states[:, 7:9] = (mv - -mv) * self.rng.rand(N, 2) + -mv
        else:
            self.forces = torch.zeros(N, K.NUM_BODIES, 3)
        self.measured_foot_clearance = foot_clearance(self.foot_positions, self.height_samples, K)
        contact = self.contact_forces[:, K.FEET_INDICES, 2] > 1.0
        self.contact_filt = torch.logical_or(contact, self.last_contacts)
        self.last_contacts = contact

    def _resample_commands(self, env_ids):
        R = self.K.CMD_RANGES
        n = len(env_ids)
        u = lambda lo, hi: ((hi - lo) * self.rng.rand(n, 1) + lo).squeeze(1)
        self.commands[env_ids, 0] = u(*R["lin_vel_x"])
        self.commands[env_ids, 1] = u(*R["lin_vel_y"])
        self.commands[env_ids, 3] = u(*R["heading"])
        nrm = _sqrt(self.commands[env_ids, 0] ** 2 + self.commands[env_ids, 1] ** 2)
        self.commands[env_ids, :2] *= (nrm > 0.1).unsqueeze(1)
        self.forces[env_ids, :] = 0.0

    def check_termination(self):
        K = self.K
        self.reset_buf = torch.zeros(self.num_envs, dtype=torch.bool)  # empty termination-contact set
        self.time_out_buf = self.episode_length_buf > self.max_episode_length
        self.reset_buf |= self.time_out_buf
        self.reset_buf |= self.projected_gravity[:, 2] > 0.2
        c = self.measured_heights[:, 10 * K.GRID_Y:(K.GRID_X - 10) * K.GRID_Y].clip(min=-0.0)
        self.reset_buf |= torch.mean(self.root_states[:, 2].unsqueeze(1) - c, dim=1) < 0.15

    def compute_reward(self):
        self.rew_buf = torch.zeros(self.num_envs)
        self.reward_terms = {}
        for name in self.reward_names:
            rew = getattr(self, "_reward_" + name)() * self.reward_scales[name]
            self.reward_terms[name] = rew
            self.rew_buf = self.rew_buf + rew
            self.episode_sums[name] += rew
        rew = (self.reset_buf * ~self.time_out_buf) * self.reward_scales["termination"]
        self.rew_buf = self.rew_buf + rew
        self.episode_sums["termination"] += rew

    def reset_idx(self, env_ids):
        K = self.K
        n = len(env_ids)
        if n == 0:
            return
        # terrain curriculum (legged_robot.py:690-711)
        dxy = self.root_states[env_ids, :2] - self.env_origins[env_ids, :2]
        distance = _sqrt(dxy[:, 0] ** 2 + dxy[:, 1] ** 2)
        move_up = distance > K.TERRAIN_LENGTH * 0.6
        cn = _sqrt(self.commands[env_ids, 0] ** 2 + self.commands[env_ids, 1] ** 2)
        move_down = (distance < cn * K.EPISODE_LENGTH_S * 0.5) * ~move_up
        self.terrain_levels[env_ids] += 1 * move_up - 1 * move_down
        lv = self.terrain_levels[env_ids]
        self.terrain_levels[env_ids] = torch.where(lv >= self.max_terrain_level,
                                                   self.rng.randint_like(lv, self.max_terrain_level),
                                                   torch.clip(lv, 0))
        self.env_origins[env_ids] = self.terrain_origins[self.terrain_levels[env_ids], self.terrain_types[env_ids]]
        # dofs, root (legged_robot.py:640-641, legged_robot_dtc.py:299-311)
        self.dof_pos[env_ids] = self.default_dof_pos * ((1.5 - 0.5) * self.rng.rand(n, 12) + 0.5)
        self.dof_vel[env_ids] = 0.0
        self.root_states[env_ids] = self.base_init_state
        self.root_states[env_ids, :3] += self.env_origins[env_ids]
        self.root_states[env_ids, :2] += (0.5 - -0.5) * self.rng.rand(n, 2) + -0.5
        self.root_states[env_ids, 7:13] = (0.5 - -0.5) * self.rng.rand(n, 6) + -0.5
        self._resample_commands(env_ids)
        lo, hi = K.MOTOR_STRENGTH_RANGE
        self.motor_strengths[env_ids, :] = self.rng.rand(n).unsqueeze(1) * (hi - lo) + lo
        self.height_noise_offset[env_ids] = self.height_noise_offset[env_ids] * 0.0
        self.height_noise_offset[env_ids] += self.rng.np_normal(0, 0.02)
        self.last_actions[env_ids] = 0.0
        self.last_actions_2[env_ids] = 0.0
        self.last_dof_vel[env_ids] = 0.0
        self.feet_air_time[env_ids] = 0.0
        self.episode_length_buf[env_ids] = 0
        self.reset_buf[env_ids] = True
        self.pitch_est[env_ids] = 0
        for b in self.lag_buffer:
            b[env_ids, :] = 0
        for b in self.stumb_buffer:
            b[env_ids, :] = False
        self.extras["episode"] = {}
        for key in self.episode_sums:
            self.extras["episode"]["rew_" + key] = torch.mean(self.episode_sums[key][env_ids]) / K.EPISODE_LENGTH_S
            self.episode_sums[key][env_ids] = 0.0
        self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())
        self.extras["time_outs"] = self.time_out_buf
        self.contact_filt[env_ids] = False
        self.last_contacts[env_ids] = False
        self.lin_vel_buffer[:, env_ids] = 0.0
        self.ang_vel_buffer[:, env_ids] = 0.0
        self.cmd_buffer[:, env_ids] = 0.0

    def compute_observations(self):
        K = self.K
        S = K.OBS_SCALES
        self.obs_buf = torch.cat((self.base_ang_vel * S["ang_vel"], self.projected_gravity,
                                  self.commands[:, :3] * self.commands_scale,
                                  (self.dof_pos - self.default_dof_pos) * S["dof_pos"],
                                  self.dof_vel * S["dof_vel"], self.actions, self.foothold_obs), dim=-1)
        self.heights = torch.clip(self.root_states[:, 2].unsqueeze(1) - K.BASE_HEIGHT_TARGET - self.measured_heights,
                                  -1, 1.0) * S["height_measurements"]
        self.privileged_obs_buf = torch.cat((
            self.heights + (2 * self.rng.rand_like(self.heights) - 1) * 0.1 + self.height_noise_offset,
            self.forces[:, 0, :] * S["force"], self.heights), dim=1)
        self.obs_buf = self.obs_buf + (2 * self.rng.rand_like(self.obs_buf) - 1) * self.noise_scale_vec

    # ------------------------------------------------------------------ R*: 23 reward terms
    def _reward_action_rate(self):
        return torch.sum(torch.square(self.last_actions - self.actions), dim=1)

    def _reward_ang_vel_xy(self):
        return torch.sum(torch.square(self.base_ang_vel[:, :2]), dim=1)

    def _reward_base_height(self):
        f2b = self.root_states[:, 2] - torch.mean(self.foot_positions[:, :, 2], dim=-1)
        return torch.square(f2b - self.K.BASE_HEIGHT_TARGET)

    def _reward_collision(self):
        f = self.contact_forces[:, self.K.PENALISED_CONTACT_INDICES, :]
        return torch.sum(1.0 * (torch.norm(f, dim=-1) > 0.1), dim=1)

    def _reward_dof_acc(self):
        return torch.sum(torch.square((self.last_dof_vel - self.dof_vel) / self.dt), dim=1)

    def _reward_dof_pos_limits(self):
        out = -(self.dof_pos - self.dof_pos_limits[:, 0]).clip(max=0.0)
        out = out + (self.dof_pos - self.dof_pos_limits[:, 1]).clip(min=0.0)
        return torch.sum(out, dim=1)

    def _reward_feet_air_time(self):
        contact = self.contact_forces[:, self.K.FEET_INDICES, 2] > 1.0
        contact_filt = torch.logical_or(contact, self.last_contacts)
        self.last_contacts = contact
        first_contact = (self.feet_air_time > 0.0) * contact_filt
        self.feet_air_time += self.dt
        rew = torch.sum((self.feet_air_time - 0.5) * first_contact, dim=1)
        rew = rew * (torch.norm(self.commands[:, :2], dim=1) > 0.1)
        self.feet_air_time *= ~contact_filt
        return rew

    def _reward_feet_slip(self):
        contact = self.contact_forces[:, self.K.FEET_INDICES, 2] > 1.0
        contact_filt = torch.logical_or(contact, self.last_contacts)
        fv = torch.square(torch.norm(self.foot_velocities[:, :, 0:2], dim=2))
        return torch.sum(contact_filt * fv, dim=1)

    def _reward_foot_acc(self):
        mask = torch.where(self.terrain_levels > 5, 0.2, 1.0)
        a = torch.norm((self.last_foot_velocities - self.foot_velocities) / self.dt, dim=-1)
        return torch.sum((mask.view(-1, 1) * (a - self.K.MAX_ACC)).clip(min=0.0), dim=1)

    def _reward_foot_clearance(self):
        f = self.contact_forces[:, self.K.FEET_INDICES, :]
        stumb = torch.norm(f[:, :, :2], dim=2) > 4 * torch.abs(f[:, :, 2])
        self.stumb_buffer = self.stumb_buffer[1:] + [stumb.clone()]
        flag = self.stumb_buffer[0] | self.stumb_buffer[1] | self.stumb_buffer[2] | self.stumb_buffer[3] | self.stumb_buffer[4]
        return torch.sum(~flag * (self.measured_foot_clearance > 0.18), dim=1)

    def _reward_foothold_miss(self):
        mz = torch.min(self.foot_positions[:, :, -1], dim=-1)[0]
        return torch.where(mz < 0, torch.tensor(1.0), torch.tensor(0.0))

    def _reward_hip_pos(self):
        return torch.sum(torch.square(self.dof_pos[:, self.K.HIP_DOF_INDICES]), dim=1)

    def _reward_lin_vel_z(self):
        return torch.square(self.base_lin_vel[:, 2])

    def _reward_orientation(self):
        X = self.measured_heights @ self.plane_op.t()  # [N,3] == bmm(A3, heights)
        pv = torch.stack([X[:, 0], X[:, 1], -torch.ones_like(X[:, 0])], dim=1)
        pv = pv / torch.norm(pv, dim=-1, keepdim=True)
        p_norm = -pv
        pitch = torch.atan(p_norm[:, 0])
        roll = -torch.atan(p_norm[:, 1])
        zero = torch.tensor(0.0)
        pitch_c = torch.where((pitch >= -0.1) & (pitch <= 0.1), zero, pitch)
        roll_c = torch.where((roll >= -0.1) & (roll <= 0.1), zero, roll)
        self.pitch_est = self.pitch_est * 0.2 + 0.8 * pitch_c
        quat = quat_from_euler_xyz(roll_c, self.pitch_est, torch.zeros_like(roll))
        loc = quat_rotate_inverse(quat, self.gravity_vec)
        return torch.sum(torch.square(self.projected_gravity[:, :1] - loc[:, :1]), dim=1)

    def _reward_pos_acc(self):
        N = self.num_envs
        v = self.base_lin_vel.reshape(N, 1, 3) + torch.linalg.cross(
            self.base_ang_vel.reshape(N, 1, 3).repeat(1, 8, 1), self.acc_point.repeat(N, 1, 1))
        return torch.sum(torch.square(torch.norm(v, dim=-1)), dim=1)

    def _reward_power(self):
        return torch.sum(torch.clip(self.torques * self.dof_vel, min=0), dim=1)

    def _reward_powerchange(self):
        co = self.commands[:, 0].clip(min=1.0)
        return (torch.sum((self.torques * self.dof_vel).clip(min=0.0), dim=1) / (self.robot_mass * 9.815 * co)) ** 2

    def _reward_smooth(self):
        return torch.sum(torch.square(self.actions - 2 * self.last_actions + self.last_actions_2), dim=1)

    def _reward_soft_tracking_ang_vel(self, tolerance=0.15, lookback=4):
        K = self.K
        d = torch.square((self.cmd_buffer[-lookback:, :, 2] - self.ang_vel_buffer[-lookback:, :].squeeze(-1))
                         / K.CMD_RANGES["ang_vel_yaw"][1])
        d = torch.where(d <= tolerance ** 2, 0.0, 1.0)
        return torch.mean(torch.exp(-d / K.TRACKING_SIGMA), dim=0)

    def _reward_soft_tracking_lin_vel(self, lookback=3):
        K = self.K
        d = torch.sum(torch.square((self.cmd_buffer[-lookback:, :, :2] - self.lin_vel_buffer[-lookback, :, :2])
                                   / K.CMD_RANGES["lin_vel_x"][1]), dim=-1)
        return torch.mean(torch.exp(-d / K.TRACKING_SIGMA), dim=0)

    def _reward_stand_still(self):
        return torch.sum(torch.abs(self.dof_pos - self.default_dof_pos), dim=1) * (torch.norm(self.commands[:, :2], dim=1) < 0.1)

    def _reward_torques(self):
        return torch.sum(torch.square(self.torques), dim=1)

    def _reward_tracking_optimal_footholds(self):
        dis = torch.norm(self.foot_positions[:, :, :-1] - self.optimal_footholds_world[:, :, :-1], dim=-1)
        per_foot = -torch.log(0.8 + dis)
        filt = torch.where(self.contact_filt.float() == 1, per_foot, torch.tensor(0.0))
        return torch.sum(filt, dim=-1)


class OracleHistoryWrapper:
    """history_wrapper.py:6-53 (obs_history is never cleared on episode resets - reference quirk)."""

    def __init__(self, env):
        self.env = env
        self.num_envs, self.num_obs = env.num_envs, env.num_obs
        self.num_privileged_obs, self.num_actions = env.num_privileged_obs, env.num_actions
        self.max_episode_length = env.max_episode_length
        self.device = env.device
        self.num_obs_history = env.K.NUM_HIST * env.num_obs
        self.obs_history = torch.zeros(env.num_envs, self.num_obs_history)
        self.episode_length_buf = env.episode_length_buf  # runner overwrites the WRAPPER attribute only

    def _pack(self, obs, priv):
        return {"obs": obs, "privileged_obs": priv, "obs_history": self.obs_history, "base_vel": self.env.get_base_vel()}

    def step(self, action):
        obs, priv, rew, done, info = self.env.step(action)
        self.obs_history = torch.cat((self.obs_history[:, self.env.num_obs:], obs), dim=-1)
        return self._pack(obs, priv), rew, done, info

    def get_observations(self):
        obs = self.env.get_observations()
        self.obs_history = torch.cat((self.obs_history[:, self.env.num_obs:], obs), dim=-1)
        return self._pack(obs, self.env.get_privileged_observations())

    def reset(self):
        obs, _ = self.env.reset()
        self.obs_history[:, :] = 0
        return {"obs": obs, "privileged_obs": self.env.get_privileged_observations(),
                "obs_history": self.obs_history, "base_vel": self.env.get_base_vel()}

    def get_reward_buf(self):
        return self.env.get_reward_buf()
