"""`LeggedRobotDTC` - the environment surface of the hot path, backed by the sm_100a kernels.

Keeps the reference's attribute and method names (legged_gym/envs/base/legged_robot_dtc.py:29-288,
legged_robot.py:92-122, base_task.py:41-119, rsl_rl/env/vec_env.py:36-59).  Everything the reference computes
between two physics calls runs in five kernel launches (csrc/dtc_env.cu, csrc/dtc_foothold.cu); this class only
owns the torch tensors, draws the two host-side random numbers the reference draws on the host
(np.random.randint lag choice, np.random.normal reset offset) and sequences the launches.

Scene construction (create_sim/_create_envs, viewer, debug drawing) is out of scope (SURVEY.md section 2 rows
1-3): the simulator is whatever object `gym` is - FakeGym on the benchmark path - exposing the four state tensors.
"""
import ctypes as C
import math

import numpy as np
import torch

from .... import _lib as B
from .... import lite3 as L
from .cfg_resolve import resolve


class _Scales:
    def __init__(self, d):
        self.__dict__.update(d)


class LeggedRobotDTC:
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True, *, gym,
                 height_samples, terrain_origins, layout, seed=0, foothold_variant=6, robot_mass=12.0, host_rng=False):
        self.cfg = cfg
        self.device = torch.device(sim_device)
        if self.device.type != "cuda":
            raise B.DtcError("LeggedRobotDTC runs on a CUDA device only (no CPU fallback)")
        self.lib = B.lib()
        self.gym = gym
        self.sim = getattr(gym, "sim", None)
        self._unwrap = getattr(gym, "unwrap_tensor", None) or (lambda t: t)  # gymtorch.unwrap_tensor for Isaac Gym's tensor API
        self.headless = headless
        self.viewer = None
        self.debug_viz = False
        self.foothold_variant = foothold_variant
        # every number below comes from the configuration object, the way _parse_cfg / _prepare_reward_function read it
        # (legged_robot.py:1230-1240, 929-952); settings the fused kernels cannot honour raise CfgError
        R = self._rc = resolve(cfg, sim_dt=getattr(sim_params, "dt", None))
        N = self.num_envs = R.num_envs
        if gym.num_envs != N:
            raise ValueError("gym.num_envs != cfg.env.num_envs")
        self.num_obs, self.num_privileged_obs, self.num_actions = L.NUM_OBS, L.NUM_PRIV, L.NUM_ACTIONS
        self.num_bodies, self.num_dof = L.NUM_BODIES, L.NUM_DOF
        self.dt = R.dt
        self.max_episode_length_s = R.episode_length_s
        self.max_episode_length = float(R.max_episode_length)
        self.obs_scales = _Scales(R.obs_scales)
        self.reward_scales = dict(R.reward_scales)  # non-zero scales x dt (legged_robot.py:934-940)
        self.reward_names = [k for k in L.REWARD_NAMES if k in R.reward_scales]
        self.command_ranges = {k: list(v) for k, v in R.command_ranges.items()}
        self.seed = int(seed)
        # the reference draws two numbers per step on the HOST (np.random.randint lag choice, np.random.normal reset offset,
        # legged_robot.py:608,230).  host_rng=False (default): both come from the kernels' Philox streams, step() touches no host
        # generator (CUDA-graph capturable); host_rng=True: a seeded numpy generator, consumed exactly where the reference does
        self.host_rng = bool(host_rng)
        self.np_rng = np.random.default_rng(seed)
        self.common_step_counter = 0
        self.init_done = True
        dev = self.device
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        u8 = lambda *s: torch.zeros(*s, device=dev, dtype=torch.uint8)
        i64 = lambda *s: torch.zeros(*s, device=dev, dtype=torch.int64)
        # simulator tensors (legged_robot.py:759-779)
        self.root_states = gym.acquire_actor_root_state_tensor(None)
        self.dof_state = gym.acquire_dof_state_tensor(None)
        self.contact_forces = gym.acquire_net_contact_force_tensor(None).view(N, -1, 3)
        self.rigid_body_state = gym.acquire_rigid_body_state_tensor(None)
        for t in (self.root_states, self.dof_state, self.contact_forces, self.rigid_body_state):
            B.require_cuda(t, "simulator tensor")
        self.dof_pos = self.dof_state.view(N, L.NUM_DOF, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, L.NUM_DOF, 2)[..., 1]
        self.base_quat = self.root_states[:, 3:7]
        self.base_pos = self.root_states[:, :3]
        # numpy (host generator) or a tensor, e.g. straight from sim_stub.make_heightmap_device / dtc_terrain_rasterize
        hs_t = height_samples if torch.is_tensor(height_samples) else torch.as_tensor(np.asarray(height_samples))
        if hs_t.dim() != 2 or hs_t.dtype != torch.int16:
            raise ValueError("height_samples must be an int16 [rows, cols] heightfield (legged_gym/utils/terrain.py:26-30)")
        self.height_samples = hs_t.to(dev).contiguous()
        levels, types, origins, tor = layout
        self.terrain_levels = levels.to(dev).clone()
        self.terrain_types = types.to(dev).clone()
        self.env_origins = origins.to(dev).float().contiguous().clone()
        self.terrain_origins = tor.to(dev).float().contiguous().clone()
        self.max_terrain_level = R.num_rows
        self.custom_origins = True
        # buffers (base_task.py:41-52; legged_robot.py:788-846)
        self.priv_ld, self.hist_ld = 1392, 268
        self._priv_store = f(N, self.priv_ld)
        self.privileged_obs_buf = self._priv_store[:, :L.NUM_PRIV]
        self.obs_buf = f(N, L.NUM_OBS)
        self._hist_store = f(N, self.hist_ld)
        self.rew_buf = f(N)
        self.reset_buf = torch.ones(N, device=dev, dtype=torch.uint8)
        self.episode_length_buf = i64(N)
        self.time_out_buf = u8(N)
        self.actions, self.torques = f(N, 12), f(N, 12)
        self._lag = f(6, N, 12)
        self.base_lin_vel, self.base_ang_vel, self.projected_gravity = f(N, 3), f(N, 3), f(N, 3)
        self._base_vel = f(N, 3)
        self.commands = f(N, 4)
        self.cmd_buffer, self.lin_vel_buffer, self.ang_vel_buffer = f(10, N, 4), f(10, N, 2), f(10, N, 1)
        self.measured_heights = f(N, L.NUM_POINTS)
        self.pred_footholds = f(N, 4, 3)
        self._optimal_idx = torch.zeros(N, 4, device=dev, dtype=torch.int32)
        self._nominal_idx = torch.zeros(N, 4, device=dev, dtype=torch.int32)
        self.foothold_obs = f(N, 8)
        self.optimal_footholds_world = f(N, 4, 3)
        self._center_clear_mean, self._plane_ab = f(N), f(N, 2)
        self.measured_foot_clearance = f(N, 4)
        self._contact_filt, self._last_contacts, self._stumb = u8(N, 4), u8(N, 4), u8(5, N, 4)
        self.feet_air_time, self.pitch_est = f(N, 4), f(N)
        self.last_actions, self.last_actions_2, self.last_dof_vel = f(N, 12), f(N, 12), f(N, 12)
        self.last_root_vel, self.last_foot_velocities = f(N, 6), f(N, 4, 3)
        self.motor_strengths = torch.ones(N, 12, device=dev)
        self.robot_mass = torch.full((N,), float(robot_mass), device=dev)
        self._height_noise_offset = f(N)
        self._forces0 = f(N, 3)
        self._episode_sums = f(24, N)
        self._reward_terms = f(24, N)
        self._episode_stats, self._episode_stats_last = f(26), f(26)
        self._time_outs_sent = u8(N)
        # keys = the non-zero reward scales, like the reference's episode_sums (legged_robot.py:950-952)
        self.episode_sums = {k: self._episode_sums[i] for i, k in enumerate(L.EPISODE_SUM_NAMES) if k in self.reward_scales}
        self.default_dof_pos = torch.tensor(R.default_dof_pos, device=dev).unsqueeze(0)
        self.dof_pos_limits = torch.tensor(R.dof_pos_limits, device=dev)
        self.torque_limits = torch.full((12,), R.torque_limit, device=dev)
        self.p_gains, self.d_gains = torch.tensor(R.p_gains, device=dev), torch.tensor(R.d_gains, device=dev)
        self.feet_indices = torch.tensor(L.FEET_INDICES, device=dev)
        self.thigh_indices = torch.tensor(L.THIGH_INDICES, device=dev)
        self.noise_scale_vec = torch.tensor(R.noise_scale_vec, device=dev)
        self.add_noise = R.add_noise
        self._debug_score = None
        self._noise = None  # injected draws (tests): dict of CUDA tensors keyed like dtc_env_noise
        self._host_draws = None  # injected host draws: dict(lag=[4 ints], reset_normal=float)
        self._make_ctx()
        self.extras = _Extras(self)
        # CUDA-graph support (include/dtc_b200.h: dtc_env_set_step_base): kernels receive common_step_counter - _step_offset and add
        # the device-side copy of _step_offset, so a captured rollout can be replayed with an advancing step counter
        self._step_offset = 0
        self._step_base = torch.zeros(1, device=dev, dtype=torch.int64)
        B.check(self.lib.dtc_env_set_step_base(self._h, B.ptr(self._step_base)), "dtc_env_set_step_base")

    # ------------------------------------------------------------------ construction helpers
    def _plane_op(self):
        # constant (A^T A)^-1 A^T of get_plane_norm (legged_robot.py:1541-1546); same torch ops, batch of one
        x = torch.tensor(self._rc.grid_x)
        y = torch.tensor(self._rc.grid_y)
        gx, gy = torch.meshgrid(x, y, indexing="ij")
        A = torch.stack([gx.flatten(), gy.flatten(), torch.ones(L.NUM_POINTS)], dim=1)[None]
        return torch.bmm(torch.linalg.inv(torch.bmm(A.transpose(1, 2), A)), A.transpose(1, 2))[0]

    def _make_ctx(self):
        R, c = self._rc, B.EnvConfig()
        c.num_envs, c.map_rows, c.map_cols = self.num_envs, int(self.height_samples.shape[0]), int(self.height_samples.shape[1])
        c.horizontal_scale, c.vertical_scale, c.border_size = R.horizontal_scale, R.vertical_scale, R.border_size
        c.dt = R.dt
        c.max_episode_length, c.resampling_steps, c.push_interval = R.max_episode_length, R.resampling_steps, R.push_interval
        c.max_push_vel_xy, c.push_robots = R.max_push_vel_xy, int(R.push_robots)
        for name, key in (("cmd_lin_x", "lin_vel_x"), ("cmd_lin_y", "lin_vel_y"), ("cmd_heading", "heading")):
            lo, hi = R.command_ranges[key]
            getattr(c, name)[0], getattr(c, name)[1] = lo, hi - lo
        lo, hi = R.motor_strength
        c.motor_strength[0], c.motor_strength[1] = lo, hi - lo
        c.cmd_lin_x_max, c.cmd_ang_yaw_max = R.command_ranges["lin_vel_x"][1], R.command_ranges["ang_vel_yaw"][1]
        c.base_height_target, c.tracking_sigma, c.max_acc = R.base_height_target, R.tracking_sigma, R.max_acc
        c.terrain_length, c.max_terrain_level, c.num_terrain_cols = R.terrain_length, R.num_rows, R.num_cols
        c.terrain_curriculum = int(R.terrain_curriculum)
        c.episode_length_s = R.episode_length_s
        c.action_scale, c.torque_limit = R.action_scale, R.torque_limit
        for j in range(12):
            c.p_gains[j], c.d_gains[j] = R.p_gains[j], R.d_gains[j]
            c.default_dof_pos[j] = R.default_dof_pos[j]
            c.dof_pos_lower[j], c.dof_pos_upper[j] = R.dof_pos_limits[j]
        for j in range(13):
            c.base_init_state[j] = R.base_init_state[j]
        for j in range(33):
            c.grid_x[j] = R.grid_x[j]
        for j in range(21):
            c.grid_y[j] = R.grid_y[j]
        po = self._plane_op()
        flat = po[:2].contiguous().flatten().tolist()
        for j, v in enumerate(flat):
            c.plane_op[j] = v
        for j, k in enumerate(L.EPISODE_SUM_NAMES):
            c.reward_scale[j] = R.reward_scales.get(k, 0.0)  # a zero scale drops the term, as legged_robot.py:936-938 does
        for j, v in enumerate(R.noise_scale_vec):
            c.noise_scale_vec[j] = v
        s = R.obs_scales
        c.obs_scale_lin_vel, c.obs_scale_ang_vel, c.obs_scale_dof_pos = s["lin_vel"], s["ang_vel"], s["dof_pos"]
        c.obs_scale_dof_vel, c.obs_scale_height, c.obs_scale_force = s["dof_vel"], s["height_measurements"], s["force"]
        c.clip_obs, c.clip_actions = R.clip_obs, R.clip_actions
        self._cfg_struct = c
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            B.check(self.lib.dtc_env_create(C.byref(c), C.byref(h)), "dtc_env_create")
        self._h = h
        self._bind()

    def _bind(self):
        t = dict(
            root_states=self.root_states, dof_state=self.dof_state, contact_forces=self.contact_forces,
            rigid_body_state=self.rigid_body_state, height_samples=self.height_samples, actions=self.actions,
            torques=self.torques, lag_buffer=self._lag, base_lin_vel=self.base_lin_vel, base_vel_scaled=self._base_vel,
            base_ang_vel=self.base_ang_vel,
            projected_gravity=self.projected_gravity, commands=self.commands, cmd_buffer=self.cmd_buffer,
            lin_vel_buffer=self.lin_vel_buffer, ang_vel_buffer=self.ang_vel_buffer, measured_heights=self.measured_heights,
            pred_footholds=self.pred_footholds, optimal_idx=self._optimal_idx, nominal_idx=self._nominal_idx,
            foothold_obs=self.foothold_obs, optimal_footholds_world=self.optimal_footholds_world,
            center_clear_mean=self._center_clear_mean, plane_ab=self._plane_ab, foot_clearance=self.measured_foot_clearance,
            contact_filt=self._contact_filt, last_contacts=self._last_contacts, stumb_buffer=self._stumb,
            feet_air_time=self.feet_air_time, pitch_est=self.pitch_est, last_actions=self.last_actions,
            last_actions_2=self.last_actions_2, last_dof_vel=self.last_dof_vel, last_root_vel=self.last_root_vel,
            last_foot_vel=self.last_foot_velocities, motor_strengths=self.motor_strengths, robot_mass=self.robot_mass,
            height_noise_offset=self._height_noise_offset, forces0=self._forces0, episode_length_buf=self.episode_length_buf,
            terrain_levels=self.terrain_levels, terrain_types=self.terrain_types, env_origins=self.env_origins,
            terrain_origins=self.terrain_origins, reset_buf=self.reset_buf, time_out_buf=self.time_out_buf,
            rew_buf=self.rew_buf, episode_sums=self._episode_sums, reward_terms=self._reward_terms, obs_buf=self.obs_buf,
            privileged_obs_buf=self._priv_store, obs_history=self._hist_store, episode_stats=self._episode_stats,
            episode_stats_last=self._episode_stats_last, time_outs_sent=self._time_outs_sent)
        b = B.EnvBuffers()
        for name in B.ENV_BUFFER_NAMES:
            x = t[name]
            assert x.is_contiguous() and x.is_cuda, name
            setattr(b, name, x.data_ptr())
        b.priv_ld, b.hist_ld = self.priv_ld, self.hist_ld
        self._keep = t
        B.check(self.lib.dtc_env_bind(self._h, C.byref(b)), "dtc_env_bind")

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.dtc_env_destroy(self._h)
        except Exception:
            pass

    # ------------------------------------------------------------------ reference-shaped views
    @property
    def optimal_foothold_indice(self):
        return self._optimal_idx.long().unsqueeze(1)  # [N,1,4] int64 like torch.topk's indices

    @property
    def nominal_footholds_indice(self):
        return self._nominal_idx.long()

    @property
    def contact_filt(self):
        return self._contact_filt.bool()

    @property
    def last_contacts(self):
        return self._last_contacts.bool()

    @property
    def height_noise_offset(self):
        return self._height_noise_offset.unsqueeze(1).expand(self.num_envs, L.NUM_POINTS)

    @property
    def obs_history(self):
        return self._hist_store[:, :L.NUM_OBS_HIST]

    @property
    def foot_positions(self):
        return self.rigid_body_state.view(self.num_envs, L.NUM_BODIES, 13)[:, L.FEET_INDICES, 0:3]

    @property
    def foot_velocities(self):
        return self.rigid_body_state.view(self.num_envs, L.NUM_BODIES, 13)[:, L.FEET_INDICES, 7:10]

    # ------------------------------------------------------------------ VecEnv surface
    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def get_reward_buf(self):
        return self.rew_buf

    def get_base_vel(self):
        return self._base_vel  # base_lin_vel * obs_scales.lin_vel, written by dtc_env_state_prep (legged_robot.py:1429-1431)

    def reset(self):
        """base_task.py:115-119: reset every robot, then one zero-action step."""
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, priv, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device))
        return obs, priv

    def reset_idx(self, env_ids):
        """legged_robot.py:200-272 for an explicit id list (host-driven: reset(), user scripts).  In-episode resets inside
        step() happen on the device in dtc_env_reward_reset (no nonzero()/len() host sync); this path
        does the same arithmetic with torch ops on the selected rows."""
        env_ids = torch.as_tensor(env_ids, device=self.device, dtype=torch.long).flatten()
        if env_ids.numel() == 0:
            return
        self._reset_ids(env_ids, self._host_draws or {})

    def _noise_struct(self):
        nz = B.EnvNoise()
        if self._noise:
            for k, v in self._noise.items():
                if v is not None:
                    B.require_cuda(v, k)
                    assert v.is_contiguous() and v.dtype == torch.float32
                    setattr(nz, k, v.data_ptr())
        return nz

    def step(self, actions):
        """legged_robot.py:92-122."""
        lib, st = self.lib, B.stream_ptr(self.device)
        B.require_cuda(actions, "actions")
        actions = actions.contiguous().float()
        hd = self._host_draws or {}
        lag = hd.get("lag") or ([int(self.np_rng.integers(1, 5)) for _ in range(L.DECIMATION)] if self.host_rng else [0] * L.DECIMATION)
        arr = (C.c_int32 * 4)(*lag)
        nstep, seed = self.common_step_counter + 1 - self._step_offset, self.seed
        gym, sim, uw = self.gym, self.sim, self._unwrap
        if getattr(gym, "static_dof_state", False):
            # stubbed simulator: the dof state does not move inside the decimation loop -> the four PD sub-steps in one launch
            B.check(lib.dtc_env_pre_physics(self._h, B.ptr(actions), arr, 0, L.DECIMATION, nstep, seed, st), "dtc_env_pre_physics")
            for _ in range(L.DECIMATION):
                gym.set_dof_actuation_force_tensor(sim, uw(self.torques))
                gym.simulate(sim)
                gym.refresh_dof_state_tensor(sim)
        else:
            # legged_robot.py:102-111: torque from the refreshed dof state in every sub-step, handed to the simulator each time
            for s in range(L.DECIMATION):
                B.check(lib.dtc_env_pre_physics(self._h, B.ptr(actions), arr, s, 1, nstep, seed, st), "dtc_env_pre_physics")
                gym.set_dof_actuation_force_tensor(sim, uw(self.torques))
                gym.simulate(sim)
                gym.refresh_dof_state_tensor(sim)
        self.post_physics_step()
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def post_physics_step(self):
        """legged_robot_dtc.py:56-223 in four launches."""
        lib, st = self.lib, B.stream_ptr(self.device)
        self.gym.refresh_actor_root_state_tensor(self.sim)
        self.gym.refresh_net_contact_force_tensor(self.sim)
        self.gym.refresh_rigid_body_state_tensor(self.sim)
        self.common_step_counter += 1
        nz = self._noise_struct()
        hd = self._host_draws or {}
        step, seed = self.common_step_counter - self._step_offset, self.seed
        B.check(lib.dtc_env_state_prep(self._h, step, seed, C.byref(nz), st), "dtc_env_state_prep")
        dbg = B.ptr(self._debug_score) if self._debug_score is not None else C.c_void_p(0)
        B.check(lib.dtc_foothold_step(self._h, self.foothold_variant, dbg, st), "dtc_foothold_step")
        rn = hd["reset_normal"] if "reset_normal" in hd else (float(self.np_rng.normal(0, 0.02)) if self.host_rng else math.nan)
        B.check(lib.dtc_env_reward_reset(self._h, step, seed, rn, C.byref(nz), st), "dtc_env_reward_reset")
        # pushes (legged_robot.py:673-678) and in-episode resets (:640-667) were written into root_states / dof_state on the device;
        # hand them to the simulator.  Which environments were reset is not known on the host (no nonzero() sync), so the whole
        # tensors go back: rows the kernel did not touch still hold what the simulator produced.
        self.gym.set_actor_root_state_tensor(self.sim, self._unwrap(self.root_states))
        self.gym.set_dof_state_tensor(self.sim, self._unwrap(self.dof_state))
        B.check(lib.dtc_env_observe(self._h, step, seed, C.byref(nz), st), "dtc_env_observe")

    def _reset_ids(self, ids, hd):
        """reset_idx (legged_robot.py:200-272) on the rows `ids`: init-time / user-script plumbing, not the hot path."""
        R, dev, n = self._rc, self.device, int(ids.numel())
        full = n == self.num_envs
        r = hd.get("reset0_u")
        if r is None:
            g = torch.Generator(device=dev).manual_seed(self.seed + 12345 + self.common_step_counter)
            r = torch.rand(n, 25, generator=g, device=dev)
        # terrain curriculum (:690-711)
        if R.terrain_curriculum:
            dxy = self.root_states[ids, :2] - self.env_origins[ids, :2]
            dist = torch.linalg.vector_norm(dxy, dim=1)
            up = dist > R.terrain_length * 0.6
            cn = torch.linalg.vector_norm(self.commands[ids, :2], dim=1)
            down = (dist < cn * R.episode_length_s * 0.5) & ~up
            lv = self.terrain_levels[ids] + up.long() - down.long()
            rnd = torch.clamp((r[:, 0] * R.num_rows).long(), max=R.num_rows - 1)
            lv = torch.where(lv >= self.max_terrain_level, rnd, torch.clip(lv, 0))
            self.terrain_levels[ids] = lv
            self.env_origins[ids] = self.terrain_origins[lv, self.terrain_types[ids]]
        # _reset_dofs (:640), _reset_root_states (dtc:299-311)
        dof = self.dof_state.view(self.num_envs, L.NUM_DOF, 2)
        dof[ids, :, 0] = self.default_dof_pos * (1.0 * r[:, 1:13] + 0.5)
        dof[ids, :, 1] = 0.0
        root = torch.tensor(R.base_init_state, device=dev).repeat(n, 1)
        root[:, :3] += self.env_origins[ids]
        root[:, :2] += 1.0 * r[:, 13:15] + -0.5
        root[:, 7:13] = 1.0 * r[:, 15:21] + -0.5
        self.root_states[ids] = root
        cmd = self.commands[ids]
        for col, key, k in ((0, "lin_vel_x", 21), (1, "lin_vel_y", 22), (3, "heading", 23)):
            lo, hi = R.command_ranges[key]
            cmd[:, col] = (hi - lo) * r[:, k] + lo
        cmd[:, :2] *= (torch.linalg.vector_norm(cmd[:, :2], dim=1) > 0.1).unsqueeze(1)
        self.commands[ids] = cmd
        self._forces0[ids] = 0.0
        lo, hi = R.motor_strength
        self.motor_strengths[ids] = (r[:, 24] * (hi - lo) + lo).unsqueeze(1).expand(n, 12)
        rn0 = hd["reset0_normal"] if "reset0_normal" in hd else float(self.np_rng.normal(0, 0.02))
        self._height_noise_offset[ids] = self._height_noise_offset[ids] * 0.0 + rn0
        for t in (self.last_actions, self.last_actions_2, self.last_dof_vel, self.feet_air_time, self.pitch_est, self._contact_filt,
                  self._last_contacts):
            t[ids] = 0
        for t in (self._lag, self._stumb, self.lin_vel_buffer, self.ang_vel_buffer, self.cmd_buffer):
            t[:, ids] = 0
        # extras of this reset_idx() call (legged_robot.py:253-264): episode means over the rows reset, current mean level
        sums = self._episode_sums[:, ids].sum(dim=1)
        self._episode_stats_last[:24] = sums
        self._episode_stats_last[24] = float(n)
        self._episode_stats_last[25:26].view(torch.int32)[0] = int(self.terrain_levels.sum().item())
        self._episode_sums[:, ids] = 0.0
        self._time_outs_sent.copy_(self.time_out_buf)
        self.episode_length_buf[ids] = 0
        self.reset_buf[ids] = 1
        # hand the new state to the simulator (legged_robot.py:643-667)
        self.gym.set_actor_root_state_tensor(self.sim, self._unwrap(self.root_states))
        self.gym.set_dof_state_tensor(self.sim, self._unwrap(self.dof_state))

    # ------------------------------------------------------------------ CUDA-graph capture of step() sequences
    def graph_supported(self):
        """step() can be captured when nothing in it depends on the host: device-side draws, no injected tables."""
        return not self.host_rng and self._noise is None and self._host_draws is None and self._debug_score is None

    def graph_capture_begin(self):
        """Call before capturing a sequence of step() calls: from here on the launches carry steps relative to the device counter."""
        self._step_offset = self.common_step_counter
        self._step_base.fill_(self._step_offset)

    def graph_capture_end(self, steps):
        """Call as the LAST captured operation: the graph itself advances the device counter by the steps it contains."""
        B.check(self.lib.dtc_counter_add(B.ptr(self._step_base), int(steps), B.stream_ptr(self.device)), "dtc_counter_add")
        self._step_offset += int(steps)

    def graph_replayed(self, steps):
        """Host mirror of one replay of a captured `steps`-step sequence."""
        self.common_step_counter += int(steps)
        self._step_offset += int(steps)

    def graph_before_replay(self):
        """Simulator hook: whatever the captured refresh_*() calls read has to be in place before the replay is queued."""
        f = getattr(self.gym, "graph_before_replay", None)
        if f is not None:
            f()

    def graph_after_replay(self):
        f = getattr(self.gym, "graph_after_replay", None)
        if f is not None:
            f()

    # The reference calls these three inside post_physics_step (legged_robot_dtc.py:205-213).  Here they are fused into
    # dtc_env_reward_reset / dtc_env_observe, so a call returns what the fused launch of the CURRENT step produced; a subclass
    # that wants different arithmetic has to supply a kernel, not a Python override.
    def check_termination(self):
        """reset_buf / time_out_buf of the current step (legged_robot_dtc.py:229-245)."""
        return self.reset_buf

    def compute_reward(self):
        """rew_buf of the current step; per-term values in `_reward_terms`, running sums in `episode_sums` (legged_robot.py:274-291)."""
        return self.rew_buf

    def compute_observations(self):
        """obs_buf / privileged_obs_buf of the current step (legged_robot_dtc.py:254-287)."""
        return self.obs_buf, self.privileged_obs_buf


class _Extras(dict):
    """`extras` of step().  In the reference this is ONE dict that lives as long as the environment and is rewritten only by a
    reset_idx() call with a non-empty id list (legged_robot.py:210,253-264): "time_outs" [N] bool and "episode" (means of the
    episode sums over the environments of that reset, `terrain_level`) therefore describe the last step that reset anything.
    The kernels keep exactly that on the device (dtc_env_buffers.time_outs_sent / episode_stats_last); "episode" needs a
    device->host read, so it is materialised only when a consumer looks - the training loop's logger - keeping step() sync-free."""

    def __init__(self, env):
        super().__init__()
        self._env = env
        self["time_outs"] = env._time_outs_sent.view(torch.bool)  # zero-copy view of the device flags

    def _episode(self):
        s = self._env._episode_stats_last.tolist()
        if s[24] <= 0:
            return None
        ep = {"rew_" + k: torch.tensor(s[i] / s[24] / self._env.max_episode_length_s) for i, k in enumerate(L.EPISODE_SUM_NAMES)
              if k in self._env.reward_scales}
        level_sum = int(self._env._episode_stats_last[25:26].view(torch.int32).item())
        ep["terrain_level"] = torch.tensor(level_sum / self._env.num_envs)
        return ep

    def __contains__(self, k):
        if k == "episode":
            return self._episode() is not None
        return dict.__contains__(self, k)

    def __getitem__(self, k):
        if k == "episode":
            ep = self._episode()
            if ep is None:
                raise KeyError(k)
            return ep
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        return self[k] if k in self else default
