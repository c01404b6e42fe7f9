"""Resolves a task configuration object into the numbers the environment kernels consume - the CUDA path's counterpart of
`LeggedRobot._parse_cfg`, `_prepare_reward_function`, `_get_noise_scale_vec` and the cfg reads of `_init_buffers`
(legged_gym/envs/base/legged_robot.py:1230-1240, 929-952, 729-752, 795-846).

Accepts the reference's own `Lite3DTCCfg` instance (nested classes, `rewards.scales` a class, `control.stiffness` a dict keyed by
a joint-name substring) as well as this package's trimmed copy; attributes a config does not carry fall back to the Lite3 DTC
defaults of `dtc_b200.lite3`.  Anything the fused kernels cannot honour raises `CfgError` instead of being ignored.
"""
import math

from .... import lite3 as L


class CfgError(ValueError):
    pass


def class_to_dict(obj):
    """legged_gym/utils/helpers.py:11-26 (alphabetical dir() order)."""
    if isinstance(obj, dict):
        return dict(obj)
    if not hasattr(obj, "__dict__") and not isinstance(obj, type):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        if isinstance(val, type):
            out[key] = class_to_dict(val)
        elif isinstance(val, list):
            out[key] = [class_to_dict(v) for v in val]
        else:
            out[key] = val
    return out


def _get(cfg, path, default):
    cur = cfg
    for name in path.split("."):
        if isinstance(cur, dict):
            if name not in cur:
                return default
            cur = cur[name]
        else:
            if not hasattr(cur, name):
                return default
            cur = getattr(cur, name)
    return cur


class Resolved:
    """Plain attribute bag; see resolve()."""


def resolve(cfg, sim_dt=None):
    r = Resolved()
    g = lambda path, default: _get(cfg, path, default)
    # ---- shapes the kernels are specialised for
    for path, want in (("env.num_observations", L.NUM_OBS), ("env.num_privileged_obs", L.NUM_PRIV), ("env.num_actions", L.NUM_ACTIONS),
                       ("env.num_observation_history", L.NUM_HIST)):
        got = g(path, want)
        if got != want:
            raise CfgError(f"{path} = {got}: the fused kernels are specialised for {want} (Lite3 / X30 DTC task)")
    r.num_envs = int(g("env.num_envs", 4096))
    r.decimation = int(g("control.decimation", L.DECIMATION))
    if r.decimation != 4:
        raise CfgError(f"control.decimation = {r.decimation}: the lag buffer / torque kernel implement the reference's 4 sub-steps")
    if g("control.control_type", "P") != "P":
        raise CfgError("control.control_type must be 'P' (position targets, legged_robot.py:612-614)")
    r.sim_dt = float(sim_dt if sim_dt is not None else g("sim.dt", L.SIM_DT))
    r.dt = r.decimation * r.sim_dt
    r.episode_length_s = float(g("env.episode_length_s", L.EPISODE_LENGTH_S))
    r.max_episode_length = math.ceil(r.episode_length_s / r.dt)
    # ---- terrain
    if g("terrain.mesh_type", "trimesh") not in ("heightfield", "trimesh"):
        raise CfgError("terrain.mesh_type must be 'heightfield' or 'trimesh': the foothold pipeline samples height_samples")
    if not g("terrain.measure_heights", True):
        raise CfgError("terrain.measure_heights must be True")
    r.horizontal_scale = float(g("terrain.horizontal_scale", L.HORIZONTAL_SCALE))
    r.vertical_scale = float(g("terrain.vertical_scale", L.VERTICAL_SCALE))
    r.border_size = float(g("terrain.border_size", L.BORDER_SIZE))
    r.grid_x = [float(v) for v in g("terrain.measured_points_x", L.MEASURED_POINTS_X)]
    r.grid_y = [float(v) for v in g("terrain.measured_points_y", L.MEASURED_POINTS_Y)]
    if len(r.grid_x) != L.GRID_X or len(r.grid_y) != L.GRID_Y:
        raise CfgError(f"terrain.measured_points_x / _y must have {L.GRID_X} / {L.GRID_Y} entries (693 height points)")
    r.terrain_length = float(g("terrain.terrain_length", L.TERRAIN_LENGTH))
    r.num_rows, r.num_cols = int(g("terrain.num_rows", L.NUM_ROWS)), int(g("terrain.num_cols", L.NUM_COLS))
    r.terrain_curriculum = bool(g("terrain.curriculum", True))
    # ---- commands
    if not g("commands.heading_command", True):
        raise CfgError("commands.heading_command must be True (the yaw command is recomputed from the heading error every step)")
    if g("commands.curriculum", False):
        raise CfgError("commands.curriculum is not supported (False in every DTC task)")
    r.resampling_steps = int(float(g("commands.resampling_time", 10.0)) / r.dt)
    ranges = class_to_dict(g("commands.ranges", {k: list(v) for k, v in L.CMD_RANGES.items()}))
    r.command_ranges = {k: [float(ranges[k][0]), float(ranges[k][1])] for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "heading")}
    # ---- control
    def gains(d, what):
        out = []
        for name in L.DOF_NAMES:
            hit = [v for k, v in d.items() if k in name]
            if not hit:
                raise CfgError(f"control.{what} has no entry matching joint {name} (legged_robot.py:1098-1109)")
            out.append(float(hit[0]))
        return out
    r.p_gains = gains(dict(g("control.stiffness", {"joint": L.P_GAIN})), "stiffness")
    r.d_gains = gains(dict(g("control.damping", {"joint": L.D_GAIN})), "damping")
    r.action_scale = float(g("control.action_scale", L.ACTION_SCALE))
    dja = g("init_state.default_joint_angles", None)
    r.default_dof_pos = [float(dja[n]) for n in L.DOF_NAMES] if dja else list(L.DEFAULT_DOF_POS)
    pos, rot = g("init_state.pos", L.BASE_INIT_STATE[0:3]), g("init_state.rot", L.BASE_INIT_STATE[3:7])
    lin, ang = g("init_state.lin_vel", [0.0, 0.0, 0.0]), g("init_state.ang_vel", [0.0, 0.0, 0.0])
    r.base_init_state = [float(v) for v in list(pos) + list(rot) + list(lin) + list(ang)]
    soft = float(g("rewards.soft_dof_pos_limit", L.SOFT_DOF_POS_LIMIT))
    r.dof_pos_limits = []
    for lo, hi in L._URDF_LIMITS:  # resources/robots/Lite3/urdf/Lite3.urdf:58,87,116 (legged_robot.py:494-500)
        m, rng = (lo + hi) / 2, hi - lo
        r.dof_pos_limits.append((m - 0.5 * rng * soft, m + 0.5 * rng * soft))
    r.torque_limit = L.TORQUE_LIMIT
    # ---- domain randomisation
    r.push_robots = bool(g("domain_rand.push_robots", True))
    r.push_interval = math.ceil(float(g("domain_rand.push_interval_s", 15.0)) / r.dt)
    r.max_push_vel_xy = float(g("domain_rand.max_push_vel_xy", L.MAX_PUSH_VEL_XY))
    if float(g("domain_rand.max_push_force_xy", 0.0)) != 0.0:
        raise CfgError("domain_rand.max_push_force_xy != 0 is not supported (0 in every DTC task)")
    ms = g("domain_rand.motor_strength", list(L.MOTOR_STRENGTH_RANGE))
    r.motor_strength = [float(ms[0]), float(ms[1])] if g("domain_rand.randomize_motor_strength", True) else [1.0, 1.0]
    for flag in ("randomize_Kp_factor", "randomize_Kd_factor"):
        if g("domain_rand." + flag, False):
            raise CfgError(f"domain_rand.{flag} is not supported")
    # ---- rewards (legged_robot.py:929-952): scale * dt per named term; zero scales drop out
    if g("rewards.only_positive_rewards", False):
        raise CfgError("rewards.only_positive_rewards = True is not supported by the fused reward kernel")
    scales = class_to_dict(g("rewards.scales", dict(L.REWARD_SCALES)))
    r.reward_scales = {}
    for name, v in scales.items():
        if v is None or float(v) == 0.0:
            continue
        if name not in L.EPISODE_SUM_NAMES:
            raise CfgError(f"rewards.scales.{name} = {v}: no fused implementation of _reward_{name} (available: {', '.join(L.EPISODE_SUM_NAMES)})")
        r.reward_scales[name] = float(v) * r.dt
    r.base_height_target = float(g("rewards.base_height_target", L.BASE_HEIGHT_TARGET))
    r.tracking_sigma = float(g("rewards.tracking_sigma", L.TRACKING_SIGMA))
    r.max_acc = float(g("rewards.max_acc", L.MAX_ACC))
    # ---- asset-derived sets the kernels hard-code
    if list(g("asset.terminate_after_contacts_on", [])):
        raise CfgError("asset.terminate_after_contacts_on must be empty (check_termination's contact term is compiled out)")
    pen = list(g("asset.penalize_contacts_on", ["TORSO", "THIGH", "SHANK"]))
    if sorted(pen) != ["SHANK", "THIGH", "TORSO"]:
        raise CfgError("asset.penalize_contacts_on must be ['TORSO', 'THIGH', 'SHANK'] (the collision reward's body set)")
    # ---- normalisation / noise (legged_robot.py:729-752)
    obs = class_to_dict(g("normalization.obs_scales", dict(L.OBS_SCALES)))
    r.obs_scales = {k: float(obs.get(k, L.OBS_SCALES[k])) for k in L.OBS_SCALES}
    r.clip_obs = float(g("normalization.clip_observations", L.CLIP_OBS))
    r.clip_actions = float(g("normalization.clip_actions", L.CLIP_ACTIONS))
    r.add_noise = bool(g("noise.add_noise", True))
    ns = class_to_dict(g("noise.noise_scales", dict(L.NOISE_SCALES)))
    lvl = float(g("noise.noise_level", 1.0))
    v = [0.0] * L.NUM_OBS
    if r.add_noise:
        v[0:3] = [float(ns["ang_vel"]) * lvl * r.obs_scales["ang_vel"]] * 3
        v[3:6] = [float(ns["gravity"]) * lvl] * 3
        v[9:21] = [float(ns["dof_pos"]) * lvl * r.obs_scales["dof_pos"]] * 12
        v[21:33] = [float(ns["dof_vel"]) * lvl * r.obs_scales["dof_vel"]] * 12
    r.noise_scale_vec = v
    return r
