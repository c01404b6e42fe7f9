"""ctypes binding of libdtc_b200.so (include/dtc_b200.h).  The product path has no CPU fallback: if the
library is missing or no CUDA device is present, the ops raise."""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libdtc_b200.so")
_lib = None

f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class EnvConfig(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("map_rows", C.c_int32), ("map_cols", C.c_int32),
        ("horizontal_scale", C.c_float), ("vertical_scale", C.c_float), ("border_size", C.c_float),
        ("dt", C.c_float), ("max_episode_length", C.c_int32), ("resampling_steps", C.c_int32),
        ("push_interval", C.c_int32), ("max_push_vel_xy", C.c_float),
        ("cmd_lin_x", C.c_float * 2), ("cmd_lin_y", C.c_float * 2), ("cmd_heading", C.c_float * 2),
        ("motor_strength", C.c_float * 2), ("cmd_lin_x_max", C.c_float), ("cmd_ang_yaw_max", C.c_float),
        ("base_height_target", C.c_float), ("tracking_sigma", C.c_float), ("max_acc", C.c_float),
        ("terrain_length", C.c_float), ("max_terrain_level", C.c_int32), ("num_terrain_cols", C.c_int32),
        ("episode_length_s", C.c_float),
        ("p_gains", C.c_float * 12), ("d_gains", C.c_float * 12), ("action_scale", C.c_float), ("torque_limit", C.c_float),
        ("terrain_curriculum", C.c_int32), ("push_robots", C.c_int32),
        ("default_dof_pos", C.c_float * 12), ("dof_pos_lower", C.c_float * 12), ("dof_pos_upper", C.c_float * 12),
        ("base_init_state", C.c_float * 13), ("grid_x", C.c_float * 33), ("grid_y", C.c_float * 21),
        ("plane_op", C.c_float * (2 * 693)), ("reward_scale", C.c_float * 24), ("noise_scale_vec", C.c_float * 53),
        ("obs_scale_lin_vel", C.c_float), ("obs_scale_ang_vel", C.c_float), ("obs_scale_dof_pos", C.c_float),
        ("obs_scale_dof_vel", C.c_float), ("obs_scale_height", C.c_float), ("obs_scale_force", C.c_float),
        ("clip_obs", C.c_float), ("clip_actions", C.c_float),
    ]


ENV_BUFFER_NAMES = [
    "root_states", "dof_state", "contact_forces", "rigid_body_state", "height_samples",
    "actions", "torques", "lag_buffer", "base_lin_vel", "base_vel_scaled", "base_ang_vel", "projected_gravity", "commands",
    "cmd_buffer", "lin_vel_buffer", "ang_vel_buffer", "measured_heights", "pred_footholds", "optimal_idx",
    "nominal_idx", "foothold_obs", "optimal_footholds_world", "center_clear_mean", "plane_ab", "foot_clearance",
    "contact_filt", "last_contacts", "stumb_buffer", "feet_air_time", "pitch_est", "last_actions", "last_actions_2",
    "last_dof_vel", "last_root_vel", "last_foot_vel", "motor_strengths", "robot_mass", "height_noise_offset",
    "forces0", "episode_length_buf", "terrain_levels", "terrain_types", "env_origins", "terrain_origins",
    "reset_buf", "time_out_buf", "rew_buf", "episode_sums", "reward_terms", "obs_buf", "privileged_obs_buf",
    "obs_history", "episode_stats", "episode_stats_last", "time_outs_sent",
]


class EnvBuffers(C.Structure):
    _fields_ = [(n, vp) for n in ENV_BUFFER_NAMES] + [("priv_ld", C.c_int32), ("hist_ld", C.c_int32)]


class EnvNoise(C.Structure):
    _fields_ = [(n, vp) for n in ("resample_u", "push_u", "reset_u", "priv_u", "obs_u")]


STORAGE_PTRS = ["hist", "priv_a", "xc", "next_obs", "actions", "mu", "sigma", "rewards", "values", "returns",
                "advantages", "logp", "dones", "hist_lo", "priv_a_lo", "xc_lo"]


class Storage(C.Structure):
    _fields_ = [(n, vp) for n in STORAGE_PTRS] + [("T", C.c_int32), ("N", C.c_int32)]


class PPOHParams(C.Structure):
    _fields_ = [("clip_param", C.c_float), ("value_loss_coef", C.c_float), ("entropy_coef", C.c_float),
                ("max_grad_norm", C.c_float), ("desired_kl", C.c_float), ("adaptive_lr", C.c_int32),
                ("use_clipped_value_loss", C.c_int32), ("reserved", C.c_int32)]


class ParamInfo(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32), ("ld", C.c_int32),
                ("nseg", C.c_int32), ("seg_src", C.c_int32 * 4), ("seg_dst", C.c_int32 * 4), ("seg_len", C.c_int32 * 4)]


# every entry point include/dtc_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
EXPORTED_SYMBOLS = ['dtc_env_bind', 'dtc_env_create', 'dtc_env_destroy', 'dtc_env_heightmap_updated', 'dtc_env_observe', 'dtc_env_pre_physics', 'dtc_env_reward_reset', 'dtc_env_set_step_base', 'dtc_env_state_prep', 'dtc_counter_add', 'dtc_count_launches', 'dtc_learner_set_act_counter_base', 'dtc_foothold_step', 'dtc_gae', 'dtc_gae_normalize', 'dtc_gather_minibatch', 'dtc_gemm_debug', 'dtc_get_gemm_mode', 'dtc_get_gemm_pair', 'dtc_get_overlap', 'dtc_gru_forward', 'dtc_gru_param_floats', 'dtc_gru_reset', 'dtc_last_error', 'dtc_launch_count', 'dtc_learner_create', 'dtc_learner_debug_buffer', 'dtc_learner_destroy', 'dtc_learner_get_adam_steps', 'dtc_learner_grad_bucket', 'dtc_learner_wait_bucket', 'dtc_learner_refresh_params', 'dtc_learner_reset_stats', 'dtc_learner_set_adam_steps', 'dtc_learner_set_lr', 'dtc_learner_stats', 'dtc_learner_workspace_bytes', 'dtc_linear_forward', 'dtc_optimizer_apply', 'dtc_param_count', 'dtc_param_get', 'dtc_param_range', 'dtc_param_total_floats', 'dtc_policy_act', 'dtc_policy_act_teacher', 'dtc_policy_evaluate', 'dtc_ppo_step', 'dtc_profile_enable', 'dtc_profile_kind', 'dtc_profile_read', 'dtc_set_gemm_mode', 'dtc_set_gemm_pair', 'dtc_set_overlap', 'dtc_store_transition', 'dtc_struct_size', 'dtc_terrain_paint', 'dtc_terrain_rasterize', 'dtc_vae_step', 'dtc_version', 'dtc_dp_create', 'dtc_dp_handles', 'dtc_dp_open', 'dtc_dp_allreduce', 'dtc_dp_error', 'dtc_dp_register', 'dtc_dp_open_registered',
                    'dtc_dp_destroy']


class DtcError(RuntimeError):
    pass


def lib():
    """Loads the library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DtcError(f"{LIB_PATH} not found - run `python __graft_entry__.py` (build()) first; there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.dtc_last_error.restype = C.c_char_p
    L.dtc_launch_count.restype = C.c_int64
    L.dtc_profile_enable.argtypes = [C.c_int]
    L.dtc_profile_enable.restype = None
    L.dtc_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.dtc_env_create.argtypes = [C.POINTER(EnvConfig), C.POINTER(vp)]
    L.dtc_env_destroy.argtypes = [vp]
    L.dtc_env_destroy.restype = None
    L.dtc_env_bind.argtypes = [vp, C.POINTER(EnvBuffers)]
    L.dtc_env_set_step_base.argtypes = [vp, vp]
    L.dtc_counter_add.argtypes = [vp, C.c_int64, vp]
    L.dtc_count_launches.argtypes = [C.c_int64]
    L.dtc_count_launches.restype = None
    L.dtc_learner_set_act_counter_base.argtypes = [vp, vp]
    L.dtc_env_pre_physics.argtypes = [vp, vp, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int64, C.c_uint64, vp]
    L.dtc_env_state_prep.argtypes = [vp, C.c_int64, C.c_uint64, C.POINTER(EnvNoise), vp]
    L.dtc_foothold_step.argtypes = [vp, C.c_int, vp, vp]
    L.dtc_env_reward_reset.argtypes = [vp, C.c_int64, C.c_uint64, C.c_float, C.POINTER(EnvNoise), vp]
    L.dtc_env_observe.argtypes = [vp, C.c_int64, C.c_uint64, C.POINTER(EnvNoise), vp]
    if not hasattr(L, "dtc_learner_create"):
        raise DtcError("libdtc_b200.so was built without the learner kernels - rebuild")
    L.dtc_param_total_floats.restype = C.c_int64
    L.dtc_learner_workspace_bytes.restype = C.c_int64
    L.dtc_learner_workspace_bytes.argtypes = [C.c_int32]
    L.dtc_learner_stats.restype = C.c_void_p
    L.dtc_learner_stats.argtypes = [vp]
    L.dtc_param_get.argtypes = [C.c_int, C.POINTER(ParamInfo)]
    L.dtc_param_range.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.dtc_learner_create.argtypes = [C.c_int32, vp, vp, vp, vp, vp, vp, vp, C.c_int64, vp, C.POINTER(vp)]
    L.dtc_learner_destroy.argtypes = [vp]
    L.dtc_learner_destroy.restype = None
    L.dtc_learner_refresh_params.argtypes = [vp, vp]
    L.dtc_learner_grad_bucket.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.dtc_learner_wait_bucket.argtypes = [vp, C.c_int, C.c_int, vp]
    i32, u64, SP = C.c_int32, C.c_uint64, C.POINTER(Storage)
    L.dtc_policy_act.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, u64, u64, SP, i32, vp, vp, vp, vp, vp, vp]
    L.dtc_policy_evaluate.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, vp, vp]
    L.dtc_policy_act_teacher.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32, vp, vp]
    L.dtc_store_transition.argtypes = [SP, i32, vp, vp, vp, vp, i32, C.c_float, vp]
    L.dtc_gae.argtypes = [SP, vp, C.c_float, C.c_float, vp, C.c_int, vp]
    L.dtc_terrain_rasterize.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, C.c_double, C.c_double, vp, vp, vp]
    L.dtc_terrain_paint.argtypes = [i32, i32, i32, i32, i32, i32, i32, vp, vp, C.POINTER(C.c_int32), C.c_double, C.c_double, C.c_double, vp, vp, vp]
    L.dtc_gru_param_floats.restype = C.c_int64
    L.dtc_gru_param_floats.argtypes = [i32, i32, i32]
    L.dtc_gru_forward.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.dtc_gru_reset.argtypes = [i32, i32, i32, vp, vp, vp]
    L.dtc_dp_create.argtypes = [i32, i32, C.c_int64, C.POINTER(vp)]
    L.dtc_dp_handles.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.dtc_dp_open.argtypes = [vp, i32, C.c_char_p, C.c_char_p]
    L.dtc_dp_allreduce.argtypes = [vp, vp, C.c_int64, vp]
    L.dtc_dp_register.argtypes = [vp, vp, C.c_int64, C.c_char_p, C.POINTER(C.c_int64)]
    L.dtc_dp_open_registered.argtypes = [vp, i32, C.c_char_p, C.c_int64]
    L.dtc_dp_error.argtypes = [vp, vp]
    L.dtc_dp_destroy.argtypes = [vp]
    L.dtc_dp_destroy.restype = None
    L.dtc_profile_kind.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.dtc_gae_normalize.argtypes = [SP, vp, vp]
    L.dtc_gather_minibatch.argtypes = [SP, SP, vp, C.c_int64, vp]
    step_args = [vp, SP, C.c_int64, i32, vp, u64, u64, C.POINTER(PPOHParams), C.c_int, vp]
    L.dtc_vae_step.argtypes = step_args
    L.dtc_ppo_step.argtypes = step_args
    L.dtc_optimizer_apply.argtypes = [vp, C.c_int, C.POINTER(PPOHParams), C.c_float, i32, vp]
    L.dtc_learner_set_lr.argtypes = [vp, C.c_double, vp]
    L.dtc_learner_reset_stats.argtypes = [vp, vp]
    L.dtc_learner_set_adam_steps.argtypes = [vp, C.c_int64, C.c_int64]
    L.dtc_learner_get_adam_steps.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.dtc_learner_debug_buffer.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.dtc_linear_forward.argtypes = [i32, i32, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp]
    L.dtc_gemm_debug.argtypes = [i32, i32, i32, vp, vp, i32, i32, vp, vp, i32, i32, vp, vp, i32, i32, vp, i32, vp]
    L.dtc_set_gemm_mode.argtypes = [C.c_int]
    L.dtc_set_gemm_mode.restype = None
    for which, st in enumerate((EnvConfig, EnvBuffers, EnvNoise, Storage, PPOHParams, ParamInfo)):
        got = L.dtc_struct_size(which)
        if got != C.sizeof(st):
            raise DtcError(f"ABI mismatch: {st.__name__} is {C.sizeof(st)} B in Python, {got} B in the library")
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise DtcError(f"{what} failed ({rc}): {lib().dtc_last_error().decode()}")


def require_cuda(t, name):
    if not t.is_cuda:
        raise DtcError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    return t


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(lib().dtc_launch_count())
