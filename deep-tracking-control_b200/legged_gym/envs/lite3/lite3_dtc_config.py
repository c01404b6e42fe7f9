"""Lite3 DTC task configuration as nested classes with the reference's attribute names
(legged_gym/envs/lite3/lite3_dtc_config.py:3-195, inheriting legged_robot_config.py).  Only the fields the hot
path reads are present; values are shared with `dtc_b200.lite3`."""
from .... import lite3 as L


class Lite3DTCCfg:
    class env:
        num_envs = 4096
        num_observations = L.NUM_OBS
        num_privileged_obs = L.NUM_PRIV
        num_obs_history = L.NUM_OBS_HIST
        num_observation_history = L.NUM_HIST
        num_actions = L.NUM_ACTIONS
        episode_length_s = L.EPISODE_LENGTH_S
        send_timeouts = True

    class terrain:
        mesh_type = "trimesh"
        horizontal_scale = L.HORIZONTAL_SCALE
        vertical_scale = L.VERTICAL_SCALE
        border_size = L.BORDER_SIZE
        curriculum = True
        measure_heights = True
        measure_foot_clearance = True
        measured_points_x = L.MEASURED_POINTS_X
        measured_points_y = L.MEASURED_POINTS_Y
        measured_x_dim, measured_y_dim = L.GRID_X, L.GRID_Y
        num_rows, num_cols = L.NUM_ROWS, L.NUM_COLS
        terrain_length = terrain_width = L.TERRAIN_LENGTH
        max_init_terrain_level = 5

    class control:
        control_type = "P"
        stiffness = {"joint": L.P_GAIN}
        damping = {"joint": L.D_GAIN}
        action_scale = L.ACTION_SCALE
        decimation = L.DECIMATION

    class sim:
        dt = L.SIM_DT

    class rewards:
        base_height_target = L.BASE_HEIGHT_TARGET
        tracking_sigma = L.TRACKING_SIGMA
        max_acc = L.MAX_ACC
        only_positive_rewards = False
        scales = dict(L.REWARD_SCALES)

    class normalization:
        obs_scales = dict(L.OBS_SCALES)
        clip_observations = L.CLIP_OBS
        clip_actions = L.CLIP_ACTIONS

    class noise:
        add_noise = True
        noise_level = 1.0
        noise_scales = dict(L.NOISE_SCALES)


class Lite3DTCCfgPPO:
    seed = 1
    runner_class_name = "OnPolicyRunner"

    class policy:
        init_noise_std = 1.0

    class algorithm:
        value_loss_coef = 1.0
        use_clipped_value_loss = True
        clip_param = 0.2
        entropy_coef = 0.003
        num_learning_epochs = 5
        num_mini_batches = 4
        learning_rate = 1.e-3
        schedule = "adaptive"
        gamma = 0.99
        lam = 0.95
        desired_kl = 0.01
        max_grad_norm = 1.0

    class runner:
        policy_class_name = "ActorCriticDecoder"
        algorithm_class_name = "PPO"
        num_steps_per_env = 24
        max_iterations = 20000
        save_interval = 50
        experiment_name = "lite3_dtc_highres"
        run_name = ""


def class_to_dict(obj):
    """legged_gym/utils/helpers.py:11-26 (alphabetical dir() order)."""
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
