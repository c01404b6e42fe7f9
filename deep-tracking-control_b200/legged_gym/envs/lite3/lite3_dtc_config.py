"""Lite3 DTC task configuration as nested classes with the reference's attribute names
(legged_gym/envs/lite3/lite3_dtc_config.py:3-195, inheriting legged_robot_config.py).  Only the fields the hot
path reads are present; values are shared with `dtc_b200.lite3`.  `LeggedRobotDTC` resolves whatever object it is given
through `envs/base/cfg_resolve.py` - an instance of the reference's own `Lite3DTCCfg` works as well - and edits to reward
scales, command ranges, gains, noise or randomisation settings reach the kernels."""
from .... import lite3 as L
from ..base.cfg_resolve import class_to_dict  # noqa: F401  (legged_gym/utils/helpers.py:11-26)


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
        static_friction = 1.0
        dynamic_friction = 1.0
        restitution = 0.0
        num_height_points = L.GRID_X * L.GRID_Y
        selected = False
        terrain_kwargs = None
        # terrain types: [smooth slope, rough slope, stairs up, stairs down, discrete, stepping stones, gap, pit] (terrain.py:79-141)
        terrain_proportions = [0.0, 0.0, 0.2, 0.2, 0.2, 0.4]
        slope_treshold = 0.75

    class commands:
        curriculum = False
        num_commands = 4
        resampling_time = 10.0
        heading_command = True

        class ranges:
            lin_vel_x = list(L.CMD_RANGES["lin_vel_x"])
            lin_vel_y = list(L.CMD_RANGES["lin_vel_y"])
            ang_vel_yaw = list(L.CMD_RANGES["ang_vel_yaw"])
            heading = list(L.CMD_RANGES["heading"])

    class init_state:
        pos = list(L.BASE_INIT_STATE[0:3])
        rot = list(L.BASE_INIT_STATE[3:7])
        lin_vel = [0.0, 0.0, 0.0]
        ang_vel = [0.0, 0.0, 0.0]
        default_joint_angles = dict(zip(L.DOF_NAMES, L.DEFAULT_DOF_POS))

    class domain_rand:
        push_robots = True
        push_interval_s = 15
        max_push_vel_xy = L.MAX_PUSH_VEL_XY
        max_push_force_xy = 0.0
        randomize_motor_strength = True
        motor_strength = list(L.MOTOR_STRENGTH_RANGE)
        randomize_Kp_factor = False
        randomize_Kd_factor = False

    class asset:
        penalize_contacts_on = ["TORSO", "THIGH", "SHANK"]
        terminate_after_contacts_on = []

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
        soft_dof_pos_limit = L.SOFT_DOF_POS_LIMIT

        class scales:
            pass

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
        resume = False
        load_run = -1    # -1 = last run
        checkpoint = -1  # -1 = last saved model


for _k, _v in L.REWARD_SCALES.items():  # reward scales as class attributes, like the reference's `class scales`
    setattr(Lite3DTCCfg.rewards.scales, _k, _v)
