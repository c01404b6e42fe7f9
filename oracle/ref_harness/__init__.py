"""TEST INFRASTRUCTURE ONLY - never imported by the product package.

Runs the UNMODIFIED reference (/root/reference, present only in the build container) on CPU through
three import stubs (`stubs/`: isaacgym, gym, params_proto) and a FakeGym, following the recipe of
SURVEY.md Appendix A.  Used by tests/golden/make_golden.py to record golden vectors that pin the
restated oracle (oracle/*.py); `/root/reference` does not exist on the GPU box, so nothing that runs
there imports this module.
"""
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "legged_gym"))


def import_reference():
    """Make `legged_gym` / `rsl_rl` importable from the read-only reference tree."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected in the build container only)")
    for p in (os.path.join(REFERENCE_ROOT, "rsl_rl"), REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import legged_gym.envs  # noqa: F401  (registers tasks)
    return sys.modules["legged_gym"]


def build_ref_env(num_envs, height_samples, terrain_origins, layout, fake_gym, device="cpu"):
    """LeggedRobotDTC over FakeGym without Isaac Gym scene construction.

    layout = (terrain_levels, terrain_types, env_origins, terrain_origins_t) from sim_stub.initial_env_layout
    """
    import_reference()
    from legged_gym.envs.base.legged_robot_dtc import LeggedRobotDTC
    from legged_gym.envs.lite3.lite3_dtc_config import Lite3DTCCfg

    cfg = Lite3DTCCfg()
    cfg.env.num_envs = num_envs
    env = object.__new__(LeggedRobotDTC)
    env.cfg = cfg
    env.sim_params = types.SimpleNamespace(dt=cfg.sim.dt, use_gpu_pipeline=False)
    env.height_samples = None
    env.debug_viz = False
    env.init_done = False
    env._parse_cfg(cfg)
    # BaseTask.__init__ (base_task.py:11-52) minus create_sim
    env.gym = fake_gym
    env.sim = None
    env.device = device
    env.headless = True
    env.viewer = None
    env.enable_viewer_sync = True
    env.num_envs = num_envs
    env.num_obs = cfg.env.num_observations
    env.num_privileged_obs = cfg.env.num_privileged_obs
    env.num_actions = cfg.env.num_actions
    env.obs_buf = torch.zeros(num_envs, env.num_obs, device=device)
    env.rew_buf = torch.zeros(num_envs, device=device)
    env.reset_buf = torch.ones(num_envs, device=device, dtype=torch.long)
    env.episode_length_buf = torch.zeros(num_envs, device=device, dtype=torch.long)
    env.time_out_buf = torch.zeros(num_envs, device=device, dtype=torch.bool)
    env.privileged_obs_buf = torch.zeros(num_envs, env.num_privileged_obs, device=device)
    env.extras = {}
    # products of _create_envs (legged_robot_dtc.py:318-457), constant for the Lite3 asset
    from dtc_b200 import lite3 as L
    env.up_axis_idx = 2
    env.num_bodies = L.NUM_BODIES
    env.num_dof = env.num_dofs = L.NUM_DOF
    env.dof_names = list(L.DOF_NAMES)
    env.feet_indices = torch.tensor(L.FEET_INDICES, dtype=torch.long, device=device)
    env.thigh_indices = torch.tensor(L.THIGH_INDICES, dtype=torch.long, device=device)
    env.hip_indices = torch.tensor(L.HIP_DOF_INDICES, dtype=torch.long, device=device)
    env.penalised_contact_indices = torch.tensor(L.PENALISED_CONTACT_INDICES, dtype=torch.long, device=device)
    env.termination_contact_indices = torch.tensor(L.TERMINATION_CONTACT_INDICES, dtype=torch.long, device=device)
    env.collision_contact_indices = env.penalised_contact_indices.clone()
    env.dof_pos_limits = torch.tensor(L.soft_dof_pos_limits(), dtype=torch.float, device=device)
    env.dof_vel_limits = torch.full((L.NUM_DOF,), L.DOF_VEL_LIMIT, device=device)
    env.torque_limits = torch.full((L.NUM_DOF,), L.TORQUE_LIMIT, device=device)
    env.base_init_state = torch.tensor(L.BASE_INIT_STATE, dtype=torch.float, device=device)
    env.default_friction = 1.0
    env.default_restitution = 0.0
    env.height_samples = torch.as_tensor(np.asarray(height_samples)).view(L.MAP_ROWS, L.MAP_COLS).to(device)
    env.terrain = types.SimpleNamespace(cfg=cfg.terrain, env_length=cfg.terrain.terrain_length,
                                        env_origins=np.asarray(terrain_origins))
    levels, ttypes, origins, tor = layout
    env.custom_origins = True
    env.terrain_levels = levels.clone().to(device)
    env.terrain_types = ttypes.clone().to(device)
    env.max_terrain_level = cfg.terrain.num_rows
    env.terrain_origins = tor.clone().to(device)
    env.env_origins = origins.clone().to(device)
    env._init_custom_buffers__()
    env.robot_mass = torch.full((num_envs,), 12.0, device=device)
    env._init_buffers()
    env._prepare_reward_function()
    env.init_done = True
    env.foothold_obs = torch.zeros(num_envs, 8, device=device)
    env.lidar_global_counter = 0
    env.height_global_counter = 0
    env.lookat_id = 0
    env.render = lambda *a, **k: None
    return env
