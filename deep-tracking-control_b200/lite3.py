"""Constants of the Lite3 DTC task, as *data*.

Every number here restates a value pinned by the reference configuration
(`legged_gym/envs/lite3/lite3_dtc_config.py:3-195`, `legged_gym/envs/base/legged_robot_config.py`)
or by the robot asset (`resources/robots/Lite3/urdf/Lite3.urdf:58,87,116` joint limits; body/DOF order
TORSO,{FL,FR,HL,HR}x{HIP,THIGH,SHANK,FOOT}).  The reference obtains the asset-derived values from
Isaac Gym at scene-construction time (`legged_robot_dtc.py:318-457`); with the simulator stubbed
they are constants.
"""
import math

NUM_BODIES = 17
NUM_DOF = 12
NUM_ACTIONS = 12
NUM_OBS = 53
NUM_PRIV = 1389
NUM_HIST = 5
NUM_OBS_HIST = NUM_OBS * NUM_HIST  # 265
GRID_X, GRID_Y = 33, 21
NUM_POINTS = GRID_X * GRID_Y  # 693

FEET_INDICES = [4, 8, 12, 16]
THIGH_INDICES = [2, 6, 10, 14]
HIP_DOF_INDICES = [0, 3, 6, 9]
PENALISED_CONTACT_INDICES = [0, 2, 6, 10, 14, 3, 7, 11, 15]  # TORSO, THIGH x4, SHANK x4 (name-match order)
TERMINATION_CONTACT_INDICES = []

DOF_NAMES = [f"{leg}_{j}_joint" for leg in ("FL", "FR", "HL", "HR") for j in ("HipX", "HipY", "Knee")]
DEFAULT_DOF_POS = [0.1, -1.0, 1.8, -0.1, -1.0, 1.8, 0.1, -1.0, 1.8, -0.1, -1.0, 1.8]
_URDF_LIMITS = [(-0.523, 0.523), (-2.67, 0.314), (0.524, 2.792)] * 4
SOFT_DOF_POS_LIMIT = 0.9
DOF_VEL_LIMIT = 20.0
TORQUE_LIMIT = 0.8 * 30.0  # legged_robot.py:503
P_GAIN, D_GAIN = 25.0, 0.5
ACTION_SCALE = 0.25
DECIMATION = 4
SIM_DT = 0.005
DT = SIM_DT * DECIMATION  # 0.02
BASE_INIT_STATE = [0.0, 0.0, 0.4, 0.0, 0.0, 0.0, 1.0, 0, 0, 0, 0, 0, 0]

# terrain (lite3_dtc_config.py:20-51)
HORIZONTAL_SCALE = 0.05
VERTICAL_SCALE = 0.005
BORDER_SIZE = 20.0
TERRAIN_LENGTH = 8.0
NUM_ROWS, NUM_COLS = 6, 2
MAP_ROWS = int(NUM_ROWS * TERRAIN_LENGTH / HORIZONTAL_SCALE) + 2 * int(BORDER_SIZE / HORIZONTAL_SCALE)  # 1760
MAP_COLS = int(NUM_COLS * TERRAIN_LENGTH / HORIZONTAL_SCALE) + 2 * int(BORDER_SIZE / HORIZONTAL_SCALE)  # 1120
MEASURED_POINTS_X = [round(-0.8 + 0.05 * i, 2) for i in range(GRID_X)]
MEASURED_POINTS_Y = [round(-0.5 + 0.05 * i, 2) for i in range(GRID_Y)]

EPISODE_LENGTH_S = 20.0
MAX_EPISODE_LENGTH = math.ceil(EPISODE_LENGTH_S / DT)  # 1000
RESAMPLING_STEPS = int(10.0 / DT)  # 500
PUSH_INTERVAL = math.ceil(15.0 / DT)  # 750
MAX_PUSH_VEL_XY = 1.0
CMD_RANGES = {"lin_vel_x": (-0.75, 0.75), "lin_vel_y": (-0.75, 0.75), "ang_vel_yaw": (-0.5, 0.5),
              "heading": (-3.14, 3.14)}
MOTOR_STRENGTH_RANGE = (0.9, 1.1)
BASE_HEIGHT_TARGET = 0.32
TRACKING_SIGMA = 0.25
MAX_ACC = 100.0

OBS_SCALES = {"lin_vel": 2.0, "ang_vel": 0.25, "dof_pos": 1.0, "dof_vel": 0.05,
              "height_measurements": 5.0, "force": 0.005}
NOISE_SCALES = {"dof_pos": 0.01, "dof_vel": 1.5, "lin_vel": 0.1, "ang_vel": 0.2, "gravity": 0.05,
                "height_measurements": 0.1}
CLIP_OBS = 100.0
CLIP_ACTIONS = 100.0

# non-zero reward scales (lite3_dtc_config.py:141-181), BEFORE the *dt of legged_robot.py:939.
# Evaluation order is alphabetical (helpers.py:11-26 iterates dir()); "termination" is applied last.
REWARD_SCALES = {
    "action_rate": -0.01, "ang_vel_xy": -0.05 / 5, "base_height": -4.0, "collision": -1.5,
    "dof_acc": -2.5e-7 / 10, "dof_pos_limits": -10.0, "feet_air_time": 1.0, "feet_slip": -0.05,
    "foot_acc": -0.007, "foot_clearance": -0.01, "foothold_miss": -0.05, "hip_pos": -0.4 / 10,
    "lin_vel_z": -2.0 / 2, "orientation": -0.5, "pos_acc": -0.005, "power": -6e-7,
    "powerchange": -0.01 / 2, "smooth": -0.015 / 5, "soft_tracking_ang_vel": 0.5,
    "soft_tracking_lin_vel": 2, "stand_still": -0.2, "termination": -0.1, "torques": -0.000001,
    "tracking_optimal_footholds": 1,
}
REWARD_NAMES = sorted(k for k in REWARD_SCALES if k != "termination")  # 23 names
EPISODE_SUM_NAMES = sorted(REWARD_SCALES)  # 24 (termination included, alphabetical position)


def soft_dof_pos_limits():
    out = []
    for lo, hi in _URDF_LIMITS:
        m = (lo + hi) / 2
        r = hi - lo
        out.append((m - 0.5 * r * SOFT_DOF_POS_LIMIT, m + 0.5 * r * SOFT_DOF_POS_LIMIT))
    return out
