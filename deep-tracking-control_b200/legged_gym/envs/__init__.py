from .base.legged_robot_dtc import LeggedRobotDTC  # noqa: F401
from .lite3.lite3_dtc_config import Lite3DTCCfg, Lite3DTCCfgPPO  # noqa: F401
