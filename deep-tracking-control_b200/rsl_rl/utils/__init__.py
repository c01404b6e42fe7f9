from . import dp  # noqa: F401
from .utils import split_and_pad_trajectories, unpad_trajectories  # noqa: F401
