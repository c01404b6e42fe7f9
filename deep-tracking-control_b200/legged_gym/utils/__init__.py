from .helpers import class_to_dict, get_load_path, make_alg_runner  # noqa: F401
from .terrain import Terrain  # noqa: F401
