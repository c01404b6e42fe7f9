from . import dp  # noqa: F401
