"""Import stub for NVIDIA Isaac Gym (closed source, absent from this image).

TEST INFRASTRUCTURE ONLY.  It lets the *unmodified* reference under /root/reference be imported
on CPU so tests/golden/make_golden.py can record golden vectors.  Nothing in the product
package imports this.  The arithmetic helpers in torch_utils restate the published BSD-licensed
formulas of isaacgymenvs/utils/torch_jit_utils.py ("parity unpinned" at this
third-party boundary - no reference test pins them).
"""
from . import gymapi, gymutil, gymtorch, terrain_utils, torch_utils  # noqa: F401
