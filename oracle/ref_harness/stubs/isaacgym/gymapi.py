"""Placeholders for the scene-construction API (never executed on the stub path)."""
SIM_PHYSX, SIM_FLEX = 1, 0
ENV_SPACE = 0
IMAGE_DEPTH = 1
FOLLOW_TRANSFORM = 1
for _k in "ESCAPE V F W S A D Q E R P N".split():
    globals()["KEY_" + _k] = _k


class _Bag:
    def __init__(self, *a, **k):
        self.__dict__.update(k)

    def __getattr__(self, name):  # permissive: any attribute reads as a fresh bag
        b = _Bag()
        object.__setattr__(self, name, b)
        return b


class Vec3(_Bag): pass
class Transform(_Bag): pass
class Quat(_Bag): pass
class AssetOptions(_Bag): pass
class SimParams(_Bag): pass
class PlaneParams(_Bag): pass
class HeightFieldParams(_Bag): pass
class TriangleMeshParams(_Bag): pass
class CameraProperties(_Bag): pass


def acquire_gym():
    raise RuntimeError("isaacgym stub: no simulator; use oracle.ref_harness.FakeGym")
