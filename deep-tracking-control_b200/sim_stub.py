"""Stub for the Isaac Gym physics call (north_star: "the Isaac Gym sim.step() call is stubbed on the
benchmark path with recorded/synthetic root/dof/contact tensors").

This is host-side set-up and plumbing, not the hot path: it builds a synthetic int16 heightmap of the
reference's shape/dtype (`legged_gym/utils/terrain.py:26-30`: 6x2 sub-terrains of 160x160 px plus a
400 px border -> [1760, 1120]) and draws seeded per-step values for the four state tensors the
reference acquires from PhysX (`legged_robot.py:759-779`): root_states[N,13], dof_state[N*12,2],
net_contact_force[N*17,3], rigid_body_state[N*17,13].  Distributions follow SURVEY.md section 8(d).

`FakeGym` exposes the handful of `gym.*` entry points the reference env touches outside scene
construction, so the same object drives the reference (CPU oracle runs) and this package.
"""
import math

import numpy as np
import torch

from . import lite3 as L


# ----------------------------------------------------------------------------- heightmaps
def _sub_origin(i, j):
    b = int(L.BORDER_SIZE / L.HORIZONTAL_SCALE)
    px = int(L.TERRAIN_LENGTH / L.HORIZONTAL_SCALE)
    return b + i * px, b + j * px, px


def _stepping_stones(rng, px, stone_size_m, gap_m, depth_m=-2.0, platform_m=1.0):
    """Stones of side ~stone_size at height 0 separated by gaps, pit at `depth` (terrain.py:133)."""
    h = np.full((px, px), int(depth_m / L.VERTICAL_SCALE), dtype=np.int16)
    stone = max(1, int(stone_size_m / L.HORIZONTAL_SCALE))
    gap = max(1, int(round(gap_m / L.HORIZONTAL_SCALE)))
    x = 0
    while x < px:
        sx = min(px, x + stone)
        y = int(rng.integers(0, stone))
        # first (partial) stone of the column
        h[x:sx, 0:max(0, y - gap)] = 0
        while y < px:
            sy = min(px, y + stone)
            h[x:sx, y:sy] = 0
            y += stone + gap
        x += stone + gap
    c = px // 2
    p = int(platform_m / L.HORIZONTAL_SCALE / 2)
    h[c - p:c + p, c - p:c + p] = 0
    return h


def _pyramid_stairs(px, step_width_m, step_height_m, platform_m=3.0):
    h = np.zeros((px, px), dtype=np.int16)
    sw = int(step_width_m / L.HORIZONTAL_SCALE)
    sh = int(step_height_m / L.VERTICAL_SCALE)
    plat = int(platform_m / L.HORIZONTAL_SCALE)
    height = 0
    lo, hi = 0, px
    while hi - lo > plat:
        lo += sw
        hi -= sw
        height += sh
        h[lo:hi, lo:hi] = height
    return h


def _discrete_obstacles(rng, px, max_height_m, n=20, platform_m=3.0):
    h = np.zeros((px, px), dtype=np.int16)
    mh = int(max_height_m / L.VERTICAL_SCALE)
    for _ in range(n):
        w = int(rng.integers(20, 41))
        l = int(rng.integers(20, 41))
        x = int(rng.integers(0, px - w))
        y = int(rng.integers(0, px - l))
        h[x:x + w, y:y + l] = int(rng.choice([-mh, -mh // 2, mh // 2, mh]))
    c = px // 2
    p = int(platform_m / L.HORIZONTAL_SCALE / 2)
    h[c - p:c + p, c - p:c + p] = 0
    return h


def make_heightmap(kind="stones", seed=0):
    """Returns (height_samples int16 [1760,1120] numpy, terrain_origins float32 [6,2,3] numpy).

    kind: "flat" (cfg 1; all-zero heightfield), "stones" (cfg 2/3/5), "curriculum" (cfg 4).
    """
    rng = np.random.default_rng(seed)
    hs = np.zeros((L.MAP_ROWS, L.MAP_COLS), dtype=np.int16)
    origins = np.zeros((L.NUM_ROWS, L.NUM_COLS, 3), dtype=np.float32)
    for j in range(L.NUM_COLS):
        for i in range(L.NUM_ROWS):
            x0, y0, px = _sub_origin(i, j)
            d = i / L.NUM_ROWS
            if kind == "flat":
                sub = np.zeros((px, px), dtype=np.int16)
            elif kind == "stones":
                size = float(rng.uniform(0.3, 1.0))
                gap = float(rng.uniform(0.06, 0.10))
                sub = _stepping_stones(rng, px, size, gap)
            elif kind == "curriculum":
                # proportions [0,0,.2,.2,.2,.4] over the 2 columns x 6 rows: column 0 rows cycle through
                # stairs-down / stairs-up / discrete, column 1 is stepping stones (weight .4 + remainder)
                if j == 0:
                    which = i % 3
                    if which == 0:
                        sub = _pyramid_stairs(px, 0.31, -(0.05 + 0.13 * d))
                    elif which == 1:
                        sub = _pyramid_stairs(px, 0.31, 0.05 + 0.13 * d)
                    else:
                        sub = _discrete_obstacles(rng, px, 0.05 + 0.15 * d)
                else:
                    sub = _stepping_stones(rng, px, 1.0 * (1.05 - d), 0.03 if d == 0 else 0.06)
            else:
                raise ValueError(kind)
            hs[x0:x0 + px, y0:y0 + px] = sub
            c = px // 2
            origins[i, j, 0] = (i + 0.5) * L.TERRAIN_LENGTH
            origins[i, j, 1] = (j + 0.5) * L.TERRAIN_LENGTH
            origins[i, j, 2] = float(sub[c - 10:c + 10, c - 10:c + 10].max()) * L.VERTICAL_SCALE
    return hs, origins


def describe_terrain(kind="stones", seed=0):
    """The random parameters of `make_heightmap(kind, seed)` - same generator, same draw order - as the two arrays
    `dtc_terrain_rasterize` (include/dtc_b200.h) takes: subs int32 [12, 5] = (type, a, b, c, platform_half) per
    sub-terrain in row-major (i, j) order and tables int32 [12, 256]."""
    rng = np.random.default_rng(seed)
    n = L.NUM_ROWS * L.NUM_COLS
    subs = np.zeros((n, 5), dtype=np.int32)
    tables = np.zeros((n, 256), dtype=np.int32)
    px = int(L.TERRAIN_LENGTH / L.HORIZONTAL_SCALE)

    def stones(s, size_m, gap_m, depth_m=-2.0, platform_m=1.0):
        stone = max(1, int(size_m / L.HORIZONTAL_SCALE))
        gap = max(1, int(round(gap_m / L.HORIZONTAL_SCALE)))
        x = k = 0
        while x < px:
            tables[s, k] = int(rng.integers(0, stone))
            k += 1
            x += stone + gap
        subs[s] = (1, stone, gap, int(depth_m / L.VERTICAL_SCALE), int(platform_m / L.HORIZONTAL_SCALE / 2))

    for j in range(L.NUM_COLS):
        for i in range(L.NUM_ROWS):
            s, d = i * L.NUM_COLS + j, i / L.NUM_ROWS
            if kind == "flat":
                continue
            if kind == "stones":
                size = float(rng.uniform(0.3, 1.0))
                gap = float(rng.uniform(0.06, 0.10))
                stones(s, size, gap)
            elif kind == "curriculum":
                if j == 0:
                    which = i % 3
                    if which < 2:
                        step_h = -(0.05 + 0.13 * d) if which == 0 else 0.05 + 0.13 * d
                        subs[s] = (2, int(0.31 / L.HORIZONTAL_SCALE), int(step_h / L.VERTICAL_SCALE), int(3.0 / L.HORIZONTAL_SCALE), 0)
                    else:
                        mh = int((0.05 + 0.15 * d) / L.VERTICAL_SCALE)
                        for r in range(20):
                            w = int(rng.integers(20, 41))
                            l = int(rng.integers(20, 41))
                            x = int(rng.integers(0, px - w))
                            y = int(rng.integers(0, px - l))
                            tables[s, 5 * r:5 * r + 5] = (x, y, w, l, int(rng.choice([-mh, -mh // 2, mh // 2, mh])))
                        subs[s] = (3, 20, 0, 0, int(3.0 / L.HORIZONTAL_SCALE / 2))
                else:
                    stones(s, 1.0 * (1.05 - d), 0.03 if d == 0 else 0.06)
            else:
                raise ValueError(kind)
    return subs, tables


def make_heightmap_device(kind="stones", seed=0, device="cuda"):
    """`make_heightmap` with the rasterisation on the device: (height_samples int16 [1760,1120], terrain_origins float32 [6,2,3])
    as CUDA tensors, bit-identical to the numpy version (tests/test_env_gpu.py::test_terrain_rasterize_matches_numpy)."""
    import ctypes as C
    from . import _lib as B
    subs, tables = describe_terrain(kind, seed)
    dev = torch.device(device)
    d_subs, d_tab = torch.from_numpy(subs).to(dev), torch.from_numpy(tables).to(dev)
    hs = torch.empty(L.MAP_ROWS, L.MAP_COLS, dtype=torch.int16, device=dev)
    origins = torch.empty(L.NUM_ROWS, L.NUM_COLS, 3, device=dev)
    B.check(B.lib().dtc_terrain_rasterize(L.MAP_ROWS, L.MAP_COLS, int(L.BORDER_SIZE / L.HORIZONTAL_SCALE),
                                          int(L.TERRAIN_LENGTH / L.HORIZONTAL_SCALE), L.NUM_ROWS, L.NUM_COLS, B.ptr(d_subs), B.ptr(d_tab),
                                          C.c_double(L.TERRAIN_LENGTH), C.c_double(L.VERTICAL_SCALE), B.ptr(hs), B.ptr(origins),
                                          B.stream_ptr(dev)), "dtc_terrain_rasterize")
    return hs, origins


# ----------------------------------------------------------------------------- synthetic state
HIP_OFFSETS = torch.tensor([[0.1745, 0.159, 0.0], [0.1745, -0.159, 0.0],
                            [-0.1745, 0.159, 0.0], [-0.1745, -0.159, 0.0]])


def _euler_to_quat(roll, pitch, yaw):
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    return torch.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp,
                        sy * cr * cp - cy * sr * sp, cy * cr * cp + sy * sr * sp], dim=-1)


def synth_state(num_envs, env_origins, gen, device="cpu"):
    """One draw of the four simulator tensors (SURVEY 8d distributions). `gen` is a torch.Generator
    on `device`.  Returns dict of float32 tensors with the reference's flat shapes."""
    N = num_envs
    kw = dict(generator=gen, device=device, dtype=torch.float32)
    o = env_origins.to(device)
    root = torch.zeros(N, 13, device=device)
    root[:, 0:2] = o[:, 0:2] + (torch.rand(N, 2, **kw) * 4.0 - 2.0)
    height = 0.30 + 0.10 * torch.rand(N, **kw)
    root[:, 2] = o[:, 2] + height
    rp = torch.rand(N, 2, **kw) * 0.2 - 0.1
    yaw = (torch.rand(N, **kw) * 2.0 - 1.0) * math.pi
    root[:, 3:7] = _euler_to_quat(rp[:, 0], rp[:, 1], yaw)
    root[:, 7:13] = torch.randn(N, 6, **kw) * 0.3

    dof = torch.zeros(N, L.NUM_DOF, 2, device=device)
    dof[..., 0] = torch.tensor(L.DEFAULT_DOF_POS, device=device) + torch.randn(N, L.NUM_DOF, **kw) * 0.2
    dof[..., 1] = torch.randn(N, L.NUM_DOF, **kw)

    rb = torch.zeros(N, L.NUM_BODIES, 13, device=device)
    rb[:, :, 0:3] = root[:, None, 0:3]
    rb[:, :, 3:7] = root[:, None, 3:7]
    rb[:, :, 7:10] = torch.randn(N, L.NUM_BODIES, 3, **kw) * 0.5
    cy, sy = torch.cos(yaw), torch.sin(yaw)
    hip = HIP_OFFSETS.to(device)
    hx = cy[:, None] * hip[None, :, 0] - sy[:, None] * hip[None, :, 1]
    hy = sy[:, None] * hip[None, :, 0] + cy[:, None] * hip[None, :, 1]
    thigh = torch.stack([root[:, None, 0] + hx, root[:, None, 1] + hy, root[:, None, 2].expand(N, 4)], dim=-1)
    rb[:, L.THIGH_INDICES, 0:3] = thigh
    foot = thigh.clone()
    foot[..., 0:2] += torch.randn(N, 4, 2, **kw) * 0.05
    foot[..., 2] = (root[:, 2] - height)[:, None] + 0.1 * torch.rand(N, 4, **kw)
    rb[:, L.FEET_INDICES, 0:3] = foot

    cf = torch.zeros(N, L.NUM_BODIES, 3, device=device)
    cf[:, L.FEET_INDICES, 2] = 40.0 * (torch.rand(N, 4, **kw) < 0.5).float()
    # sparse tangential foot forces and body contacts so collision / stumble terms are exercised
    cf[:, L.FEET_INDICES, 0:2] = torch.randn(N, 4, 2, **kw) * 60.0 * (torch.rand(N, 4, 1, **kw) < 0.05).float()
    body_hit = (torch.rand(N, L.NUM_BODIES, **kw) < 0.01).float()
    body_hit[:, L.FEET_INDICES] = 0.0
    cf[:, :, 2] += 5.0 * body_hit
    return {"root_states": root, "dof_state": dof.reshape(N * L.NUM_DOF, 2),
            "net_contact_force": cf.reshape(N * L.NUM_BODIES, 3),
            "rigid_body_state": rb.reshape(N * L.NUM_BODIES, 13)}


class FakeGym:
    """Stand-in for the `gym` object: owns the four persistent state tensors; `refresh_actor_root_state_tensor`
    (the first refresh of `post_physics_step`, legged_robot_dtc.py:61) makes the next queued/generated
    synthetic state visible, exactly where PhysX would have."""

    def __init__(self, num_envs, device="cpu"):
        N = num_envs
        self.num_envs = N
        self.device = device
        self.root_states = torch.zeros(N, 13, device=device)
        self.root_states[:, 6] = 1.0
        self.dof_state = torch.zeros(N * L.NUM_DOF, 2, device=device)
        self.net_contact_force = torch.zeros(N * L.NUM_BODIES, 3, device=device)
        self.rigid_body_state = torch.zeros(N * L.NUM_BODIES, 13, device=device)
        self.static_dof_state = True  # dof_state does not change inside the decimation loop (LeggedRobotDTC.step fuses the 4 sub-steps)
        self.queue = []  # list of state dicts consumed FIFO
        self.source = None  # optional callable() -> state dict, used when the queue is empty
        self._pf = None     # host->device double buffering (enable_prefetch)
        self.graph_step = None  # (t, T) while a T-step rollout is being captured into a CUDA graph (OnPolicyRunner)

    def load(self, st):
        self.root_states.copy_(st["root_states"], non_blocking=True)
        self.dof_state.copy_(st["dof_state"], non_blocking=True)
        self.net_contact_force.copy_(st["net_contact_force"], non_blocking=True)
        self.rigid_body_state.copy_(st["rigid_body_state"], non_blocking=True)

    def enable_prefetch(self, on=True):
        """States arriving from PINNED HOST memory through `source`: copy step t+1's four tensors host->device on a copy stream
        into a staging set while step t's kernels run; refresh() then only waits for the staging set and moves it into the
        simulator tensors device-to-device.  Same bytes over PCIe per step, off the compute stream's critical path."""
        if not on:
            self._pf = None
            return
        dev = self.root_states.device
        self._pf = dict(stream=torch.cuda.Stream(dev), ready=torch.cuda.Event(), free=torch.cuda.Event(), staged=False,
                        buf={k: torch.empty_like(v) for k, v in self._tensors().items()})
        self._pf["free"].record(torch.cuda.current_stream(dev))

    def _tensors(self):
        return {"root_states": self.root_states, "dof_state": self.dof_state, "net_contact_force": self.net_contact_force,
                "rigid_body_state": self.rigid_body_state}

    def _prefetch_next(self):
        pf, st = self._pf, self.source()
        with torch.cuda.stream(pf["stream"]):
            pf["stream"].wait_event(pf["free"])  # the previous device-to-device move out of the staging set has been queued
            for k, v in pf["buf"].items():
                v.copy_(st[k], non_blocking=True)
            pf["ready"].record(pf["stream"])
        pf["staged"] = True

    def _take_prefetched(self):
        pf = self._pf
        cur = torch.cuda.current_stream(self.root_states.device)
        if self.graph_step is not None:
            # inside a CUDA-graph capture of a T-step rollout: step t moves ring[t] into the simulator tensors device-to-device;
            # the ring itself is filled from pinned host memory OUTSIDE the graph, on the copy stream, while the previous
            # iteration's update runs (graph_after_replay / graph_before_replay)
            t, T = self.graph_step
            if t == 0:
                pf["ring"] = [{k: torch.empty_like(v) for k, v in self._tensors().items()} for _ in range(T)]
                pf["ring_pending"] = False
            for k, v in self._tensors().items():
                v.copy_(pf["ring"][t][k], non_blocking=True)
            return
        if not pf["staged"]:
            self._prefetch_next()
        cur.wait_event(pf["ready"])
        for k, v in self._tensors().items():
            v.copy_(pf["buf"][k], non_blocking=True)
        pf["free"].record(cur)
        pf["staged"] = False
        self._prefetch_next()  # the next step's state starts flowing while this step's kernels run

    def _upload_ring(self):
        pf = self._pf
        with torch.cuda.stream(pf["stream"]):
            pf["stream"].wait_event(pf["free"])  # the replay that read the ring last has finished with it
            for slot in pf["ring"]:
                # a state an eager step prefetched before the capture is consumed first (queued earlier on this same stream)
                st, pf["staged"] = (pf["buf"], False) if pf["staged"] else (self.source(), False)
                for k, v in slot.items():
                    v.copy_(st[k], non_blocking=True)
            pf["ready"].record(pf["stream"])
        pf["ring_pending"] = True

    def graph_before_replay(self):
        """Runner hook: the captured rollout is about to be replayed -> its T states must be in the ring."""
        pf = self._pf
        if pf is None or pf.get("ring") is None:
            return
        if not pf["ring_pending"]:
            self._upload_ring()
        torch.cuda.current_stream(self.root_states.device).wait_event(pf["ready"])
        pf["ring_pending"] = False

    def graph_after_replay(self):
        """Runner hook: a replay has been queued -> start uploading the next rollout's states behind it (they travel over PCIe
        while the update runs)."""
        pf = self._pf
        if pf is None or pf.get("ring") is None:
            return
        pf["free"].record(torch.cuda.current_stream(self.root_states.device))
        self._upload_ring()

    # --- tensor API
    def acquire_actor_root_state_tensor(self, sim): return self.root_states
    def acquire_dof_state_tensor(self, sim): return self.dof_state
    def acquire_net_contact_force_tensor(self, sim): return self.net_contact_force
    def acquire_rigid_body_state_tensor(self, sim): return self.rigid_body_state

    def refresh_actor_root_state_tensor(self, sim):
        if self.queue:
            self.load(self.queue.pop(0))
        elif self.source is not None:
            if self._pf is not None:
                self._take_prefetched()
            else:
                self.load(self.source())

    def refresh_dof_state_tensor(self, sim): pass
    def refresh_net_contact_force_tensor(self, sim): pass
    def refresh_rigid_body_state_tensor(self, sim): pass
    def simulate(self, sim): pass
    def fetch_results(self, sim, flag): pass
    def set_dof_actuation_force_tensor(self, sim, t): pass
    def set_dof_state_tensor_indexed(self, sim, t, ids, n): pass
    def set_dof_state_tensor(self, sim, t): pass
    def set_actor_root_state_tensor(self, sim, t): pass
    def set_actor_root_state_tensor_indexed(self, sim, t, ids, n): pass
    def clear_lines(self, viewer): pass
    def get_sim_time(self, sim): return 0.0


def initial_env_layout(num_envs, terrain_origins, seed, curriculum=True):
    """terrain_levels / terrain_types / env_origins as `_get_env_origins` (legged_robot.py:1201-1215)."""
    g = torch.Generator().manual_seed(seed)
    max_init = 5 if curriculum else L.NUM_ROWS - 1
    levels = torch.randint(0, max_init + 1, (num_envs,), generator=g)
    types = torch.div(torch.arange(num_envs), (num_envs / L.NUM_COLS), rounding_mode="floor").to(torch.long)
    to = (terrain_origins.detach().cpu() if torch.is_tensor(terrain_origins) else torch.from_numpy(np.asarray(terrain_origins))).to(torch.float)
    return levels, types, to[levels, types].clone(), to
