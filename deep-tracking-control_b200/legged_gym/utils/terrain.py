"""`Terrain` - the reference's terrain generator (legged_gym/utils/terrain.py:9-243) with the rasterisation on the device.

Same constructor, attribute names and numpy draw order as the reference class: `Terrain(cfg.terrain, num_robots)` lays
`num_rows x num_cols` sub-terrains (curriculum / randomized) inside a flat border and exposes `height_field_raw` / `heightsamples`
(int16 [tot_rows, tot_cols]), `env_origins` [num_rows, num_cols, 3], `tot_rows`, `tot_cols`, `border`, `env_length`, `env_width`.
Every generator the DTC tasks use is a sequence of `height_field_raw[x0:x1, y0:y1] = h` assignments on a constant background; the
generators here replay the loops and draws of terrain.py / isaacgym.terrain_utils and RECORD the assignments, and
`dtc_terrain_paint` (csrc/dtc_terrain.cu) paints the whole map and computes the env-origin heights - the 3.9 MB heightmap never
exists on the host.  With the same `numpy.random.seed` the map equals the reference's bit for bit
(tests/test_env_gpu.py::test_terrain_class_matches_reference_golden).

Not supported: the two sloped terrain types (proportions[0], [1]; zero in every DTC task) and the trimesh conversion (PhysX only)."""
import ctypes as C

import numpy as np
import torch

from ... import _lib as B


class _Sub:
    """One sub-terrain under construction: background + ordered rectangle list with numpy's slice semantics."""

    def __init__(self, width, length, vertical_scale, horizontal_scale):
        self.width, self.length = width, length
        self.vertical_scale, self.horizontal_scale = vertical_scale, horizontal_scale
        self.background, self.rects = 0, []

    def fill(self, h):
        self.background, self.rects = int(h), []

    def assign(self, x0, x1, y0, y1, h):
        """height_field_raw[x0:x1, y0:y1] = h, including numpy's clipping and negative-index wrap."""
        xa, xb, _ = slice(int(x0), int(x1)).indices(self.width)
        ya, yb, _ = slice(int(y0), int(y1)).indices(self.length)
        if xa < xb and ya < yb:
            self.rects.append((xa, xb, ya, yb, int(np.int16(h))))


def pyramid_stairs_terrain(t, step_width, step_height, platform_size=1.0):
    """isaacgym.terrain_utils.pyramid_stairs_terrain (published algorithm)."""
    step_width = int(step_width / t.horizontal_scale)
    step_height = int(step_height / t.vertical_scale)
    platform_size = int(platform_size / t.horizontal_scale)
    height, x0, x1, y0, y1 = 0, 0, t.width, 0, t.length
    while (x1 - x0) > platform_size and (y1 - y0) > platform_size:
        x0, x1, y0, y1 = x0 + step_width, x1 - step_width, y0 + step_width, y1 - step_width
        height += step_height
        t.assign(x0, x1, y0, y1, height)


def discrete_obstacles_terrain(t, max_height, min_size, max_size, num_rects, platform_size=1.0):
    """isaacgym.terrain_utils.discrete_obstacles_terrain (published algorithm; five numpy draws per rectangle)."""
    max_height = int(max_height / t.vertical_scale)
    min_size, max_size = int(min_size / t.horizontal_scale), int(max_size / t.horizontal_scale)
    platform_size = int(platform_size / t.horizontal_scale)
    i, j = t.width, t.length
    height_range = [-max_height, -max_height // 2, max_height // 2, max_height]
    for _ in range(num_rects):
        width = np.random.choice(range(min_size, max_size, 4))
        length = np.random.choice(range(min_size, max_size, 4))
        si = np.random.choice(range(0, i - width, 4))
        sj = np.random.choice(range(0, j - length, 4))
        t.assign(si, si + width, sj, sj + length, np.random.choice(height_range))
    t.assign((t.width - platform_size) // 2, (t.width + platform_size) // 2, (t.length - platform_size) // 2, (t.length + platform_size) // 2, 0)


def _stones(t, size_draw, dist_draw, height_draw, depth, platform_size):
    """Shared loop of stepping_stones_terrain (fixed stone size / distance) and stones_everywhere_terrain (drawn per use)."""
    t.fill(int(depth / t.vertical_scale))
    start_x = start_y = 0
    if t.length >= t.width:
        while start_y < t.length:
            stop_y = min(t.length, start_y + size_draw())
            start_x = np.random.randint(0, size_draw())
            t.assign(0, max(0, start_x - dist_draw()), start_y, stop_y, height_draw())  # first hole
            while start_x < t.width:
                t.assign(start_x, min(t.width, start_x + size_draw()), start_y, stop_y, height_draw())
                start_x += size_draw() + dist_draw()
            start_y += size_draw() + dist_draw()
    else:
        while start_x < t.width:
            stop_x = min(t.width, start_x + size_draw())
            start_y = np.random.randint(0, size_draw())
            t.assign(start_x, stop_x, 0, max(0, start_y - dist_draw()), height_draw())
            while start_y < t.length:
                t.assign(start_x, stop_x, start_y, min(t.length, start_y + size_draw()), height_draw())
                start_y += size_draw() + dist_draw()
            start_x += size_draw() + dist_draw()
    t.assign((t.width - platform_size) // 2, (t.width + platform_size) // 2, (t.length - platform_size) // 2, (t.length + platform_size) // 2, 0)


def stepping_stones_terrain(t, stone_size, stone_distance, max_height, platform_size=1.0, depth=-10):
    """isaacgym.terrain_utils.stepping_stones_terrain (published algorithm): constant stone size / distance, one draw per stone."""
    size, dist = int(stone_size / t.horizontal_scale), int(stone_distance / t.horizontal_scale)
    height_range = np.arange(-int(max_height / t.vertical_scale) - 1, int(max_height / t.vertical_scale), step=1)
    _stones(t, lambda: size, lambda: dist, lambda: np.random.choice(height_range), depth, int(platform_size / t.horizontal_scale))


def stones_everywhere_terrain(t, stone_size, stone_distance, max_height, platform_size=1.0, depth=-10):
    """terrain.py:186-243: stone size, distance and height drawn anew at every use."""
    mx = int(stone_size / t.horizontal_scale)
    sizes = np.arange(mx - 1, mx + 1, step=1)
    md = int(stone_distance / t.horizontal_scale)
    dists = np.arange(md, md + 1, step=1)
    heights = np.arange(1, 2 * int(max_height / t.vertical_scale) + 1, step=1)
    rc = np.random.choice
    _stones(t, lambda: rc(sizes), lambda: rc(dists), lambda: rc(heights), depth, int(platform_size / t.horizontal_scale))


def gap_terrain(t, gap_size, platform_size=1.0):
    """terrain.py:162-174."""
    gap_size, platform_size = int(gap_size / t.horizontal_scale), int(platform_size / t.horizontal_scale)
    cx, cy = t.length // 2, t.width // 2
    x1 = (t.length - platform_size) // 2
    y1 = (t.width - platform_size) // 2
    x2, y2 = x1 + gap_size, y1 + gap_size
    t.assign(cx - x2, cx + x2, cy - y2, cy + y2, -1000)
    t.assign(cx - x1, cx + x1, cy - y1, cy + y1, 0)


def pit_terrain(t, depth, platform_size=1.0):
    """terrain.py:176-183."""
    depth, p = int(depth / t.vertical_scale), int(platform_size / t.horizontal_scale / 2)
    t.assign(t.length // 2 - p, t.length // 2 + p, t.width // 2 - p, t.width // 2 + p, -depth)


class Terrain:
    def __init__(self, cfg, num_robots, device="cuda"):
        self.cfg, self.num_robots, self.type = cfg, num_robots, cfg.mesh_type
        if self.type in ["none", "plane"]:
            return
        # device=None: lay the sub-terrains out (every numpy draw happens, `sub_terrains` holds the rectangle lists) but do not
        # rasterise - there is no height field then.  Host-logic tests use it; anything else needs a CUDA device.
        self.device = None if device is None else torch.device(device)
        if self.device is not None and self.device.type != "cuda":
            raise B.DtcError("Terrain rasterises on a CUDA device only (no CPU fallback)")
        self.env_length, self.env_width = cfg.terrain_length, cfg.terrain_width
        self.proportions = [np.sum(cfg.terrain_proportions[:i + 1]) for i in range(len(cfg.terrain_proportions))]
        self.cfg.num_sub_terrains = cfg.num_rows * cfg.num_cols
        self.width_per_env_pixels = int(self.env_width / cfg.horizontal_scale)
        self.length_per_env_pixels = int(self.env_length / cfg.horizontal_scale)
        self.border = int(cfg.border_size / cfg.horizontal_scale)
        self.tot_cols = int(cfg.num_cols * self.width_per_env_pixels) + 2 * self.border
        self.tot_rows = int(cfg.num_rows * self.length_per_env_pixels) + 2 * self.border
        self._subs = [None] * (cfg.num_rows * cfg.num_cols)
        if cfg.curriculum:
            self.curiculum()
        elif getattr(cfg, "selected", False):
            raise NotImplementedError("cfg.terrain.selected (eval of a terrain_utils function name) is not supported")
        else:
            self.randomized_terrain()
        if self.device is not None:
            self._paint()

    @property
    def sub_terrains(self):
        """[(background, [(x0, x1, y0, y1, h), ...]), ...] in row-major (row, col) order: what dtc_terrain_paint receives."""
        return [(t.background, list(t.rects)) for t in self._subs]

    # ------------------------------------------------------------------ layout (terrain.py:45-63)
    def randomized_terrain(self):
        for k in range(self.cfg.num_sub_terrains):
            i, j = np.unravel_index(k, (self.cfg.num_rows, self.cfg.num_cols))
            choice = np.random.uniform(0, 1)
            difficulty = np.random.choice([0.25, 0.5, 0.75, 0.9])
            self._subs[i * self.cfg.num_cols + j] = self.make_terrain(choice, difficulty)

    def curiculum(self):
        for j in range(self.cfg.num_cols):
            for i in range(self.cfg.num_rows):
                self._subs[i * self.cfg.num_cols + j] = self.make_terrain(j / self.cfg.num_cols + 0.001, i / self.cfg.num_rows)

    def make_terrain(self, choice, difficulty):
        """terrain.py:79-141 with the `#! lite3` parameter set (the assignments that win in the reference)."""
        t = _Sub(self.width_per_env_pixels, self.width_per_env_pixels, self.cfg.vertical_scale, self.cfg.horizontal_scale)
        stepping_stones_size = 1 * (1.05 - difficulty)
        step_height = 0.05 + 0.13 * difficulty
        discrete_obstacles_height = 0.05 + difficulty * 0.15
        stone_distance = 0.03 if difficulty == 0 else 0.06
        max_height = 0.02 + 0.03 * difficulty
        stone_size = -0.1 * difficulty + 0.3
        gap_size = 0.8 * difficulty
        pit_depth = 0.8 * difficulty
        p = self.proportions
        if choice < p[1]:
            raise NotImplementedError("sloped terrain types (terrain_proportions[0], [1]) are not supported by the device painter")
        elif choice < p[3]:
            if choice < p[2]:
                step_height *= -1
            pyramid_stairs_terrain(t, step_width=0.31, step_height=step_height, platform_size=3.0)
        elif choice < p[4]:
            discrete_obstacles_terrain(t, discrete_obstacles_height, 1.0, 2.0, 20, platform_size=3.0)
        elif choice < p[5]:
            stepping_stones_terrain(t, stone_size=stepping_stones_size, stone_distance=stone_distance, max_height=0.0, platform_size=1.0, depth=-2)
        elif choice < p[6]:
            gap_terrain(t, gap_size=gap_size, platform_size=1.0)
        elif choice < p[7]:
            pit_terrain(t, depth=pit_depth, platform_size=1.0)
        else:
            stones_everywhere_terrain(t, stone_size=stone_size, stone_distance=stone_distance, max_height=max_height, platform_size=1.3, depth=-2)
        return t

    # ------------------------------------------------------------------ map assembly + env origins on the device (terrain.py:143-160)
    def _paint(self):
        cfg, dev = self.cfg, self.device
        subs = np.zeros((len(self._subs), 5), dtype=np.int32)
        rects, first = [], 0
        for s, t in enumerate(self._subs):
            subs[s] = (5, len(t.rects), first, t.background, 0)
            rects.extend(t.rects)
            first += len(t.rects)
        rect_arr = np.asarray(rects, dtype=np.int32).reshape(-1, 5) if rects else np.zeros((1, 5), dtype=np.int32)
        d_subs, d_rects = torch.from_numpy(subs).to(dev), torch.from_numpy(rect_arr).to(dev)
        self.height_field_raw = torch.empty(self.tot_rows, self.tot_cols, dtype=torch.int16, device=dev)
        origins = torch.empty(cfg.num_rows, cfg.num_cols, 3, device=dev)
        hs = cfg.horizontal_scale
        win = (C.c_int32 * 4)(int((self.env_length / 2.0 - 1) / hs), int((self.env_length / 2.0 + 1) / hs),
                              int((self.env_width / 2.0 - 1) / hs), int((self.env_width / 2.0 + 1) / hs))
        B.check(B.lib().dtc_terrain_paint(self.tot_rows, self.tot_cols, self.border, self.length_per_env_pixels, self.width_per_env_pixels,
                                          cfg.num_rows, cfg.num_cols, B.ptr(d_subs), B.ptr(d_rects), win, C.c_double(self.env_length),
                                          C.c_double(self.env_width), C.c_double(cfg.vertical_scale), B.ptr(self.height_field_raw),
                                          B.ptr(origins), B.stream_ptr(dev)), "dtc_terrain_paint")
        self.heightsamples = self.height_field_raw
        self.terrain_origins = origins          # device tensor [num_rows, num_cols, 3] (what LeggedRobotDTC consumes)
        self._num_rects = len(rects)

    @property
    def env_origins(self):
        """numpy [num_rows, num_cols, 3] like the reference attribute (one small device->host copy)."""
        return self.terrain_origins.double().cpu().numpy()
