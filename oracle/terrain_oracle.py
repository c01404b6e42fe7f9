"""TEST INFRASTRUCTURE ONLY - CPU restatement (oracle) of the reference's terrain generator, `legged_gym/utils/terrain.py`.

Restates `Terrain.__init__` (:9-43), `curiculum` (:56-63), `randomized_terrain` (:45-54), `make_terrain` (:79-141: the difficulty
formulas and the proportion ladder), `add_terrain_to_map` (:143-160: map assembly and env origins) and the reference's own
generators `gap_terrain`, `pit_terrain`, `stones_everywhere_terrain` (:162-243) in plain numpy, drawing from numpy's global
generator in the reference's order.  The Isaac Gym generators it calls (pyramid stairs, discrete obstacles, stepping stones) are
third-party code absent from /root/reference; their published algorithm is restated in oracle/ref_harness/stubs/isaacgym/
terrain_utils.py ("parity unpinned" at that boundary) and imported from there.

PINNING: tests/test_oracle_golden.py::test_terrain_matches_reference compares this against tests/golden/terrain_*.npz, recorded by
running the UNMODIFIED reference class through the stubs (tests/golden/make_golden.py)."""
import importlib.util
import os

import numpy as np

_spec = importlib.util.spec_from_file_location(
    "_dtc_terrain_utils_stub", os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_harness", "stubs", "isaacgym", "terrain_utils.py"))
TU = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(TU)


def gap_terrain(t, gap_size, platform_size=1.0):
    """terrain.py:162-174 (negative slice starts wrap, as numpy does for the high-difficulty rows)."""
    gap_size = int(gap_size / t.horizontal_scale)
    platform_size = int(platform_size / t.horizontal_scale)
    cx, cy = t.length // 2, t.width // 2
    x1 = (t.length - platform_size) // 2
    x2 = x1 + gap_size
    y1 = (t.width - platform_size) // 2
    y2 = y1 + gap_size
    t.height_field_raw[cx - x2:cx + x2, cy - y2:cy + y2] = -1000
    t.height_field_raw[cx - x1:cx + x1, cy - y1:cy + y1] = 0


def pit_terrain(t, depth, platform_size=1.0):
    """terrain.py:176-183."""
    depth = int(depth / t.vertical_scale)
    platform_size = int(platform_size / t.horizontal_scale / 2)
    x1, x2 = t.length // 2 - platform_size, t.length // 2 + platform_size
    y1, y2 = t.width // 2 - platform_size, t.width // 2 + platform_size
    t.height_field_raw[x1:x2, y1:y2] = -depth


def stones_everywhere_terrain(t, stone_size, stone_distance, max_height, platform_size=1.0, depth=-10):
    """terrain.py:186-243 (length >= width branch and its mirror)."""
    max_stone_size = int(stone_size / t.horizontal_scale)
    stone_size = np.arange(max_stone_size - 1, max_stone_size + 1, step=1)
    max_stone_distance = int(stone_distance / t.horizontal_scale)
    stone_distance = np.arange(max_stone_distance, max_stone_distance + 1, step=1)
    max_height = int(max_height / t.vertical_scale)
    platform_size = int(platform_size / t.horizontal_scale)
    height_range = np.arange(1, 2 * max_height + 1, step=1)
    start_x = start_y = 0
    t.height_field_raw[:, :] = int(depth / t.vertical_scale)
    rc = np.random.choice
    if t.length >= t.width:
        while start_y < t.length:
            stop_y = min(t.length, start_y + rc(stone_size))
            start_x = np.random.randint(0, rc(stone_size))
            stop_x = max(0, start_x - rc(stone_distance))
            t.height_field_raw[0:stop_x, start_y:stop_y] = rc(height_range)
            while start_x < t.width:
                stop_x = min(t.width, start_x + rc(stone_size))
                t.height_field_raw[start_x:stop_x, start_y:stop_y] = rc(height_range)
                start_x += rc(stone_size) + rc(stone_distance)
            start_y += rc(stone_size) + rc(stone_distance)
    else:
        while start_x < t.width:
            stop_x = min(t.width, start_x + rc(stone_size))
            start_y = np.random.randint(0, rc(stone_size))
            stop_y = max(0, start_y - rc(stone_distance))
            t.height_field_raw[start_x:stop_x, 0:stop_y] = rc(height_range)
            while start_y < t.length:
                stop_y = min(t.length, start_y + rc(stone_size))
                t.height_field_raw[start_x:stop_x, start_y:stop_y] = rc(height_range)
                start_y += rc(stone_size) + rc(stone_distance)
            start_x += rc(stone_size) + rc(stone_distance)
    x1, x2 = (t.width - platform_size) // 2, (t.width + platform_size) // 2
    y1, y2 = (t.length - platform_size) // 2, (t.length + platform_size) // 2
    t.height_field_raw[x1:x2, y1:y2] = 0


def make_terrain(tc, choice, difficulty, wpx):
    """terrain.py:79-141 with the `#! lite3` parameter set (the later assignments win)."""
    t = TU.SubTerrain("terrain", width=wpx, length=wpx, vertical_scale=tc.vertical_scale, horizontal_scale=tc.horizontal_scale)
    slope = difficulty * 0.4
    stepping_stones_size = 1 * (1.05 - difficulty)
    step_height = 0.05 + 0.13 * difficulty
    discrete_obstacles_height = 0.05 + difficulty * 0.15
    stone_distance = 0.03 if difficulty == 0 else 0.06
    max_height = 0.02 + 0.03 * difficulty
    stone_size = -0.1 * difficulty + 0.3
    gap_size = 0.8 * difficulty
    pit_depth = 0.8 * difficulty
    p = [np.sum(tc.terrain_proportions[:i + 1]) for i in range(len(tc.terrain_proportions))]
    if choice < p[0]:
        if choice < p[0] / 2:
            slope *= -1
        TU.pyramid_sloped_terrain(t, slope=slope, platform_size=3.0)
    elif choice < p[1]:
        TU.pyramid_sloped_terrain(t, slope=slope, platform_size=3.0)
        TU.random_uniform_terrain(t, min_height=-0.05, max_height=0.05, step=0.005, downsampled_scale=0.2)
    elif choice < p[3]:
        if choice < p[2]:
            step_height *= -1
        TU.pyramid_stairs_terrain(t, step_width=0.31, step_height=step_height, platform_size=3.0)
    elif choice < p[4]:
        TU.discrete_obstacles_terrain(t, discrete_obstacles_height, 1.0, 2.0, 20, platform_size=3.0)
    elif choice < p[5]:
        TU.stepping_stones_terrain(t, stone_size=stepping_stones_size, stone_distance=stone_distance, max_height=0.0, platform_size=1.0, depth=-2)
    elif choice < p[6]:
        gap_terrain(t, gap_size=gap_size, platform_size=1.0)
    elif choice < p[7]:
        pit_terrain(t, depth=pit_depth, platform_size=1.0)
    else:
        stones_everywhere_terrain(t, stone_size=stone_size, stone_distance=stone_distance, max_height=max_height, platform_size=1.3, depth=-2)
    return t


def terrain_map(tc):
    """Terrain.__init__ + curiculum / randomized_terrain + add_terrain_to_map: (height_field_raw int16 [tot_rows, tot_cols],
    env_origins float64 [num_rows, num_cols, 3]).  `tc`: an object with the reference's cfg.terrain attribute names."""
    wpx = int(tc.terrain_width / tc.horizontal_scale)
    lpx = int(tc.terrain_length / tc.horizontal_scale)
    border = int(tc.border_size / tc.horizontal_scale)
    tot_cols = int(tc.num_cols * wpx) + 2 * border
    tot_rows = int(tc.num_rows * lpx) + 2 * border
    hf = np.zeros((tot_rows, tot_cols), dtype=np.int16)
    origins = np.zeros((tc.num_rows, tc.num_cols, 3))

    def add(t, i, j):
        hf[border + i * lpx:border + (i + 1) * lpx, border + j * wpx:border + (j + 1) * wpx] = t.height_field_raw
        x1, x2 = int((tc.terrain_length / 2.0 - 1) / t.horizontal_scale), int((tc.terrain_length / 2.0 + 1) / t.horizontal_scale)
        y1, y2 = int((tc.terrain_width / 2.0 - 1) / t.horizontal_scale), int((tc.terrain_width / 2.0 + 1) / t.horizontal_scale)
        origins[i, j] = [(i + 0.5) * tc.terrain_length, (j + 0.5) * tc.terrain_width, np.max(t.height_field_raw[x1:x2, y1:y2]) * t.vertical_scale]

    if tc.curriculum:
        for j in range(tc.num_cols):
            for i in range(tc.num_rows):
                add(make_terrain(tc, j / tc.num_cols + 0.001, i / tc.num_rows, wpx), i, j)
    else:
        for k in range(tc.num_rows * tc.num_cols):
            i, j = np.unravel_index(k, (tc.num_rows, tc.num_cols))
            choice = np.random.uniform(0, 1)
            difficulty = np.random.choice([0.25, 0.5, 0.75, 0.9])
            add(make_terrain(tc, choice, difficulty, wpx), i, j)
    return hf, origins
