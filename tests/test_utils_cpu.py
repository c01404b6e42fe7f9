"""Host-side helpers that need no GPU: the trajectory split / pad / unpad pair around the GRU `Memory`
(reference rsl_rl/rsl_rl/utils/utils.py:33-71) against a step-by-step Python construction, and the GRU parameter count."""
import numpy as np
import torch

import dtc_b200  # noqa: F401
from dtc_b200.rsl_rl.utils import split_and_pad_trajectories, unpad_trajectories


def _brute(x, dones):
    T, N = dones.shape
    trajs = []
    for n in range(N):
        cur = []
        for t in range(T):
            cur.append(x[t, n])
            if bool(dones[t, n]) or t == T - 1:
                trajs.append(torch.stack(cur))
                cur = []
    padded = torch.zeros(T, len(trajs), *x.shape[2:])
    masks = torch.zeros(T, len(trajs), dtype=torch.bool)
    for k, tr in enumerate(trajs):
        padded[:len(tr), k] = tr
        masks[:len(tr), k] = True
    return padded, masks


def test_split_pad_unpad_roundtrip():
    g = torch.Generator().manual_seed(0)
    for T, N, F, p in ((24, 33, 53, 0.1), (24, 5, 3, 0.5), (8, 1, 2, 0.0), (24, 64, 7, 0.02), (24, 3, 4, 0.9)):
        x = torch.randn(T, N, F, generator=g)
        d = torch.rand(T, N, generator=g) < p
        padded, masks = split_and_pad_trajectories(x, d)
        bp, bm = _brute(x, d)
        assert torch.equal(padded, bp) and torch.equal(masks, bm), (T, N, p)
        assert torch.equal(unpad_trajectories(padded, masks), x)


def test_gru_param_count_matches_torch():
    from dtc_b200 import _lib as B
    for inp, H, L in ((53, 50, 2), (265, 50, 1), (16, 128, 3)):
        ref = torch.nn.GRU(input_size=inp, hidden_size=H, num_layers=L)
        assert int(B.lib().dtc_gru_param_floats(inp, H, L)) == sum(p.numel() for p in ref.parameters())


def _terrain_cell(sub, tab, x, y, px):
    """Python statement of csrc/dtc_terrain.cu: terrain_cell (the closed form one device thread evaluates)."""
    t, a, b, c, ph = (int(v) for v in sub)
    h = 0
    if t == 1:
        P = a + b
        cx = x // P
        h = c
        if x - cx * P < a:
            y0 = int(tab[cx])
            if y < max(0, y0 - b) or (y >= y0 and (y - y0) % P < a):
                h = 0
    elif t == 2:
        m = min(x, px - 1 - x, y, px - 1 - y) // a
        T = (px - c - 1) // (2 * a) + 1 if px > c else 0
        h = b * min(m, T)
    elif t == 3:
        for r in range(a):
            q = tab[5 * r:5 * r + 5]
            if q[0] <= x < q[0] + q[2] and q[1] <= y < q[1] + q[3]:
                h = int(q[4])
    if ph > 0:
        cc = px // 2
        if cc - ph <= x < cc + ph and cc - ph <= y < cc + ph:
            h = 0
    return h


def test_terrain_description_reproduces_the_host_generator():
    """sim_stub.describe_terrain draws exactly what make_heightmap draws, and the per-cell closed form that the device
    rasteriser evaluates gives the host loops' heights - checked on 2000 random cells per sub-terrain."""
    import numpy as np
    from dtc_b200 import lite3 as L, sim_stub
    px, b = int(L.TERRAIN_LENGTH / L.HORIZONTAL_SCALE), int(L.BORDER_SIZE / L.HORIZONTAL_SCALE)
    for kind, seed in (("flat", 0), ("stones", 0), ("stones", 7), ("curriculum", 0), ("curriculum", 3)):
        hs, _ = sim_stub.make_heightmap(kind, seed)
        subs, tabs = sim_stub.describe_terrain(kind, seed)
        rng = np.random.default_rng(1)
        for s in range(L.NUM_ROWS * L.NUM_COLS):
            i, j = s // L.NUM_COLS, s % L.NUM_COLS
            for x, y in zip(rng.integers(0, px, 2000), rng.integers(0, px, 2000)):
                assert _terrain_cell(subs[s], tabs[s], int(x), int(y), px) == int(hs[b + i * px + x, b + j * px + y]), (kind, seed, s, x, y)
    assert not hs[:b].any() and not hs[:, :b].any()  # flat border


# ------------------------------------------------------------------ host logic of the drop-in boundary (no GPU)
def test_terrain_host_layout_reproduces_reference_maps(golden_dir):
    """N3 on the CPU: the product `Terrain` class (same numpy draws, generators recording rectangle lists) against the heightmaps
    recorded from the unmodified reference class - the rectangles are painted here with numpy, as `dtc_terrain_paint` does on the
    device (last rectangle wins)."""
    import os
    from dtc_b200.legged_gym.utils.terrain import Terrain
    from tests.test_oracle_golden import TERRAIN_CASES, terrain_cfg
    for name, (seed, ov) in sorted(TERRAIN_CASES.items()):
        G = np.load(os.path.join(golden_dir, f"terrain_{name}.npz"))
        cfg = terrain_cfg(ov)
        np.random.seed(seed)
        t = Terrain(cfg, 16, device=None)  # layout only
        hf = np.zeros((t.tot_rows, t.tot_cols), dtype=np.int16)
        lp, wp, b = t.length_per_env_pixels, t.width_per_env_pixels, t.border
        subs = t.sub_terrains
        assert len(subs) == cfg.num_rows * cfg.num_cols
        for s, (background, rects) in enumerate(subs):
            i, j = divmod(s, cfg.num_cols)
            tile = np.full((lp, wp), background, dtype=np.int16)
            for x0, x1, y0, y1, h in rects:
                tile[x0:x1, y0:y1] = h
            hf[b + i * lp:b + (i + 1) * lp, b + j * wp:b + (j + 1) * wp] = tile
        assert np.array_equal(hf, G["height_field_raw"]), (name, int((hf != G["height_field_raw"]).sum()))


def test_cfg_resolution_follows_the_reference_rules():
    """legged_gym/envs/base/cfg_resolve.py against the values LeggedRobot._parse_cfg / _prepare_reward_function / _init_buffers derive
    from the Lite3 DTC config (legged_robot.py:929-952, 839-866, 1098-1109), and the loud failures for what the kernels cannot do."""
    import copy
    import pytest as _pt
    from dtc_b200.legged_gym.envs import Lite3DTCCfg
    from dtc_b200.legged_gym.envs.base.cfg_resolve import CfgError, resolve
    from dtc_b200 import lite3 as L
    r = resolve(Lite3DTCCfg())
    assert abs(r.dt - 4 * 0.005) < 1e-12 and r.max_episode_length == int(np.ceil(20.0 / r.dt))
    assert r.resampling_steps == int(10.0 / r.dt) and r.push_interval == int(np.ceil(15.0 / r.dt))
    assert len(r.p_gains) == 12 and len(r.d_gains) == 12 and all(p > 0 for p in r.p_gains)
    # reward scales are multiplied by dt and zero scales drop out of the episode sums
    ref = {k: v * r.dt for k, v in L.REWARD_SCALES.items() if v != 0.0}
    assert set(r.reward_scales) == set(ref)
    assert all(abs(r.reward_scales[k] - ref[k]) <= 1e-12 * max(1.0, abs(ref[k])) for k in ref)
    # edits are honoured ...
    cfg = Lite3DTCCfg()
    cfg.rewards = copy.deepcopy(cfg.rewards)

    class _S(cfg.rewards.scales):
        soft_tracking_lin_vel = 2.5
        torques = 0.0

    cfg.rewards.scales = _S
    cfg.commands = copy.deepcopy(cfg.commands)

    class _R(cfg.commands.ranges):
        lin_vel_x = [-0.3, 0.9]

    cfg.commands.ranges = _R
    r2 = resolve(cfg)
    assert abs(r2.reward_scales["soft_tracking_lin_vel"] - 2.5 * r2.dt) < 1e-12 and "torques" not in r2.reward_scales
    assert r2.command_ranges["lin_vel_x"] == [-0.3, 0.9]
    # ... and what the kernels cannot honour fails loudly
    for path, value in (("control.control_type", "T"), ("control.decimation", 2), ("commands.heading_command", False),
                        ("env.num_observations", 48), ("terrain.measure_heights", False), ("rewards.only_positive_rewards", True)):
        bad = Lite3DTCCfg()
        section, field = path.split(".")
        sec = copy.deepcopy(getattr(bad, section))
        sub = type("edited", (sec if isinstance(sec, type) else type(sec),), {field: value})
        setattr(bad, section, sub)
        with _pt.raises(CfgError):
            resolve(bad)


def test_get_load_path_picks_the_last_run_and_highest_checkpoint(tmp_path):
    """legged_gym/utils/helpers.py:73-95."""
    import pytest as _pt
    from dtc_b200.legged_gym.utils.helpers import get_load_path
    with _pt.raises(ValueError):
        get_load_path(str(tmp_path / "nothing_here"))
    for run, models in (("Jan01_10-00-00_a", (0, 50, 100)), ("Jan02_09-00-00_b", (0, 50, 1500, 200)), ("exported", ())):
        d = tmp_path / run
        d.mkdir()
        for m in models:
            (d / f"model_{m}.pt").write_bytes(b"")
    assert get_load_path(str(tmp_path)).endswith("Jan02_09-00-00_b/model_1500.pt")          # 'exported' ignored, numeric order
    assert get_load_path(str(tmp_path), load_run="Jan01_10-00-00_a").endswith("Jan01_10-00-00_a/model_100.pt")
    assert get_load_path(str(tmp_path), load_run="Jan01_10-00-00_a", checkpoint=50).endswith("Jan01_10-00-00_a/model_50.pt")
