"""Host-side helpers that need no GPU: the trajectory split / pad / unpad pair around the GRU `Memory`
(reference rsl_rl/rsl_rl/utils/utils.py:33-71) against a step-by-step Python construction, and the GRU parameter count."""
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
