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
