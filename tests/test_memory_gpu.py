"""the GRU `Memory` (rsl_rl/rsl_rl/modules/actor_critic_decoder.py:584-614) against torch.nn.GRU - the
very module the reference wraps - on the same weights: inference mode with persistent hidden state and per-env reset,
batch mode over padded trajectories."""
import pytest
import torch

import dtc_b200  # noqa: F401

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pair(input_size, H, L, seed):
    from dtc_b200.rsl_rl.modules import Memory
    torch.manual_seed(seed)
    ref = torch.nn.GRU(input_size=input_size, hidden_size=H, num_layers=L)
    mem = Memory(input_size, type="gru", num_layers=L, hidden_size=H).to(DEV)
    mem.load_state_dict({"rnn." + k: v for k, v in ref.state_dict().items()})
    assert list(mem.state_dict().keys()) == ["rnn." + k for k in ref.state_dict().keys()]
    return ref, mem


@pytest.mark.parametrize("N,input_size,H,L", [(64, 53, 50, 2), (4096, 53, 50, 2), (7, 265, 50, 1), (1000, 16, 128, 3)])
def test_inference_mode_matches_nn_gru(N, input_size, H, L):
    ref, mem = _pair(input_size, H, L, N + H)
    g = torch.Generator().manual_seed(3)
    h_ref = None
    for t in range(6):
        x = torch.randn(N, input_size, generator=g)
        with torch.no_grad():
            o_ref, h_ref = ref(x.unsqueeze(0), h_ref)
        o = mem.forward(x.to(DEV))
        assert o.shape == (1, N, H)
        assert torch.allclose(o.cpu(), o_ref, rtol=1e-5, atol=2e-6), (t, (o.cpu() - o_ref).abs().max())
        assert torch.allclose(mem.hidden_states.cpu(), h_ref, rtol=1e-5, atol=2e-6)
        if t == 2:  # Memory.reset(dones): hidden_state[..., dones, :] = 0
            dones = torch.rand(N, generator=g) < 0.3
            dones[0] = True
            h_ref = h_ref.clone()
            h_ref[..., dones, :] = 0.0
            mem.reset(dones.to(DEV))
            assert torch.equal(mem.hidden_states.cpu()[:, dones], torch.zeros(L, int(dones.sum()), H))


def test_batch_mode_over_padded_trajectories():
    from dtc_b200.rsl_rl.utils import split_and_pad_trajectories, unpad_trajectories
    T, N, input_size, H, L = 24, 33, 53, 50, 2
    ref, mem = _pair(input_size, H, L, 11)
    g = torch.Generator().manual_seed(5)
    obs = torch.randn(T, N, input_size, generator=g)
    dones = (torch.rand(T, N, generator=g) < 0.1)
    padded, masks = split_and_pad_trajectories(obs, dones)
    h0 = torch.randn(L, padded.shape[1], H, generator=g) * 0.3
    with torch.no_grad():
        o_ref, _ = ref(padded, h0)
        o_ref = unpad_trajectories(o_ref, masks)
    o = mem.forward(padded.to(DEV), masks=masks.to(DEV), hidden_states=h0.to(DEV))
    assert o.shape == o_ref.shape
    assert torch.allclose(o.cpu(), o_ref, rtol=1e-5, atol=3e-6), (o.cpu() - o_ref).abs().max()
    with pytest.raises(ValueError):
        mem.forward(padded.to(DEV), masks=masks.to(DEV))


def test_errors_mirror_the_reference_surface():
    from dtc_b200 import _lib as B
    from dtc_b200.rsl_rl.modules import Memory
    with pytest.raises(B.DtcError):
        Memory(53, type="lstm")
    with pytest.raises(B.DtcError):
        Memory(53, type="gru", hidden_size=50).forward(torch.zeros(4, 53, device=DEV))
    m = Memory(53, type="gru", hidden_size=50).to(DEV)
    m.reset(torch.ones(4, dtype=torch.bool, device=DEV))  # no hidden state yet: no-op, as in the reference
    with pytest.raises(ValueError):
        m.forward(torch.zeros(4, 54, device=DEV))
