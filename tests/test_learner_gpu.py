"""GPU parity tests of the learner half: the CUDA kernels (through the C ABI / Python mirror classes) against the CPU
oracle (oracle/learner_oracle.py, autograd as the differentiation oracle) on identical parameters, data and random draws.

Tolerance (north_star): 1e-5 relative fp32, applied as |a-b| <= 1e-5 * max(|ref|, scale) with `scale` the tensor's own
magnitude (max |ref|): GEMM summation order differs between cuBLAS-free SIMT kernels and the CPU's oneDNN/MKL kernels."""
import ctypes as C

import numpy as np
import pytest
import torch

import dtc_b200  # noqa: F401
from dtc_b200 import _lib as B
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(params=["simt", "tc3xtf32"])
def engine(request):
    """Runs a learner test on both GEMM engines: FP32 SIMT and tcgen05 3xTF32 (csrc/dtc_gemm_tc.cu)."""
    lib = B.lib()
    lib.dtc_set_gemm_mode(1 if request.param == "tc3xtf32" else 0)
    yield request.param
    lib.dtc_set_gemm_mode(1)


def _close(a, b, name, rel=1e-5, floor=0.0, flips=0.0):
    """flips: fraction of elements allowed outside the tolerance (each still within 5 % of the tensor scale).  Used for
    gradients only: a ReLU/ELU pre-activation within fp32 round-off of zero may land on either side on two correct
    implementations, which switches that unit's gradient contribution on or off."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = max(float(b.abs().max()), floor, 1e-30)
    err = (a - b).abs()
    tol = rel * torch.maximum(b.abs(), torch.full_like(b, scale))
    bad = err > tol
    if flips > 0 and bool(bad.any()) and float(err.max()) <= 0.05 * scale:
        # one flipped unit of one sample perturbs a whole row of its own weight gradient and, through dgrad, every element
        # of the upstream weight gradients by ~1e-3 of their scale: accept if few elements are off, or if the tensor as a
        # whole is within 2e-3 in the Frobenius norm
        if float(bad.double().mean()) <= flips or float((a - b).norm() / b.norm().clamp_min(1e-30)) <= 2e-3:
            return
    if bool(bad.any()):
        i = int((err / tol).argmax())
        raise AssertionError(f"{name}: |diff| {err.flatten()[i].item():.3e} > tol {tol.flatten()[i].item():.3e} at "
                             f"{np.unravel_index(i, tuple(a.shape))} (ref {b.flatten()[i].item():.6e}, got {a.flatten()[i].item():.6e}, "
                             f"scale {scale:.3e})")


# ------------------------------------------------------------------ GEMM family
def _lo(x):
    """3xTF32 companion: rn_tf32(x - trunc_tf32(x)) (csrc/dtc_common.cuh: tf32_lo)."""
    hi = (x.view(torch.int32) & -8192).view(torch.float32)
    r = x - hi
    return ((r.view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def _gemm(lib, mode, M, N, K, A, a_kc, Bm, b_kc, Cd, splits=1, ws=None, C_lo=None):
    A_lo = _lo(A) if mode == 1 else None
    B_lo = _lo(Bm) if mode == 1 else None
    B.check(lib.dtc_gemm_debug(M, N, K, B.ptr(A), B.ptr(A_lo), A.shape[1], a_kc, B.ptr(Bm), B.ptr(B_lo), Bm.shape[1], b_kc, B.ptr(Cd),
                               B.ptr(C_lo), Cd.shape[1], splits, B.ptr(ws), mode, B.stream_ptr()), "gemm")
    torch.cuda.synchronize()


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(4096, 512, 693), (300, 35, 64), (1000, 12, 128), (777, 693, 512), (64, 1, 128), (130, 588, 512),
                                   (256, 128, 265), (5000, 256, 12), (24576, 512, 693), (10000, 693, 512), (19000, 256, 512)])
def test_gemm_forward_and_dgrad(M, N, K, mode):
    """mode 0: FP32 SIMT kernels; mode 1: tcgen05 3xTF32 with companion arrays; mode 2: tcgen05 3xTF32, companions computed in shared
    memory by the splitter warps (shapes too small for a tile fall back to SIMT inside the launcher)."""
    lib = B.lib()
    g = torch.Generator().manual_seed(M + N + K)
    r4 = lambda x: (x + 3) // 4 * 4
    A = torch.zeros(M, r4(K)); A[:, :K] = torch.randn(M, K, generator=g)
    W = torch.zeros(N, r4(K)); W[:, :K] = torch.randn(N, K, generator=g) * 0.1
    bias = torch.randn(N, generator=g)
    ref0 = A[:, :K].double() @ W[:, :K].double().T
    Ad, Wd, bd = A.to(DEV), W.to(DEV), bias.to(DEV)
    if mode == 0:
        for act, f in ((0, lambda x: x), (1, torch.relu), (2, torch.nn.functional.elu)):
            Cd = torch.full((M, r4(N)), 7.0, device=DEV)
            B.check(lib.dtc_linear_forward(M, N, K, B.ptr(Ad), A.shape[1], B.ptr(Wd), W.shape[1], B.ptr(bd), act, B.ptr(Cd), Cd.shape[1],
                                           B.stream_ptr()), "fwd")
            _close(Cd[:, :N], f(ref0 + bias.double()), f"fwd act{act}")
            assert bool((Cd[:, N:] == 7.0).all()), "pad columns must not be written"
    Cd = torch.full((M, r4(N)), 7.0, device=DEV)
    Cl = torch.full((M, r4(N)), 7.0, device=DEV)
    _gemm(lib, mode, M, N, K, Ad, 1, Wd, 1, Cd, C_lo=Cl)
    _close(Cd[:, :N], ref0, f"fwd mode{mode}")
    # GemmArgs contract: the padding columns are left alone, or zeroed by the TMA-store epilogue (16-byte clipping granularity)
    assert bool(((Cd[:, N:] == 7.0) | (Cd[:, N:] == 0.0)).all()), "pad columns hold neither the old value nor zero"
    assert torch.equal(Cl[:, :N], _lo(Cd[:, :N].contiguous())), "companion output"
    # dgrad layout: C[M,K] = dY[M,N] @ W[N,K]  (A k-contiguous, B k-strided)
    dY = torch.zeros(M, r4(N)); dY[:, :N] = torch.randn(M, N, generator=g)
    dYd = dY.to(DEV)
    Cd = torch.zeros(M, r4(K), device=DEV)
    _gemm(lib, mode, M, K, N, dYd, 1, Wd, 0, Cd)
    _close(Cd[:, :K], dY[:, :N].double() @ W[:, :K].double(), f"dgrad mode{mode}")


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(512, 693, 24576), (35, 64, 4096), (12, 128, 3000), (1, 128, 2500), (53, 128, 999), (693, 512, 6144),
                                   (128, 268, 1000), (256, 588, 777), (512, 693, 1026), (512, 512, 1026), (256, 512, 1026)])
def test_gemm_wgrad_splitk(M, N, K, mode):
    """dW[M=out, N=in] = dY[K, out]^T X[K, in]; both operands k-strided; split-K through the workspace."""
    lib = B.lib()
    g = torch.Generator().manual_seed(M * 3 + N + K)
    r4 = lambda x: (x + 3) // 4 * 4
    dY = torch.zeros(K, r4(M)); dY[:, :M] = torch.randn(K, M, generator=g)
    X = torch.zeros(K, r4(N)); X[:, :N] = torch.randn(K, N, generator=g)
    ref = dY[:, :M].double().T @ X[:, :N].double()
    dYd, Xd = dY.to(DEV), X.to(DEV)
    for splits in (1, 7, 25):
        # workspace contract (csrc/dtc_gemm.cu: dtc_gemm_pick_splits): the tensor-core path keeps <= 32 k-blocks of 32 per split
        ws = torch.zeros(max(splits, -(-K // 1024)) * M * r4(N), device=DEV)
        Cd = torch.full((M, r4(N)), 3.0, device=DEV)
        _gemm(lib, mode, M, N, K, dYd, 0, Xd, 0, Cd, splits=splits, ws=ws)
        _close(Cd[:, :N], ref, f"wgrad mode{mode} splits={splits}")


@pytest.mark.parametrize("M,N,K,a_kc,b_kc,splits", [(24576, 512, 693, 1, 1, 1), (10000, 693, 512, 1, 0, 1), (24576, 752, 512, 1, 0, 1),
                                                     (512, 693, 24576, 0, 0, 25), (693, 512, 6144, 0, 0, 7), (256, 512, 24576, 0, 0, 30),
                                                     (19000, 256, 512, 1, 1, 1)])
def test_gemm_cta_pair_matches_single(M, N, K, a_kc, b_kc, splits):
    """The cta_group::2 kernel (256x128 pair tiles, B halves shared between the two SMs of a TPC) issues the same MMAs in the
    same order as the single-CTA kernel: results must be bit-identical, including ragged last tiles.  Split-K partial tiles are
    summed by the TMA unit's L2 reduction in the order the CTAs finish: equal to fp32 round-off there."""
    lib = B.lib()
    g = torch.Generator().manual_seed(M + 7 * N + K)
    r4 = lambda x: (x + 3) // 4 * 4
    A = torch.zeros(M if a_kc else K, r4(K if a_kc else M)); A[:, :(K if a_kc else M)] = torch.randn(A.shape[0], K if a_kc else M, generator=g)
    Bm = torch.zeros(N if b_kc else K, r4(K if b_kc else N)); Bm[:, :(K if b_kc else N)] = torch.randn(Bm.shape[0], K if b_kc else N, generator=g) * 0.1
    Ad, Bd = A.to(DEV), Bm.to(DEV)
    ws = torch.zeros(max(splits, -(-K // 1024)) * M * r4(N), device=DEV) if splits > 1 else None
    out = {}
    saved = lib.dtc_get_gemm_pair()
    try:
        for pair in (0, 1):
            lib.dtc_set_gemm_pair(pair)
            Cd = torch.full((M, r4(N)), 5.0, device=DEV)
            Cl = torch.full((M, r4(N)), 5.0, device=DEV)
            _gemm(lib, 1, M, N, K, Ad, a_kc, Bd, b_kc, Cd, splits=splits, ws=ws, C_lo=Cl if splits == 1 else None)
            out[pair] = (Cd.clone(), Cl.clone())
    finally:
        lib.dtc_set_gemm_pair(saved)
    Ar = (A[:, :K] if a_kc else A[:, :M].T).double()
    Br = (Bm[:, :K] if b_kc else Bm[:, :N].T).double()
    _close(out[1][0][:, :N], Ar @ Br.T, "pair vs fp64")
    if splits == 1:
        assert torch.equal(out[0][0], out[1][0]), f"pair kernel differs: {(out[0][0] - out[1][0]).abs().max().item()}"
        assert torch.equal(out[0][1], out[1][1]), "companion output differs"
    else:
        scale = float(out[0][0][:, :N].abs().max())
        assert float((out[0][0] - out[1][0]).abs().max()) <= 2e-6 * scale, "split-K: pair vs single beyond summation-order round-off"
        assert bool((out[1][0][:, N:] == 0).all() | (out[1][0][:, N:] == 5.0).all()), "pad columns: zeroed (L2-reduction path) or untouched"


# ------------------------------------------------------------------ policy forward
def _make_policies(seed, hot=True):
    """hot=True perturbs every parameter so that each one matters in single-step comparisons (the reference initialises
    later layers with gain 0.01 and zero bias); hot=False keeps the reference initialisation (trajectory tests)."""
    from oracle import learner_oracle as LO
    from dtc_b200.rsl_rl.modules import ActorCriticDecoder
    rng = H.TapRng(seed)
    torch.manual_seed(seed)
    oac = LO.ActorCriticDecoder(53, 1389, 12, rng=rng)
    with torch.no_grad():
        if not hot:
            pass
        else:
            g = torch.Generator().manual_seed(seed + 1)
            for k, p in oac.named_parameters():
                if k != "std":
                    p.add_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() == 2 else 0.02))
                else:
                    p.copy_(0.5 + torch.rand(12, generator=g))
    cac = ActorCriticDecoder(53, 1389, 12).to(DEV)
    cac.load_state_dict(oac.state_dict())
    return oac, cac, rng


def _inputs(M, seed):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(M, 53, generator=g)
    hist = torch.randn(M, 265, generator=g)
    priv = torch.randn(M, 1389, generator=g)
    bv = torch.randn(M, 3, generator=g)
    return obs, hist, priv, bv


def test_state_dict_roundtrip():
    oac, cac, _ = _make_policies(1)
    sd = cac.state_dict()
    for k, v in oac.state_dict().items():
        assert torch.equal(sd[k].cpu(), v), k
    assert list(sd.keys()) == list(oac.state_dict().keys())


@pytest.mark.parametrize("M", [64, 1000, 4096])
def test_act_parity(M, engine):
    oac, cac, rng = _make_policies(2)
    obs, hist, priv, bv = _inputs(M, 5)
    with torch.no_grad():
        a_ref = oac.act(obs, hist, priv)
        v_ref = oac.evaluate(obs, priv, bv)
        lp_ref = oac.get_actions_log_prob(a_ref)
        mu_ref, sg_ref = oac.action_mean, oac.action_std
    log = rng.take()
    cac._inject = dict(eps_z=log[0][1].to(DEV).contiguous(), eps_a=log[1][1].to(DEV).contiguous())
    o = cac._forward_act(obs.to(DEV), hist.to(DEV), priv.to(DEV), bv.to(DEV))
    _close(cac.debug_buffer("ML")[:, :19], oac.latent_mu, "latent_mu")
    _close(cac.debug_buffer("ML")[:, 19:35], oac.latent_var, "latent_var (after outlier repair)")
    _close(cac.debug_buffer("XA")[:, 568:584], oac.z, "z")
    _close(o["mean"], mu_ref, "action mean")
    _close(o["sigma"], sg_ref, "action std")
    _close(o["actions"], a_ref, "actions")
    _close(o["values"], v_ref.squeeze(1), "values")
    _close(o["logp"], lp_ref, "log prob", floor=1.0)
    _close(cac.evaluate(obs.to(DEV), priv.to(DEV), bv.to(DEV)), v_ref, "evaluate")
    with torch.no_grad():
        t_ref = oac.act_teacher(obs, hist, priv)
    _close(cac.act_teacher(obs.to(DEV), hist.to(DEV), priv.to(DEV)), t_ref, "act_teacher")


def test_philox_path_runs_and_is_standard_normal():
    _, cac, _ = _make_policies(3)
    obs, hist, priv, bv = _inputs(4096, 6)
    o = {k: v.clone() for k, v in cac._forward_act(obs.to(DEV), hist.to(DEV), priv.to(DEV), bv.to(DEV)).items()}
    eps = (o["actions"] - o["mean"]) / o["sigma"]
    assert abs(float(eps.mean())) < 0.02 and abs(float(eps.std()) - 1.0) < 0.02
    z = cac.debug_buffer("EPS")
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    o2 = cac._forward_act(obs.to(DEV), hist.to(DEV), priv.to(DEV), bv.to(DEV))
    assert not torch.equal(o2["actions"], o["actions"]), "a new call must draw new noise"


# ------------------------------------------------------------------ storage, GAE, update
def _fill_storages(N, T, seed, oac, cac, rng):
    """An oracle PPO and a CUDA PPO holding identical synthetic rollouts."""
    from oracle import learner_oracle as LO
    from dtc_b200.rsl_rl.algorithms import PPO
    kw = dict(num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
              entropy_coef=0.003, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
              desired_kl=0.01)
    oalg = LO.PPO(oac, rng=rng, **kw)
    oalg.init_storage(N, T, [53], [1389], [265], [12])
    calg = PPO(cac, device=DEV, **kw)
    calg.init_storage(N, T, [53], [1389], [265], [12])
    g = torch.Generator().manual_seed(seed)
    for t in range(T):
        obs, hist, priv, bv = _inputs(N, seed * 100 + t)
        with torch.no_grad():
            a = oalg.act(obs, priv, hist, bv)
        log = rng.take()
        cac._inject = dict(eps_z=log[0][1].to(DEV).contiguous(), eps_a=log[1][1].to(DEV).contiguous())
        ca = calg.act(obs.to(DEV), priv.to(DEV), hist.to(DEV), bv.to(DEV))
        _close(ca, a, f"rollout actions t={t}")
        rew = torch.randn(N, generator=g)
        dones = (torch.rand(N, generator=g) < 0.1)
        touts = dones & (torch.rand(N, generator=g) < 0.5)
        nobs = torch.randn(N, 53, generator=g)
        oalg.process_env_step(rew, dones, nobs, {"time_outs": touts})
        calg.process_env_step(rew.to(DEV), dones.to(DEV).to(torch.uint8), nobs.to(DEV), {"time_outs": touts.to(DEV)})
    obs, hist, priv, bv = _inputs(N, seed * 100 + T)
    with torch.no_grad():
        oalg.compute_returns(obs, priv, bv)
    calg.compute_returns(obs.to(DEV), priv.to(DEV), bv.to(DEV))
    return oalg, calg


def test_storage_and_gae_parity(engine):
    oac, cac, rng = _make_policies(4)
    oalg, calg = _fill_storages(32, 24, 7, oac, cac, rng)
    so, sc = oalg.storage, calg.storage
    for name in ("observations", "next_observations", "privileged_observations", "observation_histories", "actions", "base_vel"):
        _close(getattr(sc, name), getattr(so, name), name)
    assert torch.equal(sc.dones.cpu(), so.dones)
    _close(sc.rewards, so.rewards, "rewards (timeout bootstrap)")
    _close(sc.values, so.values, "values")
    _close(sc.actions_log_prob, so.actions_log_prob, "logp", floor=1.0)
    _close(sc.mu, so.mu, "mu")
    _close(sc.sigma, so.sigma, "sigma")
    _close(sc.returns, so.returns, "returns")
    _close(sc.advantages, so.advantages, "advantages", rel=2e-5)


def _grads_as_state_dict(cac):
    out = {}
    for k, idx in cac._idx.items():
        out[k] = cac._grads[idx].view(cac._table.shape[k]).clone()
    return out


@pytest.mark.parametrize("N", [16, 171])
def test_vae_step_gradients(N, engine):
    """Raw gradients and loss values of one VAE step (sync_grads=1) against autograd, hot parameters."""
    oac, cac, rng = _make_policies(5)
    T = 24
    oalg, calg = _fill_storages(N, T, 9, oac, cac, rng)
    mbs = N * T // 4
    lib = B.lib()
    oalg.debug = {}
    oalg.update()
    log = rng.take()
    perm = log[0][1]
    eps = [log[1][1].to(DEV).contiguous()]
    batch = calg.storage.gather(perm.to(DEV))
    hp = calg._hparams()
    h = cac._learner(mbs)
    calg._push_lr(h)
    B.check(lib.dtc_learner_reset_stats(h, B.stream_ptr()), "reset")
    B.check(lib.dtc_vae_step(h, C.byref(batch._c), 0, mbs, B.ptr(eps[0]), 0, 0, C.byref(hp), 1, B.stream_ptr()), "vae_step")
    got = _grads_as_state_dict(cac)
    for k, g_ref in oalg.debug["vae_grads"][0].items():
        _close(got["vae." + k], g_ref, "vae grad " + k, rel=2e-5, flips=1e-2)
    s = cac.stats().tolist()
    rec, vel, kld, hgt = oalg.debug["vae_losses"][0]
    assert s[2] == pytest.approx(rec, rel=1e-5) and s[3] == pytest.approx(vel, rel=1e-5)
    assert s[4] == pytest.approx(kld, rel=1e-5, abs=1e-7) and s[5] == pytest.approx(hgt, rel=1e-5)


@pytest.mark.parametrize("N,lv_bias", [(16, 0.0), (171, -1.0)])
def test_update_parity(N, lv_bias, engine):
    """Full PPO.update() from the reference initialisation: losses, learning-rate schedule and the parameters after
    2 epochs x 4 minibatches (16 Adam steps).
    lv_bias: at the reference initialisation logvar ~ 1e-5, where the reference's own KL gradient (1 - exp(logvar)) is a
    catastrophic cancellation whose fp32 value is ~1 % rounding noise of the exp() implementation (CPU Sleef vs CUDA);
    the larger case shifts latent_var.bias to -1 on both sides so that the comparison is about the kernels, not that noise."""
    oac, cac, rng = _make_policies(5, hot=False)
    if lv_bias != 0.0:
        with torch.no_grad():
            oac.vae.latent_var.bias.fill_(lv_bias)
        cac.load_state_dict(oac.state_dict())
    T = 24
    oalg, calg = _fill_storages(N, T, 9, oac, cac, rng)
    mbs = N * T // 4
    lib = B.lib()
    # --- the oracle's update with draw logging
    oalg.debug = {}
    p_before = {k: v.clone() for k, v in oac.state_dict().items()}
    o_ret = oalg.update()
    log = rng.take()
    assert log[0][0] == "randperm"
    perm = log[0][1]
    draws = [v for tag, v in log[1:]]
    assert len(draws) == 3 * 8
    eps = []
    for k in range(8):
        eps += [draws[3 * k].to(DEV).contiguous(), draws[3 * k + 1].to(DEV).contiguous()]
    hp = calg._hparams()
    # --- now the real update from the same starting point
    cac.load_state_dict(p_before)
    calg._inject = dict(perm=perm, eps=eps)
    c_ret = calg.update()
    vl, sl, ent, klm, lr = oalg.debug["ppo_losses"][-1]
    assert calg.learning_rate == pytest.approx(oalg.learning_rate, rel=1e-9), "adaptive-KL schedule must take the same branches"
    # Trajectory-level comparison.  One optimizer step is compared tightly elsewhere (test_vae_step_gradients,
    # test_policy_step_gradients at 2e-5, test_adam_and_clip_match_torch at 2e-7).  Over 16 chained steps two effects
    # legitimately amplify fp32 round-off: Adam (eps 1e-8) turns every gradient element into a step of ~lr whatever its size,
    # so elements whose gradient is ~0 take steps of either sign; and the latent_var outlier repair is discontinuous at its
    # 2-sigma threshold.  Both CUDA engines (SIMT, 3xTF32) land on the same trajectory to ~5 digits; against the CPU oracle the
    # chained comparison is a sanity bound: loss means to 5 %, mean parameter difference <= 5 % of the largest movement,
    # <= 15 % of elements off by more than 5 % of it.
    problems = []
    for a, b, name in zip(c_ret, o_ret, ("value", "surrogate", "adaptation", "decoder", "recons", "vel", "kld")):
        if not a == pytest.approx(b, rel=5e-2, abs=5e-4):
            problems.append((name, a, b))
    sd = cac.state_dict()
    for k, v in oac.state_dict().items():
        moved = (v - p_before[k]).abs().max().item()
        diff = (sd[k].cpu() - v).abs()
        if moved == 0:
            if diff.max().item() != 0:
                problems.append((k, "moved although the reference did not"))
            continue
        stats = (k, round(diff.mean().item() / moved, 5), round((diff > 0.05 * moved).float().mean().item(), 5),
                 round(diff.max().item() / moved, 4))
        if stats[1] > 5e-2 or stats[2] > 0.15 or stats[3] > 2.0:
            problems.append(stats)
    assert not problems, problems


@pytest.mark.parametrize("which", [0, 1])
def test_adam_and_clip_match_torch(which):
    """dtc_optimizer_apply (global-norm clip + Adam + adaptive-KL learning rate) against torch.nn.utils.clip_grad_norm_ +
    torch.optim.Adam fed the SAME gradients, 6 steps - isolates the optimizer from gradient round-off (see
    test_update_parity for why end-to-end trajectories can only be compared statistically)."""
    from dtc_b200.rsl_rl.modules import ActorCriticDecoder
    from dtc_b200.rsl_rl.algorithms import PPO
    torch.manual_seed(0)
    cac = ActorCriticDecoder(53, 1389, 12).to(DEV)
    calg = PPO(cac, learning_rate=1e-3, schedule="adaptive", desired_kl=0.01, device=DEV)
    lib, st = B.lib(), B.stream_ptr()
    h = cac._learner(64)
    calg._push_lr(h)
    hp = calg._hparams()
    b0, b1 = cac._table.ranges["vae" if which == 0 else "policy"]
    piggy = cac._table.ranges["policy_sync"][1] - 4
    p_ref = torch.nn.Parameter(cac._flat[b0:b1].cpu().clone())
    lr = 5e-4 if which == 0 else 1e-3
    opt = torch.optim.Adam([p_ref], lr=lr)
    g = torch.Generator().manual_seed(3)
    rows = 1000
    for step in range(6):
        scale = [3e-4, 1e-2, 1e-6, 5e-3, 1e-3, 2e-2][step]  # norms below and above max_grad_norm = 1
        grad = torch.randn(b1 - b0, generator=g) * scale
        grad[::7] = 0.0
        cac._grads[b0:b1] = grad.to(DEV)
        kl_mean = [0.05, 0.001, 0.01, 0.0, 0.03, 0.004][step]
        cac._grads[piggy] = kl_mean * rows
        B.check(lib.dtc_optimizer_apply(h, which, C.byref(hp), 1.0, rows, st), "apply")
        if which == 1:
            if kl_mean > 0.02:
                lr = max(1e-5, lr / 1.5)
            elif kl_mean < 0.005 and kl_mean > 0.0:
                lr = min(1e-2, lr * 1.5)
            for grp in opt.param_groups:
                grp["lr"] = lr
        p_ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        got = cac._flat[b0:b1].cpu()
        d = (got - p_ref.detach()).abs().max().item()
        assert d <= 2e-7, (step, d)
        if which == 1:
            assert float(cac.stats()[8]) == pytest.approx(lr, rel=1e-12)
    if which == 1:
        assert calg.learning_rate == pytest.approx(lr, rel=1e-12)


def test_checkpoint_interchange_with_torch_adam(tmp_path):
    """a `model_*.pt` as the reference's OnPolicyRunner.save writes it (on_policy_runner.py:249-255: state_dict of
    the nn.Module + state_dict of torch.optim.Adam over actor_critic.parameters()) resumes on the CUDA path, and the CUDA path
    exports the same layout.  Three optimizer steps on identical gradients on both sides, checkpoint written by the torch side,
    resumed into a fresh CUDA learner, one more step everywhere."""
    from dtc_b200.rsl_rl.modules import ActorCriticDecoder
    from dtc_b200.rsl_rl.modules.actor_critic_decoder import STATE_KEYS
    from dtc_b200.rsl_rl.algorithms import PPO
    oac, cac, _ = _make_policies(21)
    assert [k for k, _ in oac.named_parameters()] == list(STATE_KEYS)
    opt = torch.optim.Adam(oac.parameters(), lr=1e-3)
    calg = PPO(cac, learning_rate=1e-3, schedule="fixed", device=DEV)
    lib, st = B.lib(), B.stream_ptr()
    b0, b1 = cac._table.ranges["policy"]
    piggy = cac._table.ranges["policy_sync"][1] - 4
    g = torch.Generator().manual_seed(5)
    covered = torch.zeros(cac._flat.numel())
    for k in STATE_KEYS:
        covered[cac._table.index[k]] = 1.0  # zero-pad columns of the flat layout never see a gradient in training

    def step(algs, seed_scale):
        grad = torch.randn(b1 - b0, generator=g) * seed_scale * covered[b0:b1]  # global norm < 1: the clip does not engage
        for alg in algs:
            ac = alg.actor_critic
            h = ac._learner(64)
            alg._push_lr(h)
            if alg._pending_steps:
                a, b = C.c_int64(), C.c_int64()
                lib.dtc_learner_get_adam_steps(h, C.byref(a), C.byref(b))
                lib.dtc_learner_set_adam_steps(h, alg._pending_steps.get("vae", a.value), alg._pending_steps.get("main", b.value))
                alg._pending_steps = {}
            ac._grads.zero_()
            ac._grads[b0:b1] = grad.to(DEV)
            ac._grads[piggy] = 0.01 * 1000
            hp = alg._hparams()
            B.check(lib.dtc_optimizer_apply(h, 1, C.byref(hp), 1.0, 1000, st), "apply")
        per_key = _grads_as_state_dict(algs[0].actor_critic)
        live = {k for k in STATE_KEYS if b0 <= int(cac._table.index[k].min()) and int(cac._table.index[k].max()) < b1}
        for k, p in oac.named_parameters():
            p.grad = per_key[k].cpu() if k in live else None
        opt.step()
        return live

    for i in range(3):
        live = step([calg], 1e-4)
    ours, theirs = calg.optimizer.state_dict(), opt.state_dict()
    assert ours["param_groups"][0]["params"] == theirs["param_groups"][0]["params"]
    assert set(ours["state"].keys()) == set(theirs["state"].keys()) and len(ours["state"]) == len(live) == 31
    for i, stt in theirs["state"].items():
        assert float(ours["state"][i]["step"]) == float(stt["step"]) == 3.0
        for key in ("exp_avg", "exp_avg_sq"):
            a, b_ = ours["state"][i][key].cpu(), stt[key]
            assert a.shape == b_.shape
            # fp32 moment updates: fma / rounding order differs from torch's foreach kernels by a few ulp of the largest entries
            assert (a - b_).abs().max().item() <= 1e-5 * b_.abs().max().item() + 1e-30, (STATE_KEYS[i], key, (a - b_).abs().max())
    # the reference side writes the checkpoint; a fresh CUDA learner resumes from it
    path = str(tmp_path / "model_7.pt")
    torch.save({"model_state_dict": oac.state_dict(), "optimizer_state_dict": opt.state_dict(), "iter": 7, "infos": None}, path)
    loaded = torch.load(path, map_location="cpu", weights_only=True)
    cac2 = ActorCriticDecoder(53, 1389, 12).to(DEV)
    calg2 = PPO(cac2, learning_rate=3e-4, schedule="fixed", device=DEV)
    cac2.load_state_dict(loaded["model_state_dict"])
    calg2.optimizer.load_state_dict(loaded["optimizer_state_dict"])
    assert calg2.learning_rate == pytest.approx(1e-3)
    step([calg, calg2], 1e-4)
    p1, p2 = cac._flat[b0:b1].cpu(), cac2._flat[b0:b1].cpu()
    if (p1 - p2).abs().max().item() > 2e-7:
        sd1, sd2 = cac.state_dict(), cac2.state_dict()
        o1, o2 = calg.optimizer.state_dict()["state"], calg2.optimizer.state_dict()["state"]
        rep = []
        for i, k in enumerate(STATE_KEYS):
            d = (sd1[k] - sd2[k]).abs().max().item()
            if d > 2e-7:
                rep.append((k, d, (o1[i]["exp_avg"] - o2[i]["exp_avg"]).abs().max().item() if i in o1 else None,
                            (o1[i]["exp_avg_sq"] - o2[i]["exp_avg_sq"]).abs().max().item() if i in o1 else None))
        raise AssertionError(f"resumed learner diverges from the uninterrupted one: {rep}")
    sd = cac2.state_dict()
    for k, p in oac.named_parameters():
        assert torch.allclose(sd[k].cpu(), p.detach(), rtol=0, atol=3e-7), k
    assert float(calg2.optimizer.state_dict()["state"][0]["step"]) == 4.0
    # and the file this side writes loads into torch.optim.Adam over the reference module
    opt2 = torch.optim.Adam(oac.parameters(), lr=1.0)
    osd = calg2.optimizer.state_dict()
    osd["state"] = {i: {k: v.cpu() for k, v in s_.items()} for i, s_ in osd["state"].items()}
    opt2.load_state_dict(osd)
    assert opt2.param_groups[0]["lr"] == pytest.approx(1e-3)


def test_policy_step_gradients(engine):
    """Raw gradients of one policy step (sync_grads=1) against autograd, incl. the outlier->median gradient routing."""
    oac, cac, rng = _make_policies(6)
    N, T = 64, 24
    oalg, calg = _fill_storages(N, T, 11, oac, cac, rng)
    mbs = N * T // 4
    lib = B.lib()
    from oracle import learner_oracle as LO
    st = oalg.storage
    perm = torch.randperm(N * T, generator=torch.Generator().manual_seed(1))
    b = perm[:mbs]
    f = {k: getattr(st, k).flatten(0, 1) for k in LO.RolloutStorage.FIELDS}
    # oracle: the policy half of PPO.update on minibatch 0, by hand
    oac.zero_grad()
    oac.act(f["observations"][b], f["observation_histories"][b], f["privileged_observations"][b])
    log = rng.take()
    logp = oac.get_actions_log_prob(f["actions"][b])
    value = oac.evaluate(f["observations"][b], f["privileged_observations"][b], f["base_vel"][b])
    ratio = torch.exp(logp - f["actions_log_prob"][b].squeeze())
    adv = f["advantages"][b].squeeze()
    surr = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2)).mean()
    tv, ret = f["values"][b], f["returns"][b]
    vc = tv + (value - tv).clamp(-0.2, 0.2)
    vloss = torch.max((value - ret).pow(2), (vc - ret).pow(2)).mean()
    loss = surr + 1.0 * vloss - 0.003 * oac.entropy.mean()
    loss.backward()
    batch = calg.storage.gather(perm.to(DEV))
    hp = calg._hparams()
    h = cac._learner(mbs)
    B.check(lib.dtc_ppo_step(h, C.byref(batch._c), 0, mbs, B.ptr(log[0][1].to(DEV).contiguous()), 0, 0, C.byref(hp), 1,
                             B.stream_ptr()), "ppo_step")
    got = _grads_as_state_dict(cac)
    n_checked = 0
    for k, p in oac.named_parameters():
        if p.grad is not None:
            _close(got[k], p.grad, "policy grad " + k, rel=3e-5, flips=1e-2)
            n_checked += 1
    assert n_checked == 31
    n_out = int(cac.debug_buffer("OUTM").view(torch.uint8)[:, :16].sum())
    assert n_out > 0, "the outlier path must be exercised"


def test_step_gradients_at_bench_size():
    """P11 / P12 at the benchmark's minibatch size (N=4096, T=24 -> 24 576 rows): the radix-select median over 393 k latent_var
    elements, the 24-25-way split-K weight gradients and the CTA-pair GEMM schedule only exist at this size.  One VAE step and
    one policy step (sync_grads=1: backward only) on a permuted minibatch of a 98 304-row storage against autograd on the same
    rows, parameters and draws; default engine (tcgen05 3xTF32)."""
    from dtc_b200.rsl_rl.algorithms import PPO
    from dtc_b200.rsl_rl.storage import RolloutStorage
    oac, cac, rng = _make_policies(8)
    N, T = 4096, 24
    mbs = N * T // 4
    g = torch.Generator().manual_seed(77)
    perm = torch.randperm(N * T, generator=g)
    b = perm[:mbs]
    R = N * T
    f = dict(obs=torch.randn(R, 53, generator=g), hist=torch.randn(R, 265, generator=g), priv=torch.randn(R, 1389, generator=g),
             bv=torch.randn(R, 3, generator=g), nobs=torch.randn(R, 53, generator=g), adv=torch.randn(R, generator=g),
             actions=torch.zeros(R, 12), values=torch.zeros(R), returns=torch.zeros(R), logp=torch.zeros(R), mu=torch.zeros(R, 12),
             sigma=torch.ones(R, 12))
    # old policy outputs on the minibatch rows: the current policy's, perturbed, so that ratio / value clipping are exercised
    with torch.no_grad():
        a = oac.act(f["obs"][b], f["hist"][b], f["priv"][b])
        rng.take()
        v = oac.evaluate(f["obs"][b], f["priv"][b], f["bv"][b]).squeeze(1)
        f["actions"][b] = a
        f["mu"][b] = oac.action_mean + 0.05 * torch.randn(mbs, 12, generator=g)
        f["sigma"][b] = oac.action_std
        f["logp"][b] = oac.get_actions_log_prob(a) + 0.15 * torch.randn(mbs, generator=g)
        f["values"][b] = v + 0.3 * torch.randn(mbs, generator=g)
        f["returns"][b] = v + 0.5 * torch.randn(mbs, generator=g)
    calg = PPO(cac, num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0,
               entropy_coef=0.003, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
               desired_kl=0.01, device=DEV)
    calg.init_storage(N, T, [53], [1389], [265], [12])
    st = calg.storage
    tr = RolloutStorage.Transition()
    for t in range(T):
        sl = slice(t * N, (t + 1) * N)
        tr.observations, tr.observation_histories, tr.privileged_observations = (f[k][sl].to(DEV) for k in ("obs", "hist", "priv"))
        tr.base_vel, tr.next_observations, tr.actions = f["bv"][sl].to(DEV), f["nobs"][sl].to(DEV), f["actions"][sl].to(DEV)
        tr.rewards, tr.dones = torch.zeros(N, device=DEV), torch.zeros(N, device=DEV, dtype=torch.uint8)
        tr.values, tr.actions_log_prob = f["values"][sl].to(DEV), f["logp"][sl].to(DEV)
        tr.action_mean, tr.action_sigma = f["mu"][sl].to(DEV), f["sigma"][sl].to(DEV)
        st.add_transitions(tr)
    st.returns.copy_(f["returns"].view(T, N, 1))
    st.advantages.copy_(f["adv"].view(T, N, 1))
    eps_v = torch.randn(mbs, 16, generator=g)
    eps_p = torch.randn(mbs, 16, generator=g)
    lib, stream = B.lib(), B.stream_ptr()
    batch = st.gather(perm.to(DEV))
    hp, h = calg._hparams(), cac._learner(mbs)
    calg._push_lr(h)
    report = {}

    def compare(got, named, prefix, rel, tag):
        n = 0
        for k, gref in named:
            if gref is None:
                continue
            gc = got[prefix + k].double().cpu()
            report[tag + prefix + k] = (float((gc - gref.double()).abs().max() / gref.abs().max().clamp_min(1e-30)),
                                        float((gc - gref.double()).norm() / gref.double().norm().clamp_min(1e-30)))
            _close(got[prefix + k], gref, "grad " + prefix + k, rel=rel, flips=1e-2)
            n += 1
        return n

    # ---- VAE step (ppo.py:197-254)
    class _Fixed:
        def __init__(self, e): self.e = e
        def randn_like(self, t): return self.e
    oac.vae.rng = _Fixed(eps_v)
    oac.zero_grad()
    mu, lv, z = oac.vae.cenet_forward(f["hist"][b])
    l_t = oac.vae.terrain_encoder(f["priv"][b][:, :693])
    recons = oac.vae.cenet_decoder(torch.cat([z, mu[:, :3], l_t], dim=1))
    recons_loss = torch.pow(recons - f["nobs"][b], 2).mean(-1).mean()
    height_loss = torch.nn.functional.mse_loss(oac.vae.terrain_decoder(l_t), f["priv"][b][:, 696:])
    vel_loss = torch.nn.functional.mse_loss(mu[:, :3], f["bv"][b])
    kld_loss = torch.mean(-0.5 * torch.sum(1 + lv - mu[:, 3:].pow(2) - lv.exp(), dim=1))
    (recons_loss + vel_loss + 4 * kld_loss + height_loss).backward()
    B.check(lib.dtc_learner_reset_stats(h, stream), "reset")
    B.check(lib.dtc_vae_step(h, C.byref(batch._c), 0, mbs, B.ptr(eps_v.to(DEV).contiguous()), 0, 0, C.byref(hp), 1, stream), "vae_step")
    n_v = compare(_grads_as_state_dict(cac), [(k, p.grad) for k, p in oac.vae.named_parameters()], "vae.", 2e-5, "VAE:")
    s = cac.stats().tolist()
    assert s[2] == pytest.approx(recons_loss.item(), rel=1e-5) and s[3] == pytest.approx(vel_loss.item(), rel=1e-5)
    assert s[4] == pytest.approx(kld_loss.item(), rel=1e-5, abs=1e-7) and s[5] == pytest.approx(height_loss.item(), rel=1e-5)
    # ---- policy step (ppo.py:265-338)
    oac.vae.rng = _Fixed(eps_p)
    oac.rng = _Fixed(torch.zeros(mbs, 12))
    oac.zero_grad()
    oac.act(f["obs"][b], f["hist"][b], f["priv"][b])
    logp = oac.get_actions_log_prob(f["actions"][b])
    value = oac.evaluate(f["obs"][b], f["priv"][b], f["bv"][b])
    ratio = torch.exp(logp - f["logp"][b])
    adv = f["adv"][b]
    surr = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2)).mean()
    tv, ret = f["values"][b].unsqueeze(1), f["returns"][b].unsqueeze(1)
    vc = tv + (value - tv).clamp(-0.2, 0.2)
    vloss = torch.max((value - ret).pow(2), (vc - ret).pow(2)).mean()
    (surr + 1.0 * vloss - 0.003 * oac.entropy.mean()).backward()
    B.check(lib.dtc_ppo_step(h, C.byref(batch._c), 0, mbs, B.ptr(eps_p.to(DEV).contiguous()), 0, 0, C.byref(hp), 1, stream), "ppo_step")
    n_p = compare(_grads_as_state_dict(cac), [(k, p.grad) for k, p in oac.named_parameters()], "", 3e-5, "PPO:")
    assert n_v == 26 and n_p == 31, (n_v, n_p)
    s = cac.stats().tolist()
    assert s[0] == pytest.approx(vloss.item(), rel=2e-5) and s[1] == pytest.approx(surr.item(), rel=2e-5, abs=1e-6)
    worst = sorted(report.items(), key=lambda kv: -kv[1][0])[:6]
    print("[24576-row step] gradient error (max|d|/max|ref|, |d|_F/|ref|_F): worst " + ", ".join(f"{k} {v[0]:.1e}/{v[1]:.1e}" for k, v in worst))
    print("[24576-row step] median over tensors: %.1e / %.1e" % (sorted(v[0] for v in report.values())[len(report) // 2],
                                                                   sorted(v[1] for v in report.values())[len(report) // 2]))
    # 57 gradient tensors: the typical one agrees to < 1e-5 (measured 8e-6, both norms).  The 512-wide ReLU layers of the terrain
    # encoder see 12.6 M pre-activations per layer at this size; the dozen within fp32 round-off of zero land on either side on
    # two correct implementations, and one flipped unit of one sample moves that unit's weight-gradient row by 1/sqrt(24576) =
    # 6e-3 of its norm (measured worst tensor: 1.0e-2 max-norm, 9e-4 Frobenius) - the `flips` allowance of _close covers exactly that
    med = sorted(v[0] for v in report.values())[len(report) // 2]
    assert med <= 2e-5, med
    assert max(v[1] for v in report.values()) <= 2e-3
    assert int(cac.debug_buffer("OUTM").view(torch.uint8)[:, :16].sum()) > 0, "the outlier path must be exercised"
