"""Pins the CPU oracle (oracle/*.py) against golden vectors recorded from the unmodified reference
(tests/golden/make_golden.py).  Floats: rtol/atol 1e-6 (see oracle/env_oracle.py arithmetic policy);
indices: exact, except that an index may differ only where the two candidates' scores differ < 1e-6."""
import os

import numpy as np
import pytest
import torch

import dtc_b200  # noqa: F401
from dtc_b200 import lite3 as K, sim_stub
from oracle import env_oracle as EO, learner_oracle as LO
from oracle.rng import Replay

TOL = dict(rtol=2e-6, atol=2e-6)


def _close(a, b, name, **kw):
    kw = {**TOL, **kw}
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if not torch.allclose(a.float(), b.float(), **kw):
        d = (a.float() - b.float()).abs()
        i = d.argmax()
        raise AssertionError(f"{name}: max abs diff {d.max().item():.3e} at {np.unravel_index(i, d.shape)} "
                             f"(ref {b.flatten()[i].item()}, got {a.flatten()[i].item()})")


@pytest.fixture(scope="module")
def env_gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "env_n16.pt"), weights_only=False)


def _make_oracle_env(G, rng):
    hs, tor = sim_stub.make_heightmap(*G["heightmap"])
    fg = sim_stub.FakeGym(G["N"])
    env = EO.OracleEnv(K, G["N"], hs, G["layout"], fg, rng)
    return env, fg


def test_env_step_matches_reference(env_gold):
    G = env_gold
    rng = Replay(G["reset_log"])
    env, fg = _make_oracle_env(G, rng)
    fg.queue.append(G["states"][0])
    env.reset()
    assert rng.done()
    a = G["after_reset"]
    _close(env.obs_buf, a["obs"], "reset obs")
    _close(env.commands, a["commands"], "reset commands")
    _close(env.rew_buf, a["rew"], "reset rew", atol=1e-5)
    env.episode_length_buf[0:4] = 498
    env.episode_length_buf[4:6] = 999
    env.common_step_counter = 747
    n_idx_diff = n_nom_diff = n_fallback = n_exact_tie = 0
    for t, fr in enumerate(G["frames"]):
        rng = env.rng = Replay(fr["log"])
        fg.queue.append(G["states"][t + 1])
        obs, priv, rew, done, extras = env.step(fr["actions"])
        assert rng.done(), f"step {t}: oracle consumed {rng.pos} of {len(rng.log)} draws"
        tag = f"step{t} "
        assert torch.equal(env.measured_heights, fr["measured_heights"]), tag + "measured_heights must be bit-exact"
        _close(env.pred_footholds, fr["pred_footholds"], tag + "pred_footholds")
        _close(env.base_lin_vel, fr["base_lin_vel"], tag + "base_lin_vel")
        _close(env.commands, fr["commands"], tag + "commands")
        _close(env.torques, fr["torques"], tag + "torques", atol=1e-5)
        _close(env.measured_foot_clearance, fr["clearance"], tag + "clearance")
        # indices: exact; a difference is tolerated only where the two candidates tie to < 1e-6 in the reference's own score
        # (optimal) or distance (nominal) - the reference's fp32 cascade mean/var is not reproduced bit for bit (module docstring)
        mine, ref = env.optimal_foothold_indice.squeeze(1), fr["optimal_idx"]
        for n, l in (mine != ref).nonzero().tolist():
            sc = fr["foothold_score"][n, :, l]
            assert abs(sc[mine[n, l]] - sc[ref[n, l]]) < 1e-6, (tag, "optimal", n, l)
            n_idx_diff += 1
        mine_n, ref_n = env.nominal_footholds_indice, fr["nominal_idx"]
        for n, l in (mine_n != ref_n).nonzero().tolist():
            xy = env.heights_world[n, :, :2]
            d = (env.pred_footholds[n, l, :2][None] - xy).norm(dim=1)
            assert abs(d[mine_n[n, l]] - d[ref_n[n, l]]) < 1e-6, (tag, "nominal", n, l)
            n_nom_diff += 1
        smin = fr["foothold_score"].min(dim=1)[0]                      # [N,4]
        n_fallback += int((smin >= 8).sum())
        n_exact_tie += int(((fr["foothold_score"] == smin[:, None, :]).sum(dim=1) > 1).sum())
        same = (env.optimal_foothold_indice.squeeze(1) == fr["optimal_idx"]).all(dim=1)
        _close(env.foothold_obs[same], fr["foothold_obs"][same], tag + "foothold_obs")
        _close(env.optimal_footholds_world[same], fr["optimal_footholds_world"][same], tag + "opt world")
        assert torch.equal(done.bool(), fr["done"].bool()), tag + "done"
        assert torch.equal(env.time_out_buf, fr["time_outs"]), tag + "time_outs"
        _close(rew, fr["rew"], tag + "rew", atol=2e-6)
        for k, v in fr["episode_sums"].items():
            _close(env.episode_sums[k], v, tag + "episode_sums." + k, atol=2e-6)
        _close(obs[same], fr["obs"][same], tag + "obs")
        _close(priv, fr["priv"], tag + "priv", atol=1e-5)
        assert torch.equal(env.terrain_levels, fr["terrain_levels"]), tag + "terrain_levels"
        _close(env.env_origins, fr["env_origins"], tag + "env_origins")
        assert torch.equal(env.episode_length_buf, fr["episode_length"]), tag + "episode_length"
        _close(fg.root_states, fr["root_after"], tag + "root after reset")
        _close(fg.dof_state, fr["dof_after"], tag + "dof after reset")
        _close(env.motor_strengths[:, 0], fr["motor"], tag + "motor")
        _close(env.height_noise_offset[:, 0], fr["hno"], tag + "height_noise_offset")
        _close(env.feet_air_time, fr["feet_air_time"], tag + "feet_air_time")
        _close(env.pitch_est, fr["pitch_est"], tag + "pitch_est")
        _close(env.get_base_vel(), fr["base_vel"], tag + "base_vel")
        for k, v in fr["extras_episode"].items():
            _close(torch.as_tensor(extras["episode"][k]).float().reshape(()), torch.as_tensor(v).float().reshape(()),
                   tag + "extras." + k, atol=2e-6)
    n_pairs = len(G["frames"]) * G["N"] * 4
    print(f"golden index parity: {n_idx_diff} optimal / {n_nom_diff} nominal near-tie differences of {n_pairs} (env, leg) pairs; "
          f"{n_fallback} pairs on the fall-back branch (min score >= 8), {n_exact_tie} with an exact tie at the minimum")
    assert n_idx_diff == 0 and n_nom_diff == 0, (n_idx_diff, n_nom_diff)
    # the goldens must exercise the reference's topk tie-break: fall-back argmin and exact multi-way ties
    assert n_fallback >= 12 and n_exact_tie >= 8, (n_fallback, n_exact_tie)


def test_known_answers():
    """KATs derived in SURVEY.md section 4."""
    # foothold_obs decode quirk: idx 431 -> x = measured_points_x[431 % 21], y = measured_points_y[(431//21) % 21]
    idx = 431
    assert K.MEASURED_POINTS_X[idx % 21] == pytest.approx(-0.25)
    assert (K.MEASURED_POINTS_Y * 4)[idx // 21] == pytest.approx(0.5)
    # alphabetical reward order, 23 names
    assert len(K.REWARD_NAMES) == 23 and K.REWARD_NAMES[0] == "action_rate" and K.REWARD_NAMES[-1] == "tracking_optimal_footholds"
    # lowest-index argmin on ties == CPU topk(k=1, largest=False)
    # (holds on the reference's shape: 693 >= 64*k selects ATen's partial_sort path, which keeps the first minimum)
    x = torch.full((2, 693, 4), 10.0)
    x[0, 100, 1] = x[0, 400, 1] = x[0, 650, 1] = 8.5
    ti = torch.topk(x, 1, dim=1, largest=False)[1].squeeze(1)
    assert torch.equal(ti, x.argmin(dim=1)) and ti[0, 1].item() == 100 and ti[1, 0].item() == 0
    ac = LO.ActorCriticDecoder(53, 1389, 12)
    assert sum(p.numel() for p in ac.parameters()) == 3193318
    assert sum(p.numel() for p in ac.vae.parameters()) == 2178125


@pytest.fixture(scope="module")
def learner_gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "learner_n8.pt"), weights_only=False)


def _digest(t):
    t = t.detach().double().flatten()
    return torch.tensor([t.sum(), t.abs().sum(), (t * torch.arange(1, t.numel() + 1, dtype=torch.float64)).sum() / t.numel()])


def test_learner_matches_reference(learner_gold):
    G = learner_gold
    N, T = G["N"], G["T"]
    hs, tor = sim_stub.make_heightmap(*G["heightmap"])
    fg = sim_stub.FakeGym(N)
    fg.queue.extend(G["states"])
    rng = Replay(G["init_log"])
    env = EO.OracleEnv(K, N, hs, G["layout"], fg, rng)
    wenv = EO.OracleHistoryWrapper(env)
    torch.manual_seed(G["param_seed"])
    ac = LO.ActorCriticDecoder(53, 1389, 12, rng=rng)
    # P5: same construction order => identical parameters from the same seed
    for k, v in ac.state_dict().items():
        assert torch.allclose(_digest(v), G["param_digest0"][k], rtol=0, atol=0), "init " + k
    assert list(ac.state_dict().keys()) == G["checkpoint_keys"]
    alg = LO.PPO(ac, entropy_coef=0.003, learning_rate=1e-3, rng=rng)
    alg.init_storage(N, T, [53], [1389], [265], [12])
    wenv.reset()
    assert rng.done()
    for it, out in enumerate(G["iters_out"]):
        rng = Replay(out["log"])
        env.rng = ac.rng = ac.vae.rng = alg.rng = alg.storage.rng = rng
        if it == 0:
            rng.randint_like(env.episode_length_buf, 1000)  # on_policy_runner.py:91 (lands on the wrapper)
        obs_dict = wenv.get_observations()
        alg.debug = {}
        obs_dict, losses = LO.learn_iteration(wenv, alg, obs_dict, T)
        assert rng.done()
        st = alg.storage
        tag = f"iter{it} "
        # per-minibatch gradient digests recorded at the reference's clip_grad_norm_ calls (vae, policy alternating)
        ref_d = out["grad_digests"]
        assert len(ref_d) == 2 * len(alg.debug["vae_grads"]) == 40
        errs = []
        for k in range(len(ref_d)):
            gd = alg.debug["vae_grads" if k % 2 == 0 else "ppo_grads"][k // 2]
            d = _digest(torch.cat([g.flatten() for g in gd.values()]))
            errs.append(float((d - ref_d[k]).abs().max() / ref_d[k][1]))
        print(f"{tag}gradient digests vs reference, |diff| / sum|g|: first pair {errs[0]:.1e} {errs[1]:.1e}, median "
              f"{sorted(errs)[20]:.1e}, worst {max(errs):.1e}")
        if it == 0:
            # minibatch 0 starts from bit-identical parameters; later steps inherit the Adam amplification described below, and
            # the latent_var outlier repair is discontinuous at its 2-sigma threshold (one step of the 40 lands on it: 0.12)
            assert max(errs[:6]) <= 1e-6 and sorted(errs)[-3] <= 1e-4 and max(errs) <= 0.5, errs
        else:
            assert sorted(errs)[20] <= 5e-2, errs
        # iteration 0 starts from bit-identical parameters -> tight; iteration 1 starts from post-Adam
        # parameters that legitimately differ by ~1e-5 (see below) -> structural check at 3e-3
        f = 1.0 if it == 0 else 300.0
        _close(st.rewards, out["rewards"], tag + "rewards", atol=2e-6 * f)
        assert torch.equal(st.dones, out["dones"])
        _close(st.actions, out["actions"], tag + "actions", rtol=1e-5 * f, atol=1e-5 * f)
        _close(st.values, out["values"], tag + "values", rtol=1e-5 * f, atol=1e-6 * f)
        _close(st.returns, out["returns"], tag + "returns", rtol=1e-5 * f, atol=1e-5 * f)
        _close(st.advantages, out["advantages"], tag + "advantages", rtol=1e-4 * f, atol=1e-4 * f)
        assert alg.learning_rate == pytest.approx(out["lr"], rel=1e-9)
        _close(ac.std.detach(), out["std"], tag + "std", rtol=1e-5 * f, atol=1e-6 * f)
        # Post-Adam weights: Adam normalises each element's step to ~lr regardless of gradient size, so a
        # 1e-7 relative input difference (the 1-ulp sqrt/sum effects documented in oracle/env_oracle.py) can
        # move an element whose gradient is ~0 by up to lr per step.  Bound: 20 steps x lr_max(1e-3) is the
        # worst case; observed 6e-6.  Aggregate digests are compared at 1e-3 relative.
        _close(ac.actor_body[6].weight.detach(), out["actor_last_w"], tag + "actor W", rtol=1e-3, atol=3e-5 if it == 0 else 1e-3)
        _close(ac.vae.latent_var.weight.detach(), out["latent_var_w"], tag + "latent_var W", rtol=1e-3, atol=3e-5 if it == 0 else 1e-3)
        for k, v in ac.state_dict().items():
            d, r = _digest(v), out["param_digest"][k]
            assert abs(float(d[1] - r[1])) <= (1e-3 if it == 0 else 1e-2) * max(1e-3, float(r[1])), (tag, k, d, r)


TERRAIN_CASES = {  # same table as tests/golden/make_golden.py
    "lite3": (7, {}),
    "mix10": (3, dict(num_rows=3, num_cols=10)),
    "custom3": (5, dict(num_rows=4, num_cols=3, terrain_proportions=[0, 0, 0, 0, 0, 0, 1 / 3, 1 / 3, 1 / 3])),
    "random": (9, dict(num_rows=2, num_cols=4, curriculum=False, terrain_proportions=[0, 0, 0.2, 0.2, 0.2, 0.2, 0.1, 0.05, 0.05])),
}


def terrain_cfg(overrides):
    """cfg.terrain of the Lite3 DTC task (lite3_dtc_config.py:20-51) with a test case's overrides."""
    import types
    base = dict(horizontal_scale=0.05, vertical_scale=0.005, border_size=20, curriculum=True, terrain_length=8.0, terrain_width=8.0,
                num_rows=6, num_cols=2, terrain_proportions=[0.0, 0.0, 0.2, 0.2, 0.2, 0.4], mesh_type="trimesh", selected=False,
                max_init_terrain_level=5, slope_treshold=0.75)
    base.update(overrides)
    return types.SimpleNamespace(**base)


@pytest.mark.parametrize("name", sorted(TERRAIN_CASES))
def test_terrain_matches_reference(name, golden_dir):
    """N3: the terrain oracle against heightmaps / env origins recorded from the unmodified reference `Terrain` class."""
    from oracle import terrain_oracle as TO
    seed, ov = TERRAIN_CASES[name]
    G = np.load(os.path.join(golden_dir, f"terrain_{name}.npz"))
    np.random.seed(seed)
    hf, origins = TO.terrain_map(terrain_cfg(ov))
    assert hf.dtype == np.int16 and hf.shape == G["height_field_raw"].shape
    assert np.array_equal(hf, G["height_field_raw"]), int((hf != G["height_field_raw"]).sum())
    assert np.array_equal(origins, G["env_origins"])
