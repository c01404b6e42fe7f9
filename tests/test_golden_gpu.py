"""GPU parity against the REFERENCE's own outputs (tests/golden/*.pt, recorded from the unmodified reference by
tests/golden/make_golden.py), through the public Python surface and the C ABI underneath:

* the CUDA environment replays the 13 golden frames (incl. the tie / fall-back frames that pin torch.topk's tie-break);
* `HistoryWrapper.obs_history` against the oracle's wrapper over 9 steps with a mid-episode reset (history_wrapper.py:23,40);
* one full `OnPolicyRunner.learn(1)` iteration on the CUDA path replaying the reference's recorded draws: rollout storage,
  returns, advantages, learning rate and the per-minibatch gradient digests against the reference goldens.
"""
import os

import pytest
import torch

import dtc_b200  # noqa: F401
from dtc_b200 import lite3 as K, sim_stub
from oracle import env_oracle as EO, learner_oracle as LO
from oracle.rng import Replay
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(a, b, name, rtol=1e-5, atol=2e-6):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if not torch.allclose(a, b, rtol=rtol, atol=atol):
        d = (a - b).abs()
        i = int(d.argmax())
        raise AssertionError(f"{name}: max abs diff {d.max():.3e} (reference {b.flatten()[i]:.8g}, cuda {a.flatten()[i]:.8g}), "
                             f"{int((d > atol + rtol * b.abs()).sum())} of {d.numel()} elements off")


def _digest(t):
    t = t.detach().double().flatten().cpu()
    return torch.tensor([t.sum(), t.abs().sum(), (t * torch.arange(1, t.numel() + 1, dtype=torch.float64)).sum() / t.numel()])


def test_cuda_env_matches_reference_golden(golden_dir):
    G = torch.load(os.path.join(golden_dir, "env_n16.pt"), weights_only=False)
    oenv, cenv, fg_cpu, fg_gpu = H.make_pair_from_golden(G)
    H.reset_both(oenv, cenv, fg_cpu, fg_gpu, G["states"][0], log=G["reset_log"])
    a = G["after_reset"]
    _close(cenv.obs_buf, a["obs"], "reset obs")
    _close(cenv.commands, a["commands"], "reset commands")
    _close(cenv.rew_buf, a["rew"], "reset rew", atol=1e-5)
    for e in (oenv, cenv):
        e.episode_length_buf[0:4] = 498
        e.episode_length_buf[4:6] = 999
        e.common_step_counter = 747
    n_opt = n_nom = n_fallback = 0
    for t, fr in enumerate(G["frames"]):
        H.lockstep(oenv, cenv, fg_cpu, fg_gpu, G["states"][t + 1], fr["actions"], log=fr["log"])
        tag = f"frame{t} "
        assert torch.equal(cenv.measured_heights.cpu(), fr["measured_heights"]), tag + "measured_heights must be bit-exact"
        ci, ri = cenv.optimal_foothold_indice.squeeze(1).cpu(), fr["optimal_idx"]
        for n, l in (ci != ri).nonzero().tolist():
            sc = fr["foothold_score"][n, :, l]
            assert abs(float(sc[ci[n, l]] - sc[ri[n, l]])) < 1e-6, (tag, "optimal idx", n, l)
            n_opt += 1
        ni, rn = cenv.nominal_footholds_indice.cpu(), fr["nominal_idx"]
        n_nom += int((ni != rn).sum())
        n_fallback += int((fr["foothold_score"].min(dim=1)[0] >= 8).sum())
        if t >= 10:  # the tie / fall-back frames: exact, no tolerance
            assert torch.equal(ci, ri) and torch.equal(ni, rn), tag + "tie-break differs from the reference's topk / argmin"
        same = (ci == ri).all(dim=1)
        sd = same.to(cenv.device)
        _close(cenv.pred_footholds, fr["pred_footholds"], tag + "pred_footholds")
        _close(cenv.base_lin_vel, fr["base_lin_vel"], tag + "base_lin_vel")
        _close(cenv.commands, fr["commands"], tag + "commands")
        _close(cenv.torques, fr["torques"], tag + "torques", atol=1e-5)
        _close(cenv.measured_foot_clearance, fr["clearance"], tag + "clearance")
        _close(cenv.foothold_obs[sd], fr["foothold_obs"][same], tag + "foothold_obs")
        _close(cenv.optimal_footholds_world[sd], fr["optimal_footholds_world"][same], tag + "optimal_footholds_world")
        assert torch.equal(cenv.reset_buf.bool().cpu(), fr["done"].bool()), tag + "done"
        assert torch.equal(cenv.time_out_buf.bool().cpu(), fr["time_outs"]), tag + "time_outs"
        _close(cenv.rew_buf[sd], fr["rew"][same], tag + "rew", atol=5e-6)
        for k, v in fr["episode_sums"].items():
            _close(cenv.episode_sums[k][sd], v[same], tag + "episode_sums." + k, atol=5e-6)
        _close(cenv.obs_buf[sd], fr["obs"][same], tag + "obs")
        _close(cenv.privileged_obs_buf, fr["priv"], tag + "priv", atol=1e-5)
        assert torch.equal(cenv.terrain_levels.cpu(), fr["terrain_levels"]), tag + "terrain_levels"
        _close(cenv.env_origins, fr["env_origins"], tag + "env_origins")
        assert torch.equal(cenv.episode_length_buf.cpu(), fr["episode_length"]), tag + "episode_length"
        _close(cenv.root_states, fr["root_after"], tag + "root after reset")
        _close(cenv.dof_state, fr["dof_after"], tag + "dof after reset")
        _close(cenv.motor_strengths[:, 0], fr["motor"], tag + "motor")
        _close(cenv.height_noise_offset[:, 0], fr["hno"], tag + "height_noise_offset")
        _close(cenv.feet_air_time, fr["feet_air_time"], tag + "feet_air_time")
        _close(cenv.pitch_est, fr["pitch_est"], tag + "pitch_est", atol=5e-6)
        _close(cenv.get_base_vel(), fr["base_vel"], tag + "base_vel")
        if fr["extras_episode"]:
            ep = cenv.extras["episode"]
            for k, v in fr["extras_episode"].items():
                _close(torch.as_tensor(ep[k]).float().reshape(()), torch.as_tensor(v).float().reshape(()), tag + "extras." + k, atol=5e-6)
    print(f"[index parity vs reference goldens] {n_opt} optimal / {n_nom} nominal differences of {len(G['frames']) * G['N'] * 4} "
          f"(env, leg) pairs, {n_fallback} pairs on the fall-back branch")
    assert n_opt == 0 and n_nom == 0, (n_opt, n_nom)
    assert n_fallback >= 12


def test_obs_history_parity():
    """E15: the shift-concat fused into dtc_env_observe against OracleHistoryWrapper (history_wrapper.py:18-49) over nine steps:
    reset() clears it, get_observations() shifts once more, step() shifts, and an in-episode reset does NOT clear it (quirk)."""
    from dtc_b200.rsl_rl.env.wrappers import HistoryWrapper
    N = 64
    oenv, cenv, fg_cpu, fg_gpu = H.make_pair(N, "stones", seed=4)
    ow, cw = EO.OracleHistoryWrapper(oenv), HistoryWrapper(cenv)
    g = torch.Generator().manual_seed(3)
    states = [sim_stub.synth_state(N, oenv.env_origins, g) for _ in range(10)]
    states[4]["root_states"][1, 3:7] = torch.tensor([0.9, 0.0, 0.0, 0.435])  # flipped robot -> in-episode reset at step 3
    states[6]["root_states"][7, 2] -= 0.4                                    # sunk robot -> in-episode reset at step 5
    od, cd = H.reset_both(oenv, cenv, fg_cpu, fg_gpu, states[0], oreset=ow.reset, creset=cw.reset)
    assert float(cd["obs_history"].abs().max()) == 0.0 and tuple(cd["obs_history"].shape) == (N, 265)
    od, cd = ow.get_observations(), cw.get_observations()
    _close(cd["obs_history"], od["obs_history"], "history after get_observations")
    ag = torch.Generator().manual_seed(5)
    n_resets = 0
    for t in range(9):
        (od, _, odone, _), (cd, _, cdone, _) = H.lockstep(oenv, cenv, fg_cpu, fg_gpu, states[t + 1], torch.randn(N, 12, generator=ag),
                                                         ostep=ow.step, cstep=cw.step)
        n_resets += int(odone.sum())
        assert torch.equal(cdone.bool().cpu(), odone.bool())
        for k in ("obs", "privileged_obs", "obs_history", "base_vel"):
            _close(cd[k], od[k], f"step{t} {k}", atol=1e-5 if k == "privileged_obs" else 2e-6)
        # the newest frame of the history IS the observation; the oldest is the one from five steps ago
        assert torch.equal(cd["obs_history"][:, 212:], cd["obs"])
    assert n_resets >= 2, "the in-episode reset path must be exercised"
    assert bool((cd["obs_history"][torch.tensor([1, 7], device=DEV)][:, :212].abs().sum(dim=1) > 0).all())  # NOT cleared by the in-episode reset


def _replay_learner_golden(G, it_count=1):
    """Runs the oracle over the recorded draws of learner_n8.pt and cuts the log into the injection tables of the CUDA path."""
    N, T = G["N"], G["T"]
    hs, tor = sim_stub.make_heightmap(*G["heightmap"])
    fg = sim_stub.FakeGym(N)
    fg.queue.extend(G["states"])
    rng = Replay(G["init_log"])
    env = EO.OracleEnv(K, N, hs, G["layout"], fg, rng)
    wenv = EO.OracleHistoryWrapper(env)
    torch.manual_seed(G["param_seed"])
    ac = LO.ActorCriticDecoder(53, 1389, 12, rng=rng)
    alg = LO.PPO(ac, entropy_coef=0.003, learning_rate=1e-3, rng=rng)
    alg.init_storage(N, T, [53], [1389], [265], [12])
    wenv.reset()
    u, normal, rest = H.reset_tables(G["init_log"], N, DEV)
    none = torch.zeros(0, dtype=torch.long)
    hd, nz = H.step_tables(rest, N, none, env.reset_buf.nonzero().flatten(), 1, DEV)
    hd.update(reset0_u=u, reset0_normal=normal)
    plan = dict(reset=(hd, nz), iters=[])
    for it in range(it_count):
        out = G["iters_out"][it]
        log = out["log"]
        rng = Replay(log)
        env.rng = ac.rng = ac.vae.rng = alg.rng = alg.storage.rng = rng
        if it == 0:
            rng.randint_like(env.episode_length_buf, 1000)  # on_policy_runner.py:91 (lands on the wrapper)
        od = wenv.get_observations()
        acts, steps = [], []
        obs, priv, hist = od["obs"], od["privileged_obs"], od["obs_history"]
        with torch.inference_mode():
            for _ in range(T):
                p0 = rng.pos
                actions = alg.act(obs, priv, hist, od["base_vel"], wenv.get_reward_buf())
                assert rng.pos == p0 + 2
                acts.append(dict(eps_z=log[p0][1].to(DEV).contiguous(), eps_a=log[p0 + 1][1].to(DEV).contiguous()))
                resample_ids = ((env.episode_length_buf + 1) % K.RESAMPLING_STEPS == 0).nonzero().flatten()
                counter = env.common_step_counter + 1
                p0 = rng.pos
                od, rewards, dones, infos = wenv.step(actions)
                steps.append(H.step_tables(log[p0:rng.pos], N, resample_ids, env.reset_buf.nonzero().flatten(), counter, DEV))
                obs, priv, hist = od["obs"], od["privileged_obs"], od["obs_history"]
                alg.process_env_step(rewards, dones, next_obs=od["obs"], infos=infos)
            alg.compute_returns(obs, priv, od["base_vel"])
        p0 = rng.pos
        alg.debug = {}
        alg.update()
        assert rng.done()
        ulog = log[p0:]
        assert ulog[0][0] == "randperm" and len(ulog) == 1 + 3 * 20
        eps = []
        for k in range(20):
            eps += [ulog[1 + 3 * k][1].to(DEV).contiguous(), ulog[2 + 3 * k][1].to(DEV).contiguous()]
        keys = [list(alg.debug["vae_grads"][0].keys()), list(alg.debug["ppo_grads"][0].keys())]
        plan["iters"].append(dict(acts=acts, steps=steps, perm=ulog[0][1], eps=eps, grad_keys=keys))
    return plan


def test_cuda_runner_matches_reference_golden(golden_dir):
    """P13 + SURVEY section 4: one OnPolicyRunner.learn() iteration on the CUDA path (rollout through HistoryWrapper / PPO.act /
    process_env_step, compute_returns, update) with the draws the reference consumed, against what the REFERENCE runner left in
    its storage and optimizer (tests/golden/learner_n8.pt)."""
    from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg, Lite3DTCCfgPPO
    from dtc_b200.legged_gym.envs.lite3.lite3_dtc_config import class_to_dict
    from dtc_b200.rsl_rl.runners import OnPolicyRunner
    from dtc_b200.rsl_rl.modules.actor_critic_decoder import STATE_KEYS
    G = torch.load(os.path.join(golden_dir, "learner_n8.pt"), weights_only=False)
    N, T = G["N"], G["T"]
    plan = _replay_learner_golden(G, 1)
    hs, tor = sim_stub.make_heightmap(*G["heightmap"])
    fg = sim_stub.FakeGym(N, device=DEV)
    fg.queue.extend({k: v.to(DEV) for k, v in s.items()} for s in G["states"])
    cfg = Lite3DTCCfg()
    cfg.env.num_envs = N
    cenv = LeggedRobotDTC(cfg, sim_device=DEV, gym=fg, height_samples=hs, terrain_origins=tor, layout=G["layout"], seed=G["seed"])
    cenv._host_draws, cenv._noise = plan["reset"]
    torch.manual_seed(G["param_seed"])
    runner = OnPolicyRunner(cenv, class_to_dict(Lite3DTCCfgPPO()), log_dir=None, device=DEV)
    alg, ac = runner.alg, runner.alg.actor_critic
    for k, v in ac.state_dict().items():  # P5: bit-identical initial parameters from the same torch seed
        # (float64 digests: the summation order of torch.sum differs between host CPUs, hence 1e-12 rather than bit equality)
        assert torch.allclose(_digest(v), G["param_digest0"][k], rtol=1e-12, atol=1e-12), "init " + k
    P = plan["iters"][0]
    acts, steps = list(P["acts"]), list(P["steps"])
    orig_act, orig_step = alg.act, cenv.step

    def act(*a, **kw):
        ac._inject = acts.pop(0)
        return orig_act(*a, **kw)

    def step(actions):
        cenv._host_draws, cenv._noise = steps.pop(0)
        return orig_step(actions)

    alg.act, cenv.step = act, step
    alg._inject = dict(perm=P["perm"], eps=P["eps"])
    digests = []

    def tap(which):
        keys = P["grad_keys"][which]
        pre = "vae." if which == 0 else ""
        digests.append(_digest(torch.cat([ac._grads[ac._idx[pre + k]] for k in keys])))

    alg._grad_tap = tap
    runner.learn(1, init_at_random_ep_len=True)
    assert not acts and not steps
    out, st = G["iters_out"][0], alg.storage
    tag = "iter0 "
    _close(st.rewards, out["rewards"], tag + "rewards", atol=5e-6)
    assert torch.equal(st.dones.cpu(), out["dones"])
    _close(st.actions, out["actions"], tag + "actions", rtol=1e-5, atol=1e-5)
    _close(st.mu, out["mu"], tag + "mu", rtol=1e-5, atol=1e-5)
    _close(st.values, out["values"], tag + "values", rtol=1e-5, atol=1e-6)
    _close(st.actions_log_prob, out["logp"], tag + "logp", rtol=1e-5, atol=2e-5)
    _close(st.returns, out["returns"], tag + "returns", rtol=1e-5, atol=1e-5)
    _close(st.advantages, out["advantages"], tag + "advantages", rtol=1e-4, atol=1e-4)
    assert alg.learning_rate == pytest.approx(out["lr"], rel=1e-9), "the adaptive-KL schedule must take the reference's branches"
    _close(ac.std, out["std"], tag + "std", rtol=1e-5, atol=1e-6)
    # gradient digests at the reference's 40 clip_grad_norm_ calls
    ref_d = out["grad_digests"]
    assert len(digests) == len(ref_d) == 40
    errs = [float((d - r).abs().max() / r[1]) for d, r in zip(digests, ref_d)]
    print(f"[runner vs reference goldens] gradient digests |diff| / sum|g|: first pair {errs[0]:.1e} {errs[1]:.1e}, median "
          f"{sorted(errs)[20]:.1e}, worst {max(errs):.1e}")
    # same bounds as the CPU oracle gets against the reference (tests/test_oracle_golden.py): minibatch 0 starts from identical
    # parameters, later ones inherit Adam's amplification of round-off, one step sits on the outlier-repair discontinuity
    assert max(errs[:2]) <= 2e-5 and sorted(errs)[20] <= 2e-4 and sorted(errs)[-3] <= 5e-3 and max(errs) <= 0.5, errs
    # post-update parameters: same structural bounds as the oracle-vs-reference test
    # (the oracle lands within 3e-5 of the reference here, the CUDA path within 7.3e-5: 20 Adam steps of lr <= 1e-3 amplify the
    # 1e-7 gradient round-off of elements whose gradient is ~0, see tests/test_oracle_golden.py)
    _close(ac.state_dict()["actor_body.6.weight"], out["actor_last_w"], tag + "actor W", rtol=1e-3, atol=1.5e-4)
    _close(ac.state_dict()["vae.latent_var.weight"], out["latent_var_w"], tag + "latent_var W", rtol=1e-3, atol=1.5e-4)
    for k, v in ac.state_dict().items():
        d, r = _digest(v), out["param_digest"][k]
        assert abs(float(d[1] - r[1])) <= 1e-3 * max(1e-3, float(r[1])), (tag, k, d, r)
    assert list(ac.state_dict().keys()) == G["checkpoint_keys"] == list(STATE_KEYS)


def test_save_load_resume_roundtrip(tmp_path, golden_dir):
    """N1: OnPolicyRunner.save -> get_load_path -> make_alg_runner(resume) -> learn (on_policy_runner.py:249-264, helpers.py:73-95,
    task_registry.py:123-128).  The file has the structure of the reference runner's own model_<it>.pt (recorded in the goldens),
    a fresh runner resumed from it holds identical parameters / Adam moments / step counts / learning rate / iteration, and
    continues training."""
    import time
    from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg, Lite3DTCCfgPPO
    from dtc_b200.legged_gym.utils import class_to_dict, get_load_path, make_alg_runner
    G = torch.load(os.path.join(golden_dir, "learner_n8.pt"), weights_only=False)
    N = 64

    def build(seed):
        hs, tor = sim_stub.make_heightmap("stones", 0)
        layout = sim_stub.initial_env_layout(N, tor, seed)
        fg = sim_stub.FakeGym(N, device=DEV)
        g = torch.Generator(device=DEV).manual_seed(seed)
        fg.source = lambda: sim_stub.synth_state(N, layout[2].to(DEV), g, device=DEV)
        cfg = Lite3DTCCfg()
        cfg.env.num_envs = N
        return LeggedRobotDTC(cfg, sim_device=DEV, gym=fg, height_samples=hs, terrain_origins=tor, layout=layout, seed=seed)

    root = str(tmp_path / "logs")
    with pytest.raises(ValueError):
        get_load_path(root)
    torch.manual_seed(3)
    ra, _ = make_alg_runner(build(1), Lite3DTCCfgPPO(), log_root=root, device=DEV)
    ra.learn(2)
    path = get_load_path(root)
    assert os.path.basename(path) == "model_2.pt" and get_load_path(root, checkpoint=0).endswith("model_0.pt")
    ck = torch.load(path, map_location="cpu", weights_only=True)
    ref = G["checkpoint_struct"]
    assert list(ck.keys()) == ref["top_keys"] and ck["iter"] == 2
    assert {k: tuple(v.shape) for k, v in ck["model_state_dict"].items()} == ref["model"]
    og, rg = ck["optimizer_state_dict"]["param_groups"], ref["opt_groups"]
    assert len(og) == len(rg) == 1 and len(og[0]["params"]) == rg[0]["params"] and set(rg[0]) <= set(og[0])
    ost = ck["optimizer_state_dict"]["state"]
    assert set(ost.keys()) == set(ref["opt_state"].keys())
    for i, st in ref["opt_state"].items():
        assert {k: tuple(ost[i][k].shape) for k in st} == {k: v[0] for k, v in st.items()}, i
    time.sleep(1.1)  # run directories are named to the second
    class Resume(Lite3DTCCfgPPO):
        class runner(Lite3DTCCfgPPO.runner):
            resume = True
    torch.manual_seed(99)  # a different initialisation, overwritten by the checkpoint
    rb, _ = make_alg_runner(build(2), Resume(), log_root=root, device=DEV)
    assert rb.current_learning_iteration == 2
    sa, sb = ra.alg.actor_critic.state_dict(), rb.alg.actor_critic.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    oa, ob = ra.alg.optimizer.state_dict(), rb.alg.optimizer.state_dict()
    assert oa["param_groups"][0]["lr"] == ob["param_groups"][0]["lr"] and set(oa["state"]) == set(ob["state"])
    for i in oa["state"]:
        for key in ("step", "exp_avg", "exp_avg_sq"):
            assert torch.equal(oa["state"][i][key].cpu(), ob["state"][i][key].cpu()), (i, key)
    rb.learn(1)
    assert rb.current_learning_iteration == 3 and os.path.exists(os.path.join(rb.log_dir, "model_3.pt"))
    assert all(v == v for v in rb.alg.last_stats.values())
    assert float(rb.alg.optimizer.state_dict()["state"][0]["step"]) == float(oa["state"][0]["step"]) + 20


def test_graph_rollout_equals_eager_rollout():
    """The rollout captured into one CUDA graph (OnPolicyRunner: T x {PPO.act, env.step, process_env_step} with device-side step /
    Philox counters) leaves bit-identical storage and environment state to the same iterations launched eagerly -- with the
    simulator tensors coming from device memory, and from pinned host memory through the staging ring (bench.py's e2e leg)."""
    from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg, Lite3DTCCfgPPO
    from dtc_b200.legged_gym.envs.lite3.lite3_dtc_config import class_to_dict
    from dtc_b200.rsl_rl.runners import OnPolicyRunner
    N = 256
    hs, tor = sim_stub.make_heightmap("stones", 0)
    layout = sim_stub.initial_env_layout(N, tor, 5)
    g = torch.Generator().manual_seed(5)
    host = [{k: v.pin_memory() for k, v in sim_stub.synth_state(N, layout[2], g).items()} for _ in range(8)]
    # a flipped robot: resets and resampling happen inside the captured steps as well
    host[3]["root_states"][1, 3:7] = torch.tensor([0.9, 0.0, 0.0, 0.435])
    dev = [{k: v.to(DEV) for k, v in s.items()} for s in host]

    def build(use_graph, pool, prefetch):
        fg = sim_stub.FakeGym(N, device=DEV)
        st = {"i": 0}

        def source():
            st["i"] = (st["i"] + 1) % 8  # 24 steps per iteration: every iteration sees the same sequence in every mode
            return pool[st["i"]]

        fg.source = source
        cfg = Lite3DTCCfg()
        cfg.env.num_envs = N
        env = LeggedRobotDTC(cfg, sim_device=DEV, gym=fg, height_samples=hs, terrain_origins=tor, layout=layout, seed=5)
        torch.manual_seed(7)
        tc = class_to_dict(Lite3DTCCfgPPO())
        tc["runner"]["cuda_graph"] = use_graph
        r = OnPolicyRunner(env, tc, log_dir=None, device=DEV)
        if prefetch:
            fg.enable_prefetch(True)
        r.env.env.episode_length_buf[5:9] = 980  # episodes about to time out
        return r

    runners = [build(False, dev, False), build(True, dev, False), build(True, host, True)]
    ra = runners[0]
    for it in range(4):  # iteration 0 eager everywhere, then the graph runners capture (1) and replay (2, 3)
        for r in runners:
            torch.manual_seed(100 + it)  # the minibatch permutation comes from torch's generator
            r.learn(1)
        for r in runners[1:]:
            what = f"iteration {it}, {'host ring' if r is runners[2] else 'device pool'}"
            assert (r._graph is not None) == (it >= 1)
            ea, eb = ra.env.env, r.env.env
            assert ea.common_step_counter == eb.common_step_counter and ra.alg.actor_critic._calls == r.alg.actor_critic._calls
            for name in ("actions", "values", "rewards", "returns", "advantages", "hist", "priv_a", "xc", "dones"):
                assert torch.equal(getattr(ra.alg.storage, name), getattr(r.alg.storage, name)), (what, name)
            for name in ("obs_buf", "rew_buf", "episode_length_buf", "commands", "terrain_levels", "_episode_sums", "measured_heights"):
                assert torch.equal(getattr(ea, name), getattr(eb, name)), (what, name)
            # The UPDATE is not run-to-run deterministic (float atomics in the weight-gradient reductions; Adam then normalises that
            # noise to +-lr on parameters whose gradient is zero, and the KL-adaptive learning rate can branch on it), with or
            # without the graph -- so the parameters are re-synchronised and the claim checked here is the rollout's.
            r.alg.actor_critic._flat.copy_(ra.alg.actor_critic._flat)
            r.alg.actor_critic._params_written()
    assert runners[1]._graph_launches > 24 * 20
    assert int(ra.env.env.episode_length_buf.min()) < 24 * 4, "resets must have happened inside the captured rollouts"
