"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python tests/golden/make_golden.py
The reference is imported through oracle/ref_harness (import stubs + FakeGym); every random draw it
makes is logged by oracle.rng.Recorder so the oracle restatement can replay it.
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dtc_b200  # noqa: E402
from dtc_b200 import sim_stub  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402
from oracle.rng import Recorder  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def digest(t):
    t = t.detach().double().flatten()
    return torch.tensor([t.sum(), t.abs().sum(), (t * torch.arange(1, t.numel() + 1, dtype=torch.float64)).sum() / t.numel()])


def _place(state, n, xy=None, dz=0.0, lin_vel=None):
    """Moves robot n of a synthetic state rigidly (root and every body) and/or overrides its world linear velocity."""
    rs, rb = state["root_states"], state["rigid_body_state"].view(-1, 17, 13)
    if xy is not None:
        d = torch.tensor(xy) - rs[n, 0:2]
        rs[n, 0:2] += d
        rb[n, :, 0:2] += d
    if dz:
        rs[n, 2] += dz
        rb[n, :, 2] += dz
    if lin_vel is not None:
        rs[n, 7:10] = torch.tensor(lin_vel)


def env_golden(N=16, steps=13, seed=11):
    torch.manual_seed(seed)
    np.random.seed(seed)
    hs, tor = sim_stub.make_heightmap("stones", 0)
    layout = sim_stub.initial_env_layout(N, tor, seed)
    fg = sim_stub.FakeGym(N)
    g = torch.Generator().manual_seed(seed + 1)
    states = [sim_stub.synth_state(N, layout[2], g) for _ in range(steps + 1)]
    # make sure the nasty paths are hit: tilt one robot over, drop one into a pit
    states[3]["root_states"][1, 3:7] = torch.tensor([0.9, 0.0, 0.0, 0.435])
    states[5]["root_states"][2, 2] -= 0.4
    # frames 10..12 pin the tie / fall-back branch of topk(k=1, largest=False) by reference output:
    states[11]["root_states"][5:7, 2] += 1.5                    # root only: every grid point is an exception point -> score == 10 everywhere
    _place(states[11], 9, dz=1.5)                               # same with the whole robot lifted
    _place(states[12], 7, lin_vel=[40.0, 0.0, 0.0])             # nominal footholds 1.6 m away: no in-radius candidate on stones
    _place(states[12], 8, xy=[-10.0, -10.0])                    # flat border: identical terrain score at all 693 points
    states[12]["root_states"][8, 2] = 0.35
    _place(states[13], 10, xy=[-8.0, 12.5], lin_vel=[-40.0, 3.0, 0.0])  # flat border AND no in-radius candidate: 693-way exact tie
    states[13]["root_states"][10, 2] = 0.33
    _place(states[13], 11, xy=[4.0, -11.0])
    states[13]["root_states"][11, 2] = 0.31
    env = RH.build_ref_env(N, hs, tor, layout, fg)
    rec = Recorder()
    out = dict(N=N, steps=steps, seed=seed, states=states, heightmap=("stones", 0), layout=layout, frames=[])
    with rec:
        fg.queue.append(states[0])
        env.reset()
        out["reset_log"] = rec.take()
        out["after_reset"] = dict(obs=env.obs_buf.clone(), priv_digest=digest(env.privileged_obs_buf),
                                  commands=env.commands.clone(), rew=env.rew_buf.clone())
        # pokes applied identically to the oracle (exercise resample @500, timeout @1000, push @750)
        env.episode_length_buf[0:4] = 498
        env.episode_length_buf[4:6] = 999
        env.common_step_counter = 747
        ag = torch.Generator().manual_seed(seed + 2)
        for t in range(steps):
            actions = torch.randn(N, 12, generator=ag) * (150.0 if t == 2 else 1.0)
            fg.queue.append(states[t + 1])
            obs, priv, rew, done, extras = env.step(actions)
            fr = dict(actions=actions, log=rec.take(), obs=obs.clone(), priv=priv.clone(), rew=rew.clone(),
                      done=done.clone(), time_outs=env.time_out_buf.clone(),
                      measured_heights=env.measured_heights.clone(), pred_footholds=env.pred_footholds.clone(),
                      optimal_idx=env.optimal_foothold_indice.squeeze(1).clone(),
                      nominal_idx=env.nominal_footholds_indice.clone(), foothold_obs=env.foothold_obs.clone(),
                      optimal_footholds_world=env.optimal_footholds_world.clone(),
                      foothold_score=env.foothold_score.clone(), slope=env.slope.clone(),
                      commands=env.commands.clone(), torques=env.torques.clone(),
                      base_lin_vel=env.base_lin_vel.clone(), clearance=env.measured_foot_clearance.clone(),
                      episode_sums={k: v.clone() for k, v in env.episode_sums.items()},
                      terrain_levels=env.terrain_levels.clone(), env_origins=env.env_origins.clone(),
                      episode_length=env.episode_length_buf.clone(), root_after=env.root_states.clone(),
                      dof_after=env.dof_state.clone(), motor=env.motor_strengths[:, 0].clone(),
                      hno=env.height_noise_offset[:, 0].clone(), feet_air_time=env.feet_air_time.clone(),
                      pitch_est=env.pitch_est.clone(), base_vel=env.get_base_vel().clone(),
                      extras_episode={k: (v.clone() if torch.is_tensor(v) else v) for k, v in extras.get("episode", {}).items()})
            out["frames"].append(fr)
    torch.save(out, os.path.join(OUT, "env_n16.pt"))
    print("env golden: resets per step", [int(f["done"].sum()) for f in out["frames"]])


def learner_golden(N=8, iters=2, seed=5):
    RH.import_reference()
    from rsl_rl.runners import OnPolicyRunner
    from legged_gym.envs.lite3.lite3_dtc_config import Lite3DTCCfgPPO
    from legged_gym.utils.helpers import class_to_dict
    torch.manual_seed(seed)
    np.random.seed(seed)
    hs, tor = sim_stub.make_heightmap("stones", 0)
    layout = sim_stub.initial_env_layout(N, tor, seed)
    fg = sim_stub.FakeGym(N)
    T = 24
    g = torch.Generator().manual_seed(seed + 1)
    states = [sim_stub.synth_state(N, layout[2], g) for _ in range(1 + iters * T)]
    env = RH.build_ref_env(N, hs, tor, layout, fg)
    fg.queue.extend(states)
    train_cfg = class_to_dict(Lite3DTCCfgPPO())
    rec = Recorder()
    out = dict(N=N, T=T, iters=iters, seed=seed, states=states, layout=layout, heightmap=("stones", 0))
    with tempfile.TemporaryDirectory() as d, rec:
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(seed + 7)
            runner = OnPolicyRunner(env, train_cfg, log_dir=d, device="cpu")
        out["init_log"] = rec.take()
        ac = runner.alg.actor_critic
        out["param_seed"] = seed + 7
        out["param_digest0"] = {k: digest(v) for k, v in ac.state_dict().items()}
        out["num_params"] = sum(p.numel() for p in ac.parameters())
        # capture per-minibatch gradients of the first iteration through a hook on clip_grad_norm_
        grads = []
        orig_clip = torch.nn.utils.clip_grad_norm_

        def clip(params, max_norm, *a, **k):
            params = list(params)
            grads.append(digest(torch.cat([p.grad.flatten() for p in params if p.grad is not None])))
            return orig_clip(params, max_norm, *a, **k)

        import rsl_rl.algorithms.ppo as ppo_mod
        ppo_mod.nn.utils.clip_grad_norm_ = clip
        out["iters_out"] = []
        for it in range(iters):
            with contextlib.redirect_stdout(io.StringIO()):
                # learn(1) == one iteration; the reference saves a checkpoint into the temp dir
                st_before = None
                runner.learn(1, init_at_random_ep_len=(it == 0))
            st = runner.alg.storage
            out["iters_out"].append(dict(
                log=rec.take(), grad_digests=list(grads), lr=runner.alg.learning_rate,
                param_digest={k: digest(v) for k, v in ac.state_dict().items()},
                std=ac.std.detach().clone(),
                advantages=st.advantages.clone(), returns=st.returns.clone(), values=st.values.clone(),
                rewards=st.rewards.clone(), dones=st.dones.clone(), actions=st.actions.clone(),
                mu=st.mu.clone(), logp=st.actions_log_prob.clone(),
                actor_last_w=ac.actor_body[6].weight.detach().clone(),
                latent_var_w=ac.vae.latent_var.weight.detach().clone()))
            grads.clear()
        ppo_mod.nn.utils.clip_grad_norm_ = orig_clip
        ck = torch.load(os.path.join(d, "model_%d.pt" % iters), weights_only=True)
        out["checkpoint_keys"] = list(ck["model_state_dict"].keys())
        # structure of the reference's model_<it>.pt (on_policy_runner.py:249-255); the tensors themselves (38 MB) are not committed
        osd = ck["optimizer_state_dict"]
        out["checkpoint_struct"] = dict(
            top_keys=list(ck.keys()), iter=ck["iter"], model={k: tuple(v.shape) for k, v in ck["model_state_dict"].items()},
            opt_groups=[{k: (len(v) if k == "params" else v) for k, v in g.items()} for g in osd["param_groups"]],
            opt_state={i: {k: (tuple(v.shape), str(v.dtype)) for k, v in st.items()} for i, st in osd["state"].items()})
    torch.save(out, os.path.join(OUT, "learner_n8.pt"))
    print("learner golden: lr", [o["lr"] for o in out["iters_out"]], "params", out["num_params"])


TERRAIN_CASES = {
    # name: (seed, overrides of Lite3DTCCfg.terrain)
    "lite3": (7, {}),                                                       # the task's own 6 x 2 curriculum: stairs down | discrete obstacles
    "mix10": (3, dict(num_rows=3, num_cols=10)),                            # + stairs up and Isaac Gym stepping stones
    "custom3": (5, dict(num_rows=4, num_cols=3, terrain_proportions=[0, 0, 0, 0, 0, 0, 1 / 3, 1 / 3, 1 / 3])),  # gap | pit | stones_everywhere
    "random": (9, dict(num_rows=2, num_cols=4, curriculum=False, terrain_proportions=[0, 0, 0.2, 0.2, 0.2, 0.2, 0.1, 0.05, 0.05])),
}


def terrain_cfg(overrides):
    RH.import_reference()
    from legged_gym.envs.lite3.lite3_dtc_config import Lite3DTCCfg
    return type("terrain", (Lite3DTCCfg.terrain,), dict(overrides))


def terrain_golden():
    """N3: the UNMODIFIED reference `Terrain` (legged_gym/utils/terrain.py) over the isaacgym.terrain_utils stand-in."""
    RH.import_reference()
    from legged_gym.utils.terrain import Terrain
    for name, (seed, ov) in TERRAIN_CASES.items():
        np.random.seed(seed)
        t = Terrain(terrain_cfg(ov), 16)
        np.savez_compressed(os.path.join(OUT, f"terrain_{name}.npz"), height_field_raw=t.height_field_raw, env_origins=t.env_origins, seed=seed)
        print("terrain golden", name, t.height_field_raw.shape, int(t.height_field_raw.min()), int(t.height_field_raw.max()))


if __name__ == "__main__":
    terrain_golden()
    if "--terrain-only" in sys.argv:
        sys.exit(0)
    env_golden()
    learner_golden()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
