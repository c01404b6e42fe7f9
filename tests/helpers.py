"""Shared test helpers: run the CPU oracle with recorded draws and convert them to the per-env injection tables
of the CUDA path (include/dtc_b200.h: dtc_env_noise)."""
import numpy as np
import torch

from dtc_b200 import lite3 as K, sim_stub
from oracle import env_oracle as EO
from oracle.rng import Live


class TapRng(Live):
    """Live draws, logged in order as (tag, value)."""

    def __init__(self, seed=0):
        super().__init__(seed)
        self.log = []

    def _t(self, tag, v):
        self.log.append((tag, v.clone() if torch.is_tensor(v) else v))
        return v

    def rand(self, *shape): return self._t("rand", super().rand(*shape))
    def rand_like(self, t): return self._t("rand_like", super().rand_like(t))
    def randn_like(self, t): return self._t("randn_like", super().randn_like(t))
    def randint_like(self, t, high): return self._t("randint_like", super().randint_like(t, high))
    def randperm(self, n): return self._t("randperm", super().randperm(n))
    def np_randint(self, lo, hi): return self._t("np_randint", super().np_randint(lo, hi))
    def np_normal(self, mu, sigma): return self._t("np_normal", super().np_normal(mu, sigma))

    def take(self):
        out, self.log = self.log, []
        return out


def make_pair(N, kind="stones", seed=3, device="cuda", K=K, cfg=None):
    """Oracle env (CPU) and CUDA env fed by twin FakeGyms holding identical synthetic states.
    K / cfg: the task constants on the oracle side and the configuration object on the CUDA side (defaults: Lite3 DTC)."""
    from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg
    hs, tor = sim_stub.make_heightmap(kind, 0)
    layout = sim_stub.initial_env_layout(N, tor, seed)
    rng = TapRng(seed)
    fg_cpu = sim_stub.FakeGym(N)
    oenv = EO.OracleEnv(K, N, hs, layout, fg_cpu, rng)
    fg_gpu = sim_stub.FakeGym(N, device=device)
    cfg = cfg if cfg is not None else Lite3DTCCfg()
    cfg.env.num_envs = N
    cenv = LeggedRobotDTC(cfg, sim_device=device, gym=fg_gpu, height_samples=hs, terrain_origins=tor, layout=layout, seed=seed)
    return oenv, cenv, fg_cpu, fg_gpu


def reset_tables(log, N, device):
    """Draws of a full-batch reset_idx -> reset0_u [N,25], reset0_normal."""
    it = iter(log)
    nxt = lambda tag: _expect(next(it), tag)
    u = torch.zeros(N, 25)
    u[:, 0] = (nxt("randint_like").float() + 0.5) / K.NUM_ROWS
    u[:, 1:13] = nxt("rand")
    u[:, 13:15] = nxt("rand")
    u[:, 15:21] = nxt("rand")
    for k in (21, 22, 23):
        u[:, k] = nxt("rand")[:, 0]
    u[:, 24] = nxt("rand")
    normal = nxt("np_normal")
    return u.to(device), normal, list(it)


def _expect(rec, tag):
    assert rec[0] == tag, (rec[0], tag)
    return rec[1]


def step_tables(log, N, resample_ids, reset_ids, counter, device):
    """Draws of one oracle step() -> (host_draws, noise dict of CUDA tensors)."""
    it = iter(log)
    nxt = lambda tag: _expect(next(it), tag)
    lag = [nxt("np_randint") for _ in range(4)]
    res = torch.zeros(N, 3)
    for k in range(3):
        v = nxt("rand")
        res[resample_ids, k] = v[:, 0]
    push = torch.zeros(N, 2)
    m = counter % K.PUSH_INTERVAL
    if m == 0:
        nxt("rand"); nxt("rand")
    if m in (0, 1):
        push = nxt("rand").clone()
    ru = torch.zeros(N, 25)
    normal = 0.0
    n = len(reset_ids)
    if n:
        ru[reset_ids, 0] = (nxt("randint_like").float() + 0.5) / K.NUM_ROWS
        ru[reset_ids, 1:13] = nxt("rand")
        ru[reset_ids, 13:15] = nxt("rand")
        ru[reset_ids, 15:21] = nxt("rand")
        for k in (21, 22, 23):
            ru[reset_ids, k] = nxt("rand")[:, 0]
        ru[reset_ids, 24] = nxt("rand")
        normal = nxt("np_normal")
    priv_u = nxt("rand_like")
    obs_u = nxt("rand_like")
    rest = list(it)
    assert not rest, f"{len(rest)} unconsumed draws"
    noise = dict(resample_u=res, push_u=push, reset_u=ru, priv_u=priv_u, obs_u=obs_u)
    return dict(lag=lag, reset_normal=normal), {k: v.contiguous().to(device) for k, v in noise.items()}


def lockstep(oenv, cenv, fg_cpu, fg_gpu, state, actions, log=None, ostep=None, cstep=None):
    """One step of both envs on the same state/actions/draws.  Returns (oracle outputs, CUDA outputs).
    log: recorded draws of the reference for this step (golden frames) - the oracle replays them instead of drawing live.
    ostep / cstep: step callables (e.g. the HistoryWrappers' step); default the environments' own."""
    from oracle.rng import Replay
    N = oenv.num_envs
    dev = cenv.device
    resample_ids = ((oenv.episode_length_buf + 1) % K.RESAMPLING_STEPS == 0).nonzero().flatten()
    counter = oenv.common_step_counter + 1
    if log is not None:
        oenv.rng = Replay(log)
    fg_cpu.queue.append(state)
    out = (ostep or oenv.step)(actions)
    reset_ids = oenv.reset_buf.nonzero().flatten()
    if log is not None:
        assert oenv.rng.done(), f"oracle consumed {oenv.rng.pos} of {len(log)} recorded draws"
    hd, nz = step_tables(log if log is not None else oenv.rng.take(), N, resample_ids, reset_ids, counter, dev)
    cenv._host_draws, cenv._noise = hd, nz
    fg_gpu.queue.append({k: v.to(dev) for k, v in state.items()})
    cout = (cstep or cenv.step)(actions.to(dev))
    return out, cout


def reset_both(oenv, cenv, fg_cpu, fg_gpu, state, log=None, oreset=None, creset=None):
    from oracle.rng import Replay
    N, dev = oenv.num_envs, cenv.device
    if log is not None:
        oenv.rng = Replay(log)
    fg_cpu.queue.append(state)
    oout = (oreset or oenv.reset)()
    log = log if log is not None else oenv.rng.take()
    u, normal, rest = reset_tables(log, N, dev)
    none = torch.zeros(0, dtype=torch.long)
    hd, nz = step_tables(rest, N, none, oenv.reset_buf.nonzero().flatten(), 1, dev)
    hd.update(reset0_u=u, reset0_normal=normal)
    cenv._host_draws, cenv._noise = hd, nz
    fg_gpu.queue.append({k: v.to(dev) for k, v in state.items()})
    cout = (creset or cenv.reset)()
    return oout, cout


def make_pair_from_golden(G, device="cuda"):
    """Oracle env + CUDA env on the heightmap / layout a golden file was recorded with."""
    from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg
    N = G["N"]
    hs, tor = sim_stub.make_heightmap(*G["heightmap"])
    fg_cpu = sim_stub.FakeGym(N)
    oenv = EO.OracleEnv(K, N, hs, G["layout"], fg_cpu, None)
    fg_gpu = sim_stub.FakeGym(N, device=device)
    cfg = Lite3DTCCfg()
    cfg.env.num_envs = N
    cenv = LeggedRobotDTC(cfg, sim_device=device, gym=fg_gpu, height_samples=hs, terrain_origins=tor, layout=G["layout"], seed=G["seed"])
    return oenv, cenv, fg_cpu, fg_gpu
