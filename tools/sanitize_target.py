"""Small pass over every kernel family of libdtc_b200.so, meant to run UNDER compute-sanitizer (tests/test_sanitizer_gpu.py,
tools/sanitize.sh): env step with each foothold variant (incl. the TMA-staged default), terrain paint, policy act, one runner
iteration (rollout, GAE, VAE + PPO minibatch steps, optimizer) eagerly and as a replayed CUDA graph, GRU forward.  Prints
"sanitize target ok" when it ran to the end; memory errors are the sanitizer's to report."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dtc_b200  # noqa: E402,F401
from dtc_b200 import sim_stub  # noqa: E402
from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg, Lite3DTCCfgPPO  # noqa: E402
from dtc_b200.legged_gym.envs.lite3.lite3_dtc_config import class_to_dict  # noqa: E402
from dtc_b200.legged_gym.utils.terrain import Terrain  # noqa: E402
from dtc_b200.rsl_rl.modules.memory import Memory  # noqa: E402
from dtc_b200.rsl_rl.runners import OnPolicyRunner  # noqa: E402

DEV = "cuda:0"


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    N = 64
    cfg = Lite3DTCCfg()
    cfg.env.num_envs = N
    if what in ("all", "terrain"):
        Terrain(cfg.terrain, N, device=DEV)  # generators -> rectangle lists -> dtc_terrain_paint
    hs, tor = sim_stub.make_heightmap("stones", 0)
    layout = sim_stub.initial_env_layout(N, tor, 3)
    g = torch.Generator().manual_seed(3)
    pool = [{k: v.to(DEV) for k, v in sim_stub.synth_state(N, layout[2], g).items()} for _ in range(8)]
    pool[2]["root_states"][1, 3:7] = torch.tensor([0.9, 0.0, 0.0, 0.435], device=DEV)  # a flipped robot: device-side reset
    pool[5]["root_states"][2, 0] += 500.0                                               # a robot off the map: clamped patch
    fg = sim_stub.FakeGym(N, device=DEV)
    st = {"i": 0}

    def source():
        st["i"] = (st["i"] + 1) % 8
        return pool[st["i"]]

    fg.source = source
    env = LeggedRobotDTC(cfg, sim_device=DEV, gym=fg, height_samples=hs, terrain_origins=tor, layout=layout, seed=3)
    if what in ("all", "env"):
        for variant in (0, 4, 5, 6):
            env.foothold_variant = variant
            for _ in range(3):
                env.step(torch.randn(N, 12, device=DEV))
        env.foothold_variant = 6
    if what in ("all", "learn"):
        torch.manual_seed(1)
        runner = OnPolicyRunner(env, class_to_dict(Lite3DTCCfgPPO()), log_dir=None, device=DEV)
        runner.learn(3)  # eager, capture + first replay, replay
        assert runner._graph is not None
        m = Memory(53, type="gru", hidden_size=64, num_layers=1, device=DEV)
        m.forward(torch.randn(N, 53, device=DEV))
    torch.cuda.synchronize()
    print("sanitize target ok")


if __name__ == "__main__":
    main()
