"""Foothold-scoring kernel microbench (BASELINE.json configs[4]): 16384 envs, stepping-stone map, HBM GB/s vs roofline.
Algorithmic bytes per launch = 3048 B/env + 3,942,400 B heightmap."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dtc_b200  # noqa: E402
from dtc_b200 import _lib as B, sim_stub  # noqa: E402
from dtc_b200.legged_gym.envs import LeggedRobotDTC, Lite3DTCCfg  # noqa: E402


def run(N=16384, iters=50, warmup=5, variants=(6,)):
    dev = "cuda"
    hs, tor = sim_stub.make_heightmap("stones", 0)
    layout = sim_stub.initial_env_layout(N, tor, 1)
    fg = sim_stub.FakeGym(N, device=dev)
    cfg = Lite3DTCCfg()
    cfg.env.num_envs = N
    env = LeggedRobotDTC(cfg, sim_device=dev, gym=fg, height_samples=hs, terrain_origins=tor, layout=layout, seed=1)
    g = torch.Generator(device=dev).manual_seed(2)
    fg.load(sim_stub.synth_state(N, env.env_origins, g, device=dev))
    env.reset()
    # one full environment step on a fresh synthetic state, then restore that state: reset_idx rewrote the root states of the
    # environments that terminated, and the kernel should see what it sees inside the training loop (root / thigh / command
    # tensors of one consistent simulator state)
    st0 = sim_stub.synth_state(N, env.env_origins, g, device=dev)
    fg.queue.append(st0)
    env.step(torch.zeros(N, 12, device=dev))
    fg.load(st0)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    out = {}
    st = B.stream_ptr()
    for v in variants:
        ts = []
        for i in range(warmup + iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            B.check(env.lib.dtc_foothold_step(env._h, v, C.c_void_p(0), st), "foothold")
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        med = ts[len(ts) // 2]
        bytes_ = 3048 * N + 3942400
        out[f"variant{v}"] = dict(us_median=med, us_min=ts[0], GBps=bytes_ / med / 1e3, frac=bytes_ / med / 1e3 / peak)
    # warm-L2 timing too (back to back, no flush)
    for v in variants:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            B.check(env.lib.dtc_foothold_step(env._h, v, C.c_void_p(0), st), "foothold")
        e1.record()
        torch.cuda.synchronize()
        out[f"variant{v}"]["us_warm"] = e0.elapsed_time(e1) * 1e3 / iters
    return out


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    variants = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (6,)
    print(json.dumps({"N": N, **run(N, variants=variants)}))
