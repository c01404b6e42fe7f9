"""Kernel census of one PPO.update() (20 VAE + 20 policy optimizer steps) with torch.profiler: launches and device time per kernel,
side streams running as in training.  python tools/update_kernels.py [envs]"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    env, fg, runner, state, pool_host, pool_dev = bench.build_world(N, 0, "cuda:0")
    runner.cuda_graph = False
    runner.learn(2)
    alg, wenv = runner.alg, runner.env
    od = wenv.get_observations()
    with torch.inference_mode():
        for _ in range(runner.num_steps_per_env):
            actions = alg.act(od["obs"], od["privileged_obs"], od["obs_history"], od["base_vel"], None)
            od, rewards, dones, infos = wenv.step(actions)
            alg.process_env_step(rewards, dones, next_obs=od["obs"], infos=infos)
        alg.compute_returns(od["obs"], od["privileged_obs"], od["base_vel"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        e0.record()
        alg.update()
        e1.record()
        torch.cuda.synchronize()
    cnt, tim = collections.Counter(), collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            cnt[e.name[:100]] += 1
            tim[e.name[:100]] += e.device_time
    rows = sorted(cnt, key=lambda k: -tim[k])
    print(json.dumps({"envs": N, "update_ms": round(e0.elapsed_time(e1), 2), "launches": sum(cnt.values()),
                      "sum_of_kernel_ms": round(sum(tim.values()) / 1e3, 2)}))
    for k in rows:
        print(f"{cnt[k]:6d} x {tim[k] / cnt[k]:8.1f} us = {tim[k] / 1e3:8.2f} ms  {k}")


if __name__ == "__main__":
    main()
