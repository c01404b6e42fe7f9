"""Kernel census of the rollout loop (PPO.act + env.step + process_env_step) with torch.profiler: which kernels run per environment
step and whether any of them is an ATen kernel (the loop should launch only libdtc_b200.so kernels)."""
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
    steps = 4
    env, fg, runner, state, pool_host, pool_dev = bench.build_world(N, 0, "cuda:0")
    runner.learn(2)
    alg, wenv = runner.alg, runner.env
    od = wenv.get_observations()
    obs, priv, hist = od["obs"], od["privileged_obs"], od["obs_history"]
    alg.storage.clear()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        with torch.inference_mode():
            for _ in range(steps):
                actions = alg.act(obs, priv, hist, od["base_vel"], None)
                od, rewards, dones, infos = wenv.step(actions)
                obs, priv, hist = od["obs"], od["privileged_obs"], od["obs_history"]
                alg.process_env_step(rewards, dones, next_obs=od["obs"], infos=infos)
        torch.cuda.synchronize()
    cnt, tim = collections.Counter(), collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            cnt[e.name[:90]] += 1
            tim[e.name[:90]] += e.device_time
    rows = sorted(cnt, key=lambda k: -tim[k])
    aten = [k for k in rows if "at::native" in k or "native::" in k]
    print(json.dumps({"envs": N, "steps": steps, "launches_per_step": sum(v for k, v in cnt.items() if "Memcpy" not in k and "Memset" not in k) / steps,
                      "device_us_per_step": round(sum(tim.values()) / steps, 1), "aten_kernels": aten}))
    for k in rows:
        print(f"{cnt[k] / steps:6.2f}/step {tim[k] / steps:8.1f} us/step  {k}")


if __name__ == "__main__":
    main()
