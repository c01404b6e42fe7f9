"""Build-container only: wall time of ONE learn() iteration of the UNMODIFIED reference (through oracle/ref_harness) against the
oracle port (oracle/*.py, what bench.py's CPU arm times) on the same synthetic states - the ratio that maps the port's
env-steps/s to the real reference's.  /root/reference does not exist on the GPU box, so this number is measured here and quoted.

    python tests/tools/port_vs_reference.py [N=1024] [iters=2]
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dtc_b200  # noqa: E402,F401
from dtc_b200 import sim_stub  # noqa: E402
import bench  # noqa: E402


def reference_iterations(N, iters, seed=1000):
    from oracle import ref_harness as RH
    RH.import_reference()
    from rsl_rl.runners import OnPolicyRunner
    from legged_gym.envs.lite3.lite3_dtc_config import Lite3DTCCfgPPO
    from legged_gym.utils.helpers import class_to_dict
    torch.manual_seed(1)
    np.random.seed(1)
    hs, tor = sim_stub.make_heightmap("stones", 0)
    layout = sim_stub.initial_env_layout(N, tor, seed)
    fg = sim_stub.FakeGym(N)
    g = torch.Generator().manual_seed(seed)
    pool = [sim_stub.synth_state(N, layout[2], g) for _ in range(4)]
    st = {"i": 0}

    def source():
        st["i"] = (st["i"] + 1) % 4
        return pool[st["i"]]

    fg.source = source
    env = RH.build_ref_env(N, hs, tor, layout, fg)
    with tempfile.TemporaryDirectory() as d, contextlib.redirect_stdout(io.StringIO()):
        runner = OnPolicyRunner(env, class_to_dict(Lite3DTCCfgPPO()), log_dir=d, device="cpu")
        runner.learn(1)  # warm-up
        t0 = time.perf_counter()
        runner.learn(iters)
        dt = time.perf_counter() - t0
    return N * 24 * iters / dt, dt


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    v_ref, dt_ref = reference_iterations(N, iters)
    v_port, dt_port, _ = bench._cpu_iterations(N, iters, warmup=1)
    print(json.dumps({"envs": N, "iters": iters, "cores": cores, "reference_env_steps_per_s": round(v_ref, 1),
                      "port_env_steps_per_s": round(v_port, 1), "port_over_reference": round(v_port / v_ref, 3)}))


if __name__ == "__main__":
    main()
