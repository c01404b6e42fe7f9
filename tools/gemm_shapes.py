"""Per-shape table of the GEMM family inside one training iteration: every launch timed alone with CUDA events (the step's side
streams serialised), grouped by (kernel, M, N, K, operand layout).  Usage: python tools/gemm_shapes.py [envs]"""
import collections
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = tempfile.mktemp(suffix=".txt")
os.environ["DTC_PROF_DUMP"] = path
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    import ctypes as C
    from dtc_b200 import _lib as B
    env, fg, runner, state, pool_host, pool_dev = bench.build_world(N, 0, "cuda:0")
    runner.learn(3)
    lib = B.lib()
    lib.dtc_profile_enable(1)
    runner.learn(1)
    torch.cuda.synchronize()
    z = C.c_double()
    n = C.c_int64()
    lib.dtc_profile_read(C.byref(z), C.byref(z), C.byref(n), C.byref(z), C.byref(n))
    lib.dtc_profile_enable(0)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for line in open(path):
        kind, m, n_, k, layout, ms = line.split()
        if int(kind) == 1:
            continue
        key = (int(layout), int(m), int(n_), int(k))
        agg[key][0] += 1
        agg[key][1] += float(ms)
    tot = sum(v[1] for v in agg.values())
    print(f"GEMM family: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms serialised per iteration")
    print("layout: 1xx SIMT, 0x single-CTA tcgen05, 1x CTA-pair tcgen05; +1000*splits; low digit = 2*A_mn_major + B_mn_major")
    for key, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        layout, m, n_, k = key
        tf = 2.0 * m * n_ * k * cnt / (ms * 1e-3) / 1e12
        print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{cnt:4d}  {ms / cnt * 1e3:8.1f} us  {tf:7.1f} TFLOP/s  layout {layout:5d}  M={m} N={n_} K={k}")
    os.unlink(path)


if __name__ == "__main__":
    main()
