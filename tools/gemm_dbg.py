import json, os, sys
sys.path.insert(0, "/root/repo")
import torch, dtc_b200
from dtc_b200 import _lib as B
lib = B.lib(); st = B.stream_ptr()
def bench(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
M = 24576
out = {}
lo = lambda x: x - (x.view(torch.int32) & -8192).view(torch.float32)
for (N, K) in ((512, 693), (512, 512), (256, 512), (128, 256), (64, 128)):
    r4 = lambda x: (x + 3) // 4 * 4
    A = torch.randn(M, r4(K), device="cuda"); W = torch.randn(N, r4(K), device="cuda")
    Cc = torch.empty(M, r4(N), device="cuda"); Cl = torch.empty_like(Cc)
    Al, Wl = lo(A), lo(W)
    t = bench(lambda: lib.dtc_gemm_debug(M, N, K, B.ptr(A), B.ptr(Al), A.shape[1], 1, B.ptr(W), B.ptr(Wl), W.shape[1], 1, B.ptr(Cc), B.ptr(Cl), Cc.shape[1], 1, None, 1, st))
    out[f"fwd_{N}x{K}"] = round(2 * M * N * K / t / 1e12, 1)
    dY = torch.randn(M, r4(N), device="cuda"); dX = torch.empty(M, r4(K), device="cuda"); dYl = lo(dY)
    t = bench(lambda: lib.dtc_gemm_debug(M, K, N, B.ptr(dY), B.ptr(dYl), dY.shape[1], 1, B.ptr(W), B.ptr(Wl), W.shape[1], 0, B.ptr(dX), None, dX.shape[1], 1, None, 1, st))
    out[f"dgrad_{N}x{K}"] = round(2 * M * N * K / t / 1e12, 1)
print(os.environ.get("DTC_TC_DEBUG", "0"), json.dumps(out))
