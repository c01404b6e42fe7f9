"""Microbenchmark of the GEMM family on the learner's shapes (TFLOP/s, CUDA events, inputs > L2 not needed: compute bound)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtc_b200
from dtc_b200 import _lib as B

def bench(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3

def main():
    lib = B.lib(); st = B.stream_ptr()
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 24576
    out = {}
    for (N, K) in ((512, 693), (512, 512), (512, 752), (256, 512), (693, 512), (128, 256), (64, 128)):
        r4 = lambda x: (x + 3) // 4 * 4
        A = torch.randn(M, r4(K), device="cuda"); W = torch.randn(N, r4(K), device="cuda"); b = torch.randn(N, device="cuda")
        Cc = torch.empty(M, r4(N), device="cuda")
        t = bench(lambda: lib.dtc_linear_forward(M, N, K, B.ptr(A), A.shape[1], B.ptr(W), W.shape[1], B.ptr(b), 1, B.ptr(Cc), Cc.shape[1], st))
        out[f"fwd_{N}x{K}"] = round(2 * M * N * K / t / 1e12, 2)
        dY = torch.randn(M, r4(N), device="cuda"); dX = torch.empty(M, r4(K), device="cuda")
        lo = lambda x: x - (x.view(torch.int32) & -8192).view(torch.float32)
        Al, Wl, dYl = lo(A), lo(W), lo(dY)
        Cl = torch.empty_like(Cc)
        for mode, tag in ((0, "simt"), (1, "tc")):
            t = bench(lambda: lib.dtc_gemm_debug(M, N, K, B.ptr(A), B.ptr(Al), A.shape[1], 1, B.ptr(W), B.ptr(Wl), W.shape[1], 1, B.ptr(Cc), B.ptr(Cl), Cc.shape[1], 1, None, mode, st))
            out[f"{tag}_fwd_{N}x{K}"] = round(2 * M * N * K / t / 1e12, 2)
            t = bench(lambda: lib.dtc_gemm_debug(M, K, N, B.ptr(dY), B.ptr(dYl), dY.shape[1], 1, B.ptr(W), B.ptr(Wl), W.shape[1], 0, B.ptr(dX), None, dX.shape[1], 1, None, mode, st))
            out[f"{tag}_dgrad_{N}x{K}"] = round(2 * M * N * K / t / 1e12, 2)
            tiles = ((N + 127) // 128) * ((K + 127) // 128)
            splits = min(128, max(-(-592 // tiles), -(-M // 1024)))  # as dtc_gemm_pick_splits: >= ceil(M/32/32) for the TMEM accumulation limit
            ws = torch.empty(splits * N * r4(K), device="cuda"); dW = torch.empty(N, r4(K), device="cuda")
            t = bench(lambda: lib.dtc_gemm_debug(N, K, M, B.ptr(dY), B.ptr(dYl), dY.shape[1], 0, B.ptr(A), B.ptr(Al), A.shape[1], 0, B.ptr(dW), None, dW.shape[1], splits, B.ptr(ws), mode, st))
            out[f"{tag}_wgrad_{N}x{K}"] = round(2 * M * N * K / t / 1e12, 2)
        t = bench(lambda: torch.mm(A, W.t()))
        out[f"torch_mm_{N}x{K}"] = round(2 * M * N * r4(K) / t / 1e12, 2)
    print(json.dumps({"M": M, "tflops": out}))

if __name__ == "__main__":
    main()
