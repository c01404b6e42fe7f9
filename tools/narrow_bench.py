"""Per-launch time (us) of the narrow-layer GEMMs in isolation (back-to-back launches, CUDA events), against their HBM floor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtc_b200
from dtc_b200 import _lib as B

def bench(fn, iters=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def main():
    lib = B.lib(); st = B.stream_ptr()
    r4 = lambda x: (x + 3) // 4 * 4
    lo = lambda x: x - (x.view(torch.int32) & -8192).view(torch.float32)
    Ms = [int(a) for a in sys.argv[1:]] or [24576]
    for M in Ms:
        for (N, K) in ((128, 256), (256, 128), (128, 64), (64, 128), (64, 35), (35, 64), (128, 12), (256, 512), (512, 256), (512, 512)):
            A = torch.randn(M, r4(K), device="cuda"); W = torch.randn(N, r4(K), device="cuda")
            Cc = torch.empty(M, r4(N), device="cuda"); Cl = torch.empty_like(Cc)
            Al, Wl = lo(A), lo(W)
            t = bench(lambda: lib.dtc_gemm_debug(M, N, K, B.ptr(A), B.ptr(Al), A.shape[1], 1, B.ptr(W), B.ptr(Wl), W.shape[1], 1, B.ptr(Cc), B.ptr(Cl), Cc.shape[1], 1, None, 1, st))
            t2 = bench(lambda: lib.dtc_gemm_debug(M, N, K, B.ptr(A), B.ptr(Al), A.shape[1], 1, B.ptr(W), B.ptr(Wl), W.shape[1], 1, B.ptr(Cc), None, Cc.shape[1], 1, None, 1, st))
            t3 = bench(lambda: lib.dtc_gemm_debug(M, N, K, B.ptr(A), None, A.shape[1], 1, B.ptr(W), None, W.shape[1], 1, B.ptr(Cc), None, Cc.shape[1], 1, None, 2, st))
            byt = (M * r4(K) * 8 + M * r4(N) * 8 + N * r4(K) * 8)
            print(f"M={M:6d} N={N:4d} K={K:4d}  {t:7.1f} us   hbm floor {byt / 6.5e6:6.1f} us   {2*M*N*K/t/1e6:6.1f} TF | no C_lo: {t2:7.1f} us {2*M*N*K/t2/1e6:6.1f} TF | in-SM split: {t3:7.1f} us {2*M*N*K/t3/1e6:6.1f} TF", flush=True)

if __name__ == "__main__":
    main()
