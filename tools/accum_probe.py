"""Error of the tensor-core GEMM against fp64 on ALL-POSITIVE operands (no cancellation: systematic errors show): with both TF32
companions, with one, with none; K-major and MN-major operands; split-K.  DTC_TC_DEBUG=16 puts the correction products into the main
TMEM accumulator."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dtc_b200
from dtc_b200 import _lib as B
lib = B.lib(); st = B.stream_ptr()
def lo(x):
    hi = (x.view(torch.int32) & -8192).view(torch.float32)
    r = x - hi
    return ((r.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
g = torch.Generator().manual_seed(0)
cases = ((24576, 512, 512, 1, 1, 1, "fwd  KK 512x512"), (24576, 512, 512, 1, 0, 1, "dgrad K/MN 512x512"), (512, 512, 1024, 0, 0, 1, "MN/MN K=1024 1 split"),
         (512, 512, 1024, 1, 1, 1, "KK    K=1024 1 split"), (512, 512, 24576, 0, 0, 27, "wgrad MN/MN K=24576 / 27"), (512, 512, 24576, 1, 1, 27, "KK K=24576 / 27"))
for (M, N, K, akc, bkc, splits, tag) in cases:
    A = (torch.randn(M if akc else K, K if akc else M, generator=g).abs() + 0.1).cuda()
    Bm = (torch.randn(N if bkc else K, K if bkc else N, generator=g).abs() * 0.1 + 0.01).cuda()
    Ar = (A if akc else A.T).double(); Br = (Bm if bkc else Bm.T).double()
    ref = Ar @ Br.T
    out = []
    for name, al, bl in (("both", lo(A), lo(Bm)), ("A_lo only", lo(A), None), ("none", None, None)):
        C = torch.empty(M, N, device="cuda")
        ws = torch.zeros(max(splits, 24) * M * N, device="cuda") if splits > 1 else None
        B.check(lib.dtc_gemm_debug(M, N, K, B.ptr(A), B.ptr(al), A.shape[1], akc, B.ptr(Bm), B.ptr(bl), Bm.shape[1], bkc, B.ptr(C), None, N, splits, B.ptr(ws), 1, st), "gemm")
        torch.cuda.synchronize()
        err = C.double() - ref
        out.append(f"{name}: {float((err / ref).mean()):+.2e}")
    print(f"{tag:26s} mean signed rel err  " + "   ".join(out), flush=True)
