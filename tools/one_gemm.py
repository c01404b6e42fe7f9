"""One tensor-core GEMM launched four times (ncu capture target): python tools/one_gemm.py M N K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dtc_b200
from dtc_b200 import _lib as B
lib = B.lib(); st = B.stream_ptr()
M, N, K = (int(a) for a in sys.argv[1:4])
r4 = lambda x: (x + 3) // 4 * 4
lo = lambda x: x - (x.view(torch.int32) & -8192).view(torch.float32)
A = torch.randn(M, r4(K), device="cuda"); W = torch.randn(N, r4(K), device="cuda")
Cc = torch.empty(M, r4(N), device="cuda"); Cl = torch.empty_like(Cc)
Al, Wl = lo(A), lo(W)
for _ in range(4):
    lib.dtc_gemm_debug(M, N, K, B.ptr(A), B.ptr(Al), A.shape[1], 1, B.ptr(W), B.ptr(Wl), W.shape[1], 1, B.ptr(Cc), B.ptr(Cl), Cc.shape[1], 1, None, 1, st)
torch.cuda.synchronize()
