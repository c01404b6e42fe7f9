"""Diagnostic: run one VAE step on both GEMM engines and report the first activation / gradient buffer that differs."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dtc_b200
from dtc_b200 import _lib as B
from tests import test_learner_gpu as TL
N = int(sys.argv[1]) if len(sys.argv) > 1 else 171
lib = B.lib()
oac, cac, rng = TL._make_policies(5)
lib.dtc_set_gemm_mode(0)
oalg, calg = TL._fill_storages(N, 24, 9, oac, cac, rng)
mbs = N * 24 // 4
perm = torch.randperm(N * 24, generator=torch.Generator().manual_seed(1))
batch = calg.storage.gather(perm.cuda())
hp = calg._hparams()
h = cac._learner(mbs)
eps = torch.randn(mbs, 16, generator=torch.Generator().manual_seed(2)).cuda()
names = ["H1", "E", "ML", "T1", "T2", "XD", "D1", "D2", "REC", "U1", "U2", "HR", "dREC", "dHR", "dD2", "dD1", "dU2", "dU1", "dX", "dML", "dE", "dH1", "dT2", "dT1"]
res = {}
for mode in (0, 1):
    lib.dtc_set_gemm_mode(mode)
    B.check(lib.dtc_vae_step(h, C.byref(batch._c), 0, mbs, B.ptr(eps), 0, 0, C.byref(hp), 1, B.stream_ptr()), "vae")
    torch.cuda.synchronize()
    res[mode] = {n: cac.debug_buffer(n).clone() for n in names}
    res[mode]["grads"] = cac._grads.clone()
for n in names + ["grads"]:
    a, b = res[0][n].double(), res[1][n].double()
    d = (a - b).abs()
    i = int(d.argmax())
    print(f"{n:6s} max|simt| {a.abs().max().item():.3e}  max diff {d.max().item():.3e}  rel {d.max().item() / max(a.abs().max().item(), 1e-30):.2e}  at {i // a.shape[-1] if a.dim() == 2 else i},{i % a.shape[-1] if a.dim() == 2 else 0}")
idx = cac._idx["vae.terrain_encoder.0.weight"]
for nm in ("vae.terrain_encoder.0.weight", "vae.terrain_encoder.2.weight", "vae.terrain_decoder.4.weight", "vae.cenet_encoder.0.weight"):
    idx = cac._idx[nm]
    a, b = res[0]["grads"][idx].double(), res[1]["grads"][idx].double()
    sc = a.abs().max().item()
    e = (a - b).abs() / sc
    qs = torch.quantile(e, torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], dtype=torch.float64, device=e.device))
    print(nm, "scale", f"{sc:.3e}", "err/scale quantiles 50/90/99/99.9/100:", [f"{q:.2e}" for q in qs.tolist()], "frac>2e-5:", float((e > 2e-5).double().mean()))
