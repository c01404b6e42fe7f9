"""Diagnostic: step-by-step gradient comparison of PPO.update() against the oracle."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dtc_b200
from dtc_b200 import _lib as B
from tests import test_learner_gpu as TL
DEV = "cuda"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
oac, cac, rng = TL._make_policies(5)
T = 24
oalg, calg = TL._fill_storages(N, T, 9, oac, cac, rng)
mbs = N * T // 4
lib = B.lib()
oalg.debug = {}
p_before = {k: v.clone() for k, v in oac.state_dict().items()}
o_ret = oalg.update()
log = rng.take()
perm = log[0][1]
draws = [v for tag, v in log[1:]]
eps = []
for k in range(8):
    eps += [draws[3 * k].to(DEV).contiguous(), draws[3 * k + 1].to(DEV).contiguous()]
batch = calg.storage.gather(perm.to(DEV))
hp = calg._hparams()
h = cac._learner(mbs)
calg._push_lr(h)
st = B.stream_ptr()
def rel(got, ref):
    worst = (0, "")
    for k, g in ref.items():
        a = got[k].cpu().double(); b = g.double()
        e = ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
        if e > worst[0]: worst = (e, k)
    return worst
k = 0
for epoch in range(2):
    for i in range(4):
        B.check(lib.dtc_vae_step(h, C.byref(batch._c), i * mbs, mbs, B.ptr(eps[2 * k]), 0, 0, C.byref(hp), 1, st), "vae")
        got = TL._grads_as_state_dict(cac)
        w = rel({kk[4:]: v for kk, v in got.items() if kk.startswith("vae.")}, oalg.debug["vae_grads"][k])
        B.check(lib.dtc_optimizer_apply(h, 0, C.byref(hp), 1.0, mbs, st), "apply")
        B.check(lib.dtc_ppo_step(h, C.byref(batch._c), i * mbs, mbs, B.ptr(eps[2 * k + 1]), 0, 0, C.byref(hp), 1, st), "ppo")
        got = TL._grads_as_state_dict(cac)
        w2 = rel(got, oalg.debug["ppo_grads"][k])
        B.check(lib.dtc_optimizer_apply(h, 1, C.byref(hp), 1.0, mbs, st), "apply")
        s = cac.stats().tolist()
        print(f"step {k}: vae grad worst rel {w[0]:.2e} ({w[1]})  ppo grad worst rel {w2[0]:.2e} ({w2[1]})  lr {s[8]:.6g} vs {oalg.debug['ppo_losses'][k][4]:.6g} kl {s[7]:.6g} vs {oalg.debug['ppo_losses'][k][3]:.6g} gn {s[9]:.5g} {s[10]:.5g}")
        k += 1
sd = cac.state_dict()
for kk, v in oac.state_dict().items():
    moved = (v - p_before[kk]).abs().max().item()
    d = (sd[kk].cpu() - v).abs()
    if moved > 0:
        print(f"{kk:40s} moved {moved:.3e} maxdiff {d.max().item():.3e} meandiff {d.mean().item():.3e}")
