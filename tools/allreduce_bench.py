"""Latency of the gradient all-reduce (7.4 / 7.8 MB flat fp32 ranges): peer-memory kernel (csrc/dtc_dp.cu) vs NCCL.
torchrun --nproc-per-node N tools/allreduce_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import dtc_b200  # noqa: F401
from dtc_b200.rsl_rl.utils import dp

def main():
    rank = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
    dev = torch.device(f"cuda:{rank}")
    peer = dp.PeerAllReduce(1 << 21, dev)
    big = torch.zeros(1 << 21, device=dev)
    for n in (1 << 16, 1 << 19, 1851000 // 4 * 4, 1950000 // 4 * 4):
        x = torch.randn(n, device=dev)
        out = {}
        y = big[:n]
        for name, fn in (("nccl", lambda: dist.all_reduce(x)), ("peer", lambda: peer.allreduce_sum_(x)), ("inplace", lambda: peer.allreduce_sum_(y))):
            if name == "inplace" and not getattr(peer, "_registered", None) is big:
                peer.register(big)
            for _ in range(10): fn()
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100): fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 100 * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[name] = float(t)
        if rank == 0:
            print(f"world {dist.get_world_size()}  n = {n:8d} floats ({n * 4 / 1e6:5.2f} MB)   nccl {out['nccl']:7.1f} us   peer (exchange buffers, 3 launches) {out['peer']:7.1f} us   peer in-place (1 launch) {out['inplace']:7.1f} us", flush=True)
    peer.check(); peer.close()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
