"""Data-parallel host logic on CPU: world_size 2 over gloo.

What must hold for N-rank training to equal 1-rank training on the concatenated batch:
  * advantage normalisation from all-reduced {sum, sumsq, count} == normalisation of the concatenated advantages;
  * SUM all-reduce of per-rank gradients of batch-MEAN losses, scaled by 1/world, == gradient on the concatenated minibatch
    (checked with the oracle's critic value loss, which has no batch-global statistics inside);
  * the KL sum riding in the piggy-back slot, divided by the global row count, == KL mean of the concatenated minibatch;
  * environment sharding / per-rank seeds are disjoint and cover everything.
The CUDA kernels themselves are exercised by the -m gpu tests; the same helpers (rsl_rl/utils/dp.py) carry both backends."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dtc_b200  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dtc_b200.rsl_rl.utils import dp
        from oracle import learner_oracle as LO
        torch.manual_seed(3)  # identical replicas
        ac = LO.ActorCriticDecoder(53, 1389, 12)
        g = torch.Generator().manual_seed(11)
        rows = 64
        obs, priv, bv = torch.randn(rows, 53, generator=g), torch.randn(rows, 1389, generator=g), torch.randn(rows, 3, generator=g)
        ret, adv_raw = torch.randn(rows, 1, generator=g), torch.randn(rows, generator=g) * 3 + 1
        kl_rows = torch.rand(rows, generator=g)
        b, e = dp.shard_envs(rows)
        assert (e - b) == rows // world
        # --- gradient averaging
        loss = (ac.evaluate(obs[b:e], priv[b:e], bv[b:e]) - ret[b:e]).pow(2).mean()
        loss.backward()
        flat = torch.cat([p.grad.flatten() for p in ac.critic_body.parameters()])
        piggy = kl_rows[b:e].sum().reshape(1)
        buf = torch.cat([flat, piggy])
        dp.allreduce_sum_(buf)
        grad_avg, kl_mean = buf[:-1] / world, buf[-1] / rows
        # --- advantage moments
        x = adv_raw[b:e].double()
        st3 = torch.tensor([x.sum(), (x * x).sum(), float(x.numel())], dtype=torch.float64)
        dp.combine_moments_(st3)
        mean, std = dp.moments_to_mean_std(st3)
        if rank == 0:
            ac.zero_grad()
            (ac.evaluate(obs, priv, bv) - ret).pow(2).mean().backward()
            ref = torch.cat([p.grad.flatten() for p in ac.critic_body.parameters()])
            out_q.put(dict(grad_err=float((grad_avg - ref).abs().max() / ref.abs().max()),
                           kl_err=abs(float(kl_mean) - float(kl_rows.mean())),
                           mean_err=abs(mean - float(adv_raw.double().mean())), std_err=abs(std - float(adv_raw.double().std())),
                           seeds=[dp.rank_seed(1000), dp.world_size()]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_equivalence():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["grad_err"] < 1e-5, res
    assert res["kl_err"] < 1e-6 and res["mean_err"] < 1e-12 and res["std_err"] < 1e-12, res
    assert res["seeds"] == [1000, 2]


def test_single_process_helpers_are_identity():
    from dtc_b200.rsl_rl.utils import dp
    assert dp.world_size() == 1 and dp.rank() == 0
    t = torch.arange(5.0)
    assert torch.equal(dp.allreduce_sum_(t.clone()), t)
    assert dp.shard_envs(4096) == (0, 4096)
    m, s = dp.moments_to_mean_std(torch.tensor([10.0, 30.0, 4.0], dtype=torch.float64))
    assert m == 2.5 and abs(s - (5.0 / 3.0) ** 0.5) < 1e-12
