"""Data-parallel plumbing: environments shard by rank, parameters are replicated, and each optimizer step
all-reduces ONE flat gradient range; the advantage moments are all-reduced once per iteration.  Backend-agnostic
(`nccl` on the GPUs, `gloo` in the CPU tests): these helpers only touch torch.distributed."""
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def allreduce_sum_(flat, group=None):
    """In-place SUM all-reduce of a flat gradient range (the caller applies 1/world through grad_scale)."""
    if world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def combine_moments_(stats3, group=None):
    """stats3 = [sum(x), sum(x^2), count] of the local raw advantages (float64).  After the SUM all-reduce every rank
    holds the global triple, from which mean / unbiased std follow exactly as rollout_storage.py:152 computes them on
    the concatenated batch."""
    if world_size(group) > 1:
        dist.all_reduce(stats3, op=dist.ReduceOp.SUM, group=group)
    return stats3


def moments_to_mean_std(stats3):
    s, q, n = (float(v) for v in stats3[:3])
    mean = s / n
    var = max((q - n * mean * mean) / (n - 1.0), 0.0)
    return mean, var ** 0.5


def rank_seed(base_seed, group=None):
    """Environment-side seed of this rank (policy initialisation uses the SAME seed on every rank)."""
    return int(base_seed) + rank(group)


def shard_envs(total_envs, group=None):
    """Contiguous [begin, end) environment range of this rank."""
    w, r = world_size(group), rank(group)
    per = total_envs // w
    return r * per, (r + 1) * per if r < w - 1 else total_envs
