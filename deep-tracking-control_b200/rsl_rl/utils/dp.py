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


class PeerAllReduce:
    """SUM all-reduce of a flat float32 CUDA range over NVLink peer memory (include/dtc_b200.h: dtc_dp_*; csrc/dtc_dp.cu): three
    launches on the caller's stream, device-side flag handshakes, no NCCL call.  torch.distributed is only the channel that carries
    the two CUDA IPC handles of every rank to the others, once.  `available()` says whether this process group can use it (CUDA,
    world size 2..16, one node)."""

    def __init__(self, max_floats, device, group=None):
        import ctypes as C
        from ... import _lib as B
        self._B, self._C = B, C
        self.world, self.rank, self.device = world_size(group), rank(group), torch.device(device)
        lib = B.lib()
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            B.check(lib.dtc_dp_create(self.rank, self.world, int(max_floats), C.byref(self._h)), "dtc_dp_create")
            hb, hf = C.create_string_buffer(64), C.create_string_buffer(64)
            B.check(lib.dtc_dp_handles(self._h, hb, hf), "dtc_dp_handles")
            mine = (self.rank, hb.raw, hf.raw)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            for r, b, f in everyone:
                if r != self.rank:
                    B.check(lib.dtc_dp_open(self._h, int(r), b, f), "dtc_dp_open")
        dist.barrier(group=group)  # every rank has mapped every buffer before the first flag is raised
        self.max_floats = int(max_floats)
        self._group = group

    def register(self, flat):
        """Maps `flat` (this rank's own buffer, same size on every rank; e.g. the flat gradient buffer) into every other rank, so
        that all-reduces of its sub-ranges run as one in-place kernel.  Collective: every rank calls it once."""
        B, C = self._B, self._C
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
        hb, off = C.create_string_buffer(64), C.c_int64()
        rc = B.lib().dtc_dp_register(self._h, B.ptr(flat), flat.numel(), hb, C.byref(off))
        mine = (self.rank, hb.raw, int(off.value), flat.numel()) if rc == 0 else None  # e.g. memory that has no IPC handle
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self._group)
        if any(e is None for e in everyone):
            return False  # every rank takes the same decision: the exchange-buffer path stays in use
        assert all(e[3] == flat.numel() for e in everyone), "every rank must register a range of the same size"
        with torch.cuda.device(self.device):
            for r, h, o, _ in everyone:
                if r != self.rank:
                    B.check(B.lib().dtc_dp_open_registered(self._h, int(r), h, o), "dtc_dp_open_registered")
        dist.barrier(group=self._group)
        self._registered = flat  # keeps the tensor (and with it the peers' view of this memory) alive
        return True

    @staticmethod
    def available(device, group=None):
        w = world_size(group)
        return torch.device(device).type == "cuda" and 2 <= w <= 16 and dist.get_backend(group) == "nccl"

    def allreduce_sum_(self, flat):
        B = self._B
        n = flat.numel()
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and n % 4 == 0
        B.check(B.lib().dtc_dp_allreduce(self._h, B.ptr(flat), n, B.stream_ptr(self.device)), "dtc_dp_allreduce")
        return flat

    def check(self):
        """Synchronises the current stream; raises if any rank ever timed out waiting for a peer."""
        B = self._B
        B.check(B.lib().dtc_dp_error(self._h, B.stream_ptr(self.device)), "dtc_dp_error")

    def close(self):
        if self._h:
            self._B.lib().dtc_dp_destroy(self._h)
            self._h = None
