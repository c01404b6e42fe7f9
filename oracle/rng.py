"""TEST INFRASTRUCTURE ONLY (oracle package) - random-draw providers.

The reference draws randomness at eight call sites on the hot path (SURVEY.md "RNG parity").  Parity is
only definable with *injected* draws, so both the reference run (tests/golden/make_golden.py patches
torch / numpy entry points with `Recorder`) and the oracle restatement (which calls `rng.<fn>` in the
reference's call order) consume the same logged tensors through `Replay`.  `Live` draws fresh values
from a seeded torch.Generator and is what the CPU-baseline timing uses.
"""
import numpy as np
import torch


class Live:
    def __init__(self, seed=0):
        self.g = torch.Generator().manual_seed(seed)
        self.np = np.random.default_rng(seed)

    def rand(self, *shape):
        return torch.rand(*shape, generator=self.g)

    def rand_like(self, t):
        return torch.rand(t.shape, generator=self.g, dtype=t.dtype)

    def randn_like(self, t):
        return torch.randn(t.shape, generator=self.g, dtype=t.dtype)

    def randint_like(self, t, high):
        return torch.randint(0, int(high), t.shape, generator=self.g, dtype=t.dtype)

    def randperm(self, n):
        return torch.randperm(n, generator=self.g)

    def np_randint(self, lo, hi):
        return int(self.np.integers(lo, hi))

    def np_normal(self, mu, sigma):
        return float(self.np.normal(mu, sigma))


class Replay:
    """Pops (tag, value) records written by `Recorder` in order, checking the tag."""

    def __init__(self, log):
        self.log = list(log)
        self.pos = 0

    def _pop(self, tag):
        assert self.pos < len(self.log), f"replay exhausted at {tag}"
        t, v = self.log[self.pos]
        assert t == tag, f"replay order mismatch: wanted {tag}, log has {t} at {self.pos}"
        self.pos += 1
        return v

    def rand(self, *shape):
        v = self._pop("rand")
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        assert tuple(v.shape) == tuple(shape), (v.shape, shape)
        return v.clone()

    def rand_like(self, t):
        v = self._pop("rand_like")
        assert v.shape == t.shape
        return v.clone()

    def randn_like(self, t):
        v = self._pop("randn_like")
        assert v.shape == t.shape, (v.shape, t.shape)
        return v.clone()

    def randint_like(self, t, high):
        v = self._pop("randint_like")
        assert v.shape == t.shape
        return v.clone()

    def randperm(self, n):
        v = self._pop("randperm")
        assert v.numel() == n
        return v.clone()

    def np_randint(self, lo, hi):
        return int(self._pop("np_randint"))

    def np_normal(self, mu, sigma):
        return float(self._pop("np_normal"))

    def done(self):
        return self.pos == len(self.log)


class Recorder:
    """Context manager: patches torch.rand / rand_like / randn_like / randint_like / randperm / normal and
    np.random.randint / normal so every draw of the reference is appended to `self.log`."""

    def __init__(self):
        self.log = []

    def __enter__(self):
        self._saved = dict(rand=torch.rand, rand_like=torch.rand_like, randn_like=torch.randn_like,
                           randint_like=torch.randint_like, randperm=torch.randperm, normal=torch.normal,
                           np_randint=np.random.randint, np_normal=np.random.normal)
        s = self._saved
        log = self.log

        def rand(*a, **k):
            v = s["rand"](*a, **k)
            log.append(("rand", v.clone()))
            return v

        def rand_like(t, **k):
            v = s["rand_like"](t, **k)
            log.append(("rand_like", v.clone()))
            return v

        def randn_like(t, **k):
            v = s["randn_like"](t, **k)
            log.append(("randn_like", v.clone()))
            return v

        def randint_like(t, *a, **k):
            v = s["randint_like"](t, *a, **k)
            log.append(("randint_like", v.clone()))
            return v

        def randperm(n, **k):
            v = s["randperm"](n, **k)
            log.append(("randperm", v.clone()))
            return v

        def normal(mean, std, **k):
            # torch.distributions.Normal.sample: torch.normal(mean.expand(shape), std.expand(shape));
            # restated as mean + std * eps with a logged standard-normal eps (ATen CPU: normal_(0,1)*std+mean)
            eps = s["randn_like"](mean)
            log.append(("randn_like", eps.clone()))
            return mean + std * eps

        def np_randint(lo, hi=None, *a, **k):
            v = s["np_randint"](lo, hi, *a, **k)
            log.append(("np_randint", int(v)))
            return v

        def np_normal(mu=0.0, sigma=1.0, *a, **k):
            v = s["np_normal"](mu, sigma, *a, **k)
            log.append(("np_normal", float(v)))
            return v

        torch.rand, torch.rand_like, torch.randn_like = rand, rand_like, randn_like
        torch.randint_like, torch.randperm, torch.normal = randint_like, randperm, normal
        np.random.randint, np.random.normal = np_randint, np_normal
        return self

    def __exit__(self, *exc):
        s = self._saved
        torch.rand, torch.rand_like, torch.randn_like = s["rand"], s["rand_like"], s["randn_like"]
        torch.randint_like, torch.randperm, torch.normal = s["randint_like"], s["randperm"], s["normal"]
        np.random.randint, np.random.normal = s["np_randint"], s["np_normal"]
        return False

    def take(self):
        out, self.log[:] = list(self.log), []
        return out
