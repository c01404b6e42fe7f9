"""CPU-side checks of the C-ABI boundary: the shared library loads without a GPU, exports every symbol that
include/dtc_b200.h declares, and the ctypes struct mirrors have the library's sizes.  No compute call is made."""
import ctypes as C
import os
import re

import pytest
import torch

import dtc_b200  # noqa: F401
from dtc_b200 import _lib as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as G
    return C.CDLL(G.build())


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "dtc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dtc_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(B.EXPORTED_SYMBOLS), "include/dtc_b200.h and _lib.EXPORTED_SYMBOLS disagree"
    for sym in declared:
        assert hasattr(lib, sym), f"libdtc_b200.so does not export {sym}"


def test_struct_sizes_match(lib):
    for which, st in enumerate((B.EnvConfig, B.EnvBuffers, B.EnvNoise, B.Storage, B.PPOHParams, B.ParamInfo)):
        assert lib.dtc_struct_size(which) == C.sizeof(st), st.__name__


def test_param_table_covers_reference_state_dict(lib):
    """The flat layout holds every reference parameter exactly once (3 193 318 floats) and the two
    optimizer ranges are contiguous."""
    from dtc_b200.rsl_rl.modules.actor_critic_decoder import STATE_KEYS, _ParamTable, reference_init_state_dict
    t = _ParamTable.get()
    assert sorted(t.index) == sorted(STATE_KEYS)
    allidx = torch.cat([t.index[k] for k in STATE_KEYS])
    assert allidx.numel() == 3193318 and allidx.unique().numel() == allidx.numel()
    assert int(allidx.max()) < t.total
    sd = reference_init_state_dict()
    assert [tuple(v.shape) for v in sd.values()] == [t.shape[k] for k in STATE_KEYS]
    (v0, v1), (p0, p1) = t.ranges["vae"], t.ranges["policy"]
    assert v0 == 0 and p0 < v1 < p1
    vae_trained = [k for k in STATE_KEYS if k.startswith("vae.") and "memory_mlp" not in k and "gb_encoder" not in k]
    for k in STATE_KEYS:
        lo, hi = int(t.index[k].min()), int(t.index[k].max())
        assert (v0 <= lo and hi < v1) == (k in vae_trained), k
        in_policy = k == "std" or k.startswith(("actor_body", "critic_body", "vae.cenet_encoder", "vae.latent", "vae.terrain_encoder"))
        assert (p0 <= lo and hi < p1) == in_policy, k


def test_no_cpu_fallback():
    from dtc_b200.rsl_rl.modules import ActorCriticDecoder
    from dtc_b200.rsl_rl.storage import RolloutStorage
    with pytest.raises(B.DtcError):
        ActorCriticDecoder(53, 1389, 12).to("cpu")
    with pytest.raises(B.DtcError):
        RolloutStorage(4, 2, [53], [1389], [265], [12], device="cpu")


def test_reference_init_is_bit_identical_to_oracle():
    from dtc_b200.rsl_rl.modules import reference_init_state_dict
    from oracle import learner_oracle as LO
    torch.manual_seed(12)
    a = reference_init_state_dict()
    torch.manual_seed(12)
    b = LO.ActorCriticDecoder(53, 1389, 12).state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_nvtx_ranges_are_emitted(tmp_path):
    """Every hot C-ABI entry point opens an NVTX range named after itself (dtc_common.cuh: DTC_NVTX).  No profiler in this image, so
    a test-only injection library (tests/support/nvtx_inject.c, loaded by NVTX3 through NVTX_INJECTION64_PATH) records the names.
    The calls below fail their argument checks (no GPU needed) -- the range is opened before them and closed on return."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    inj = tmp_path / "libnvtx_inject.so"
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-I/usr/local/cuda/include", os.path.join(here, "support", "nvtx_inject.c"), "-o", str(inj)],
                   check=True)
    log = tmp_path / "nvtx.log"
    code = ("import ctypes as C, dtc_b200\nfrom dtc_b200 import _lib as B\nlib = B.lib()\n"
            "assert lib.dtc_env_state_prep(None, 0, 0, None, None) != 0\n"
            "assert lib.dtc_foothold_step(None, 6, None, None) != 0\n"
            "assert lib.dtc_env_observe(None, 0, 0, None, None) != 0\n")
    env = dict(os.environ, NVTX_INJECTION64_PATH=str(inj), DTC_NVTX_LOG=str(log), PYTHONPATH=os.path.dirname(here))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=os.path.dirname(here))
    lines = log.read_text().split("\n")
    assert lines[:6] == ["push dtc_env_state_prep", "pop", "push dtc_foothold_step", "pop", "push dtc_env_observe", "pop"], lines
