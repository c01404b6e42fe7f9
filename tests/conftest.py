import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the built library: skip (not fail) elsewhere, e.g. a plain `pytest tests` on a CPU box."""
    import torch
    lib = os.path.join(ROOT, "deep-tracking-control_b200", "libdtc_b200.so")
    why = None
    if not torch.cuda.is_available():
        why = "no CUDA device"
    elif not os.path.exists(lib):
        why = "libdtc_b200.so not built (python __graft_entry__.py)"
    if why:
        skip = pytest.mark.skip(reason=why)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)
