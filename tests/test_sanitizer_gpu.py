"""compute-sanitizer memcheck over every kernel family, kept as a test (the reference has no race / memory
checking; VERDICT round 1 asked for the ad-hoc run to become one).  tools/sanitize.sh all adds racecheck / synccheck / initcheck."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sanitizer():
    return shutil.which("compute-sanitizer") or ("/usr/local/cuda/bin/compute-sanitizer" if os.path.exists("/usr/local/cuda/bin/compute-sanitizer") else None)


@pytest.mark.timeout(1500)
def test_memcheck_clean():
    cs = _sanitizer()
    if cs is None:
        pytest.skip("compute-sanitizer not installed")
    r = subprocess.run([cs, "--tool", "memcheck", "--error-exitcode", "9", "--print-limit", "20", sys.executable,
                        os.path.join(ROOT, "tools", "sanitize_target.py"), "all"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=1400)
    out = r.stdout.decode(errors="replace")
    assert "sanitize target ok" in out, out[-3000:]
    assert r.returncode == 0 and "ERROR SUMMARY: 0 errors" in out, out[-3000:]
