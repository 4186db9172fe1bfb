"""Engine lifetime: b200_finalize releases every stream, event, staging buffer and scheduler counter, and the next call
re-initialises lazily with the same results (the reference's bli_init/bli_finalize pair may be cycled by an application,
frame/base/bli_init.c:87-99).  Runs in its own process so that the session-wide engine is left alone."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_finalize_then_reinit():
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "lifecycle_driver.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["before"] < 1e-12 and out["host_before"] < 1e-12, out
    for cycle in range(3):
        assert out[f"same_bits_{cycle}"], out
        assert out[f"host_{cycle}"] < 1e-12 and out[f"batch_{cycle}"] < 1e-12, out
    assert out["launches"] >= 4 * 2 + 3 * 3
