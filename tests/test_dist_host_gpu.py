"""DistGemm.step_host (shards in pinned host memory) == DistGemm.step (device-resident shards) to 1e-14 absolute on O(1)
data (one ulp of epilogue rounding when another tile shape is selected for the column blocks), on one GPU with a one-rank
process group; tests/dist_host_driver.py is the same check under torchrun at any world size."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_step_host_matches_device_resident_step():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "dist_host_driver.py")], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["4steps_steps"] >= 4 and out["1step_steps"] >= 1
    for tag in ("4steps", "1step"):
        for rep in range(2):
            assert out[f"{tag}_rep{rep}_maxdiff"] <= 1e-14, out
