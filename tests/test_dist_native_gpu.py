"""The multi-GPU C ABI (b200_dist_gemm / b200_dist_gemm_1d / b200_dist_trsm) on hardware: tests/dist_native_driver.py on one
GPU (one-rank communicator: the whole pipeline runs, the gathers are local) and, when the box has at least two GPUs, under
torchrun with NCCL between two (and four) ranks -- every rank's block is compared bit for bit with the single-GPU engine."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _check(stdout, world):
    lines = [json.loads(ln) for ln in stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == world, stdout[-2000:]
    for out in lines:
        bad = {k: v for k, v in out.items() if k.endswith("_ok") and v is not True}
        assert not bad, (bad, out)
        assert sum(1 for k in out if k.endswith("_ok")) >= 20
        assert out["launches"] > 20


def test_dist_entry_points_world_1():
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()))
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "dist_native_driver.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    _check(r.stdout, 1)


@pytest.mark.parametrize("world", [2, 4])
def test_dist_entry_points_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(_port()), str(ROOT / "tests" / "dist_native_driver.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    _check(r.stdout, world)
