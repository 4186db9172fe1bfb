"""bench.py --impl reference (CPU only, no GPU needed): the contract of the reference arm.

Under torchrun only rank 0 works and prints; the other ranks exit 0 without output.  Rank 0's line carries the same
metric/config as the b200 arm plus `impl`, `cpu_baseline` and an `e2e` object that repeats the line's own value with
zero transfer bytes."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(rank: int, world: int = 2, timeout: int = 300):
    env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), B200_REF_BUDGET="3")    # bounded sample: CPU-only container
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", str(world), "--steps", "1",
                           "--warmup", "3"], capture_output=True, text=True, timeout=timeout, env=env, cwd=str(ROOT))


def test_other_ranks_exit_without_work():
    r = _run(rank=1)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_rank0_prints_the_contract_line():
    if not (ROOT / "oracle" / "_ref" / "libblis_ref.so").exists() and not Path("/root/reference").exists():
        pytest.skip("reference library not built")
    r = _run(rank=0)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GFLOPS" and line["higher_is_better"] is True
    assert line["n_gpus"] == 2 and line["value"] > 0
    assert "dgemm m=n=k=16384" in line["config"]["workload"]
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # both halves of the metric: dtrsm (T1) rides in the same line
    assert line["dtrsm"]["value"] > 0 and line["dtrsm"]["cpu_baseline"]["kind"] == "reference"
    assert "sub-config" in line["cpu_baseline"]["sample"]
