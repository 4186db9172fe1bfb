"""The drop-in boundary end to end on a GPU: the REAL reference libblis (built by
oracle/build_ref.py) with the glue of blis_b200/blis_glue/ on top.

1. in a fresh process the plugin is registered and the reference's own entry points
   (dgemm_, cblas_dgemm, bli_?gemm, dtrsm_, bli_?trsm) are called on host arrays: results must be
   right AND must have been computed by the engine (b200_launch_count grows);
2. the reference's own testsuite binary (testsuite/src, unmodified) is run with
   LD_PRELOAD=libblis_b200_glue.so BLIS_B200_PLUGIN=1: every line of the eleven served level-3 operations must say PASS
   (thresholds testsuite/src/test_gemm.c:44-47, test_trsm.c:44-47), same count as the CPU run.
"""
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
REFDIR = ROOT / "oracle" / "_ref"
GLUE = REFDIR / "libblis_b200_glue.so"
TS = REFDIR / "testsuite"


def _need(*paths):
    for p in paths:
        if not Path(p).exists():
            pytest.skip(f"{p} not built (needs /root/reference at build time)")


def test_reference_entry_points_run_on_the_engine():
    _need(GLUE, REFDIR / "libblis_ref.so")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "dropin_driver.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["launches_after_dgemm_"] >= 1, "dgemm_ did not reach the CUDA engine"
    assert out["launches_trsm"] >= 5, "trsm did not reach the CUDA engine"
    for k in ("dgemm_", "cblas_dgemm", "bli_dgemm", "bli_zgemm", "dtrsm_", "bli_dtrsm", "bli_ztrsm"):
        assert out[k] < 1e-11, (k, out[k])
    for k in ("bli_sgemm", "bli_cgemm", "bli_strsm", "bli_ctrsm"):
        assert out[k] < 2e-3, (k, out[k])
    assert out["launches_gemmt_family"] >= 6, "the gemmt family did not reach the CUDA engine"
    for k in ("dsyrk_", "bli_dgemmt", "bli_zherk", "bli_zher2k", "bli_dsyr2k"):
        assert out[k] < 1e-11, (k, out[k])
    assert out["bli_cherk"] < 2e-3
    assert out["launches_symm_trmm"] >= 4, "symm/trmm did not reach the CUDA engine"
    assert out["launches_gemm_batch"] >= 5 and out["dgemm_batch_"] < 1e-11, ("dgemm_batch_", out["launches_gemm_batch"], out["dgemm_batch_"])
    for k in ("dsymm_", "dtrmm_"):
        assert out[k] < 1e-11, (k, out[k])


def _run_testsuite(general, preload, operations="input.operations.l3"):
    env = dict(os.environ)
    if preload:
        env.update(LD_PRELOAD=str(GLUE), BLIS_B200_PLUGIN="1", BLIS_B200_VERBOSE="1")
    r = subprocess.run([str(TS / "test_libblis.x"), "-g", str(TS / general), "-o", str(TS / operations)],
                       capture_output=True, text=True, timeout=1200, env=env, cwd=str(TS))
    lines = [ln for ln in r.stdout.splitlines() if re.match(r"^blis_[sdcz](gemm|trsm|gemmt|syrk|herk|syr2k|her2k|hemm|symm|trmm|trmm3)_", ln)]
    return r, lines


# n100t4: the reference testsuite's own concurrency mode ("Simulate application-level threading: 4"): four testsuite
# threads call the BLIS API at the same time, so the engine's shared state (stream, tile-scheduler counters, staging
# ring, stream-ordered workspace) is exercised by concurrent callers (SURVEY.md 8b "Threading").
@pytest.mark.parametrize("general", ["input.general.n100", "input.general.n1000d", "input.general.n100t4"])
def test_reference_testsuite_passes_on_the_engine(general):
    _need(GLUE, TS / "test_libblis.x", TS / general)
    r, lines = _run_testsuite(general, preload=True)
    assert r.returncode == 0, r.stderr[-3000:]
    m = re.search(r"libblis \(b200\): (\d+) CUDA kernels", r.stderr)
    assert m and int(m.group(1)) > len(lines), "the engine was not used: " + r.stderr[-500:]
    assert lines, r.stdout[-2000:]
    bad = [ln for ln in lines if not ln.rstrip().endswith("PASS")]
    assert not bad, "\n".join(bad[:10])
    # same set of experiments as the plain CPU run of the same binary
    _, cpu_lines = _run_testsuite(general, preload=False)
    names, cpu_names = [ln.split()[0] for ln in lines], [ln.split()[0] for ln in cpu_lines]
    if general.endswith("t4"):           # the testsuite threads print in whatever order they finish
        names, cpu_names = sorted(names), sorted(cpu_names)
    assert names == cpu_names
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"testsuite_{general}.b200.txt").write_text("\n".join(lines) + "\n" + (m.group(0) if m else ""))


def test_reference_testsuite_mixed_datatype_gemm_on_the_engine():
    """The reference testsuite's mixed-domain + mixed-precision gemm sweep (every A/B/C datatype combination and both
    computation precisions, testsuite/src/test_gemm.c) with bli_gemm_ex bound to the engine: all PASS, same experiments
    as the CPU run."""
    general = "input.general.n100mixed"
    _need(GLUE, TS / "test_libblis.x", TS / general, TS / "input.operations.gemm")
    r, lines = _run_testsuite(general, preload=True, operations="input.operations.gemm")
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if re.match(r"^blis_[sdcz]+gemm_", ln)]
    m = re.search(r"libblis \(b200\): (\d+) CUDA kernels", r.stderr)
    assert m and int(m.group(1)) > len(lines), "the engine was not used: " + r.stderr[-500:]
    assert len(lines) > 10000, r.stdout[-2000:]
    bad = [ln for ln in lines if not ln.rstrip().endswith("PASS")]
    assert not bad, "\n".join(bad[:10])
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "testsuite_mixed.b200.txt").write_text("\n".join(lines[::97]) + f"\n{len(lines)} experiments, all PASS\n" + m.group(0))
