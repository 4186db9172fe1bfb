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


# ----------------------------------------------------------------------------------------------------------------------
# The boundary without symbol interposition ("make the boundary real")
GLUE_SUP = REFDIR / "libblis_b200_glue_sup.so"
CFG = REFDIR / "b200cfg"
BLAT_REF = REFDIR / "blastest_ref"
L3_LINE = re.compile(r"^blis_[sdcz](gemm|trsm|gemmt|syrk|herk|syr2k|her2k|hemm|symm|trmm|trmm3)_")
KERNELS = re.compile(r"libblis \(b200\): (\d+) CUDA kernels")


def _save(name, text):
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / name).write_text(text)


def test_gemm_served_by_the_gemmsup_handler_slot_only():
    """The plugin route proper (SURVEY 8b "whole-op gemm hook"): the glue flavour built WITHOUT any -DBLIS_B200_OVERRIDE_*
    defines no bli_<op>_ex symbol, so the unmodified libblis keeps its own bli_gemm_ex and reaches the engine only through
    the gemmsup_oft slot of its context (frame/3/bli_l3_oapi_ex.c:76-77 -> frame/3/bli_l3_sup.c:37-135 -> bli_gemmsup_b200).
    The reference testsuite's gemm rows (s/d/c/z, every transa/transb, row/column/general storage) must PASS and must
    have launched engine kernels."""
    _need(GLUE_SUP, TS / "test_libblis.x", TS / "input.general.n100", TS / "input.operations.gemm")
    nm = subprocess.run(["nm", "-D", "--defined-only", str(GLUE_SUP)], capture_output=True, text=True).stdout
    assert " bli_gemmsup_b200" in nm and not re.search(r" T bli_(gemm|trsm)_ex$", nm, re.M), "the sup flavour must not interpose bli_*_ex"
    env = dict(os.environ, LD_PRELOAD=str(GLUE_SUP), BLIS_B200_PLUGIN="1", BLIS_B200_VERBOSE="1")
    r = subprocess.run([str(TS / "test_libblis.x"), "-g", str(TS / "input.general.n100"), "-o", str(TS / "input.operations.gemm")],
                       capture_output=True, text=True, timeout=1200, env=env, cwd=str(TS))
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if re.match(r"^blis_[sdcz]gemm_", ln)]
    m = KERNELS.search(r.stderr)
    assert lines and m and int(m.group(1)) >= len(lines), "the handler slot was not used: " + r.stderr[-500:]
    bad = [ln for ln in lines if not ln.rstrip().endswith("PASS")]
    assert not bad, "\n".join(bad[:10])
    _save("testsuite_gemm_via_gemmsup_slot.b200.txt", "\n".join(lines) + "\n" + m.group(0) + "\n")


def _run_cfg_testsuite(general, operations):
    env = dict(os.environ, BLIS_B200_VERBOSE="1")
    env.pop("LD_PRELOAD", None)
    return subprocess.run([str(CFG / "test_libblis.x"), "-g", str(TS / general), "-o", str(operations)],
                          capture_output=True, text=True, timeout=1500, env=env, cwd=str(CFG))


def test_config_b200_library_passes_the_reference_testsuite():
    """Route 2 of INTEGRATION.md: libblis compiled WITH the b200 sub-configuration (tests/b200cfg_build.py), the
    reference's testsuite linked against it, no LD_PRELOAD and no run-time registration.  bli_arch_string() must say
    b200, all level-3 experiments must PASS, and they must have run on the engine."""
    _need(CFG / "test_libblis.x", REFDIR / "libblis_b200cfg.so", TS / "input.general.n100", TS / "input.operations.l3")
    ldd = subprocess.run(["ldd", str(CFG / "test_libblis.x")], capture_output=True, text=True).stdout
    assert "libblis_b200cfg.so" in ldd and "libblis_ref.so" not in ldd, ldd
    r = _run_cfg_testsuite("input.general.n100", TS / "input.operations.l3")
    assert r.returncode == 0, r.stderr[-3000:]
    assert re.search(r"^% active sub-configuration\s+b200\s*$", r.stdout, re.M), r.stdout[:1500]
    lines = [ln for ln in r.stdout.splitlines() if L3_LINE.match(ln)]
    m = KERNELS.search(r.stderr)
    assert len(lines) >= 3000 and m and int(m.group(1)) > len(lines), (len(lines), r.stderr[-500:])
    bad = [ln for ln in lines if not ln.rstrip().endswith("PASS")]
    assert not bad, "\n".join(bad[:10])
    _save("testsuite_config_b200_n100.txt", "\n".join(lines[::16]) + f"\n{len(lines)} experiments, all PASS\n{m.group(0)}\n"
          + "\n".join(ln for ln in r.stdout.splitlines() if ln.startswith("% active sub-conf") or ln.startswith("% version")) + "\n")


def test_config_b200_microkernel_slots_run_on_the_engine():
    """The kernel slots of the b200 context stay truthful (hard part 1): the reference testsuite's level-3
    MICROKERNEL modules (gemm_ukr, trsm_ukr, gemmtrsm_ukr: testsuite/src/test_gemm_ukr.c etc.) call whatever is registered
    under BLIS_GEMM_UKR / BLIS_GEMMTRSM_?_UKR / BLIS_TRSM_?_UKR with packed micropanels of the context's MR x NR; with the
    b200 context the gemm slot is bli_?gemm_b200_ukr (config/b200/bli_gemm_b200_ukr.c -> b200_gemm)."""
    _need(CFG / "test_libblis.x", CFG / "input.operations.ukr", TS / "input.general.n100")
    r = _run_cfg_testsuite("input.general.n100", CFG / "input.operations.ukr")
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if re.match(r"^blis_[sdcz](gemm|trsm|gemmtrsm)_ukr", ln)]
    m = KERNELS.search(r.stderr)
    assert lines and m and int(m.group(1)) >= len([ln for ln in lines if "gemm" in ln]), (len(lines), r.stderr[-500:])
    bad = [ln for ln in lines if not ln.rstrip().endswith("PASS")]
    assert not bad, "\n".join(bad[:10])
    _save("testsuite_config_b200_ukr.txt", "\n".join(lines) + "\n" + m.group(0) + "\n")


def _run_blat3(exe_dir, ch, env):
    out = exe_dir / f"out.{ch}blat3"
    if out.exists():
        out.unlink()
    with open(exe_dir / f"{ch}blat3.in") as fin:
        r = subprocess.run([str(exe_dir / f"{ch}blat3.x")], stdin=fin, capture_output=True, text=True, timeout=1500, env=env, cwd=str(exe_dir))
    text = out.read_text() if out.exists() else ""
    return r, text


def _check_blat3(ch, r, text, tag):
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    nops = {"s": 6, "d": 6, "c": 9, "z": 9}[ch]          # gemm symm trmm trsm syrk syr2k (+ hemm herk her2k)
    assert len(re.findall(r"PASSED THE TESTS OF ERROR-EXITS", text)) == nops, text[-2500:]
    assert len(re.findall(r"PASSED THE COMPUTATIONAL TESTS", text)) == nops, text[-2500:]
    assert not re.search(r"FAIL|SUSPECT|ILLEGAL|\*\*\*\*\*\*", text), text[-2500:]
    m = KERNELS.search(r.stderr)
    calls = sum(int(x) for x in re.findall(r"\(\s*(\d+) CALLS\)", text))
    assert m and int(m.group(1)) > calls // 4, ("the engine was not used", r.stderr[-400:], calls)
    _save(f"blastest_{ch}blat3_{tag}.txt", text + m.group(0) + "\n")


@pytest.mark.parametrize("ch", ["d", "s", "z", "c"])
def test_netlib_blat3_on_config_b200(ch):
    """blastest/src/?blat3.c (the netlib level-3 testers: error exits through xerbla + exact computational checks at the
    ?gemm_/?trsm_/... boundary, Makefile:863-940) linked against libblis_b200cfg.so -- no LD_PRELOAD."""
    _need(CFG / f"{ch}blat3.x", CFG / f"{ch}blat3.in", REFDIR / "libblis_b200cfg.so")
    env = dict(os.environ, BLIS_B200_VERBOSE="1")
    env.pop("LD_PRELOAD", None)
    r, text = _run_blat3(CFG, ch, env)
    _check_blat3(ch, r, text, "config_b200")


@pytest.mark.parametrize("ch", ["d", "z"])
def test_netlib_blat3_on_the_plugin(ch):
    """The same testers linked against the UNMODIFIED reference, with the plugin preloaded (route 1)."""
    _need(BLAT_REF / f"{ch}blat3.x", GLUE)
    env = dict(os.environ, LD_PRELOAD=str(GLUE), BLIS_B200_PLUGIN="1", BLIS_B200_VERBOSE="1")
    r, text = _run_blat3(BLAT_REF, ch, env)
    _check_blat3(ch, r, text, "plugin")


_OBJ_BUFFER_SCRIPT = r"""
import ctypes as C, json, sys
import numpy as np
blis = C.CDLL(sys.argv[1], mode=C.RTLD_GLOBAL)          # pulls in libblis_b200.so through its RUNPATH
blis.bli_init()
blis.bli_arch_string.restype = C.c_char_p; blis.bli_arch_query_id.restype = C.c_int
blis.bli_malloc_user.restype = C.c_void_p; blis.bli_malloc_user.argtypes = [C.c_size_t, C.POINTER(C.c_int)]
blis.bli_free_user.argtypes = [C.c_void_p]
blis.b200_pointer_kind.argtypes = [C.c_void_p]; blis.b200_launch_count.restype = C.c_uint64
n = 512; err = C.c_int(0)
ptrs = [blis.bli_malloc_user(8 * n * n, C.byref(err)) for _ in range(3)]
kinds = [blis.b200_pointer_kind(p) for p in ptrs]
a, b, c = (np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n * n,)).reshape(n, n, order="F") for p in ptrs)
rng = np.random.default_rng(7)
a[:] = rng.integers(-3, 4, (n, n)); b[:] = rng.integers(-3, 4, (n, n)); c[:] = rng.integers(-3, 4, (n, n))
want = 2.0 * (a @ b) + 3.0 * c                           # small integers: exact in fp64 whatever the summation order
l0 = blis.b200_launch_count()
i = C.c_int(n); al = C.c_double(2.0); be = C.c_double(3.0)
blis.dgemm_(b"N", b"N", C.byref(i), C.byref(i), C.byref(i), C.byref(al), C.c_void_p(ptrs[0]), C.byref(i), C.c_void_p(ptrs[1]), C.byref(i),
            C.byref(be), C.c_void_p(ptrs[2]), C.byref(i))
heap = np.zeros(16); pageable = blis.b200_pointer_kind(heap.ctypes.data)
out = dict(arch=blis.bli_arch_string(blis.bli_arch_query_id()).decode(), kinds=kinds, pageable=pageable, exact=bool((c == want).all()),
           launches=int(blis.b200_launch_count() - l0))
for p in ptrs: blis.bli_free_user(p)
print(json.dumps(out))
"""


def test_config_b200_obj_buffers_are_page_locked():
    """SURVEY 8f rank 4 / 8b "memory-allocation hooks": under config/b200, BLIS_MALLOC_USER is b200_malloc_pinned
    (config/b200/bli_family_b200.h), so what bli_obj_create hands out (bli_malloc_user, frame/base/bli_obj.c:190;
    frame/base/bli_malloc.c) is page-locked and the engine's copies take the direct-DMA path: b200_pointer_kind says 1
    for such a buffer (2 for ordinary heap memory), and dgemm_ on them is served by the engine, exactly."""
    _need(REFDIR / "libblis_b200cfg.so")
    env = dict(os.environ); env.pop("LD_PRELOAD", None)
    r = subprocess.run([sys.executable, "-c", _OBJ_BUFFER_SCRIPT, str(REFDIR / "libblis_b200cfg.so")], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["arch"] == "b200" and out["kinds"] == [1, 1, 1] and out["pageable"] == 2, out
    assert out["exact"] and out["launches"] >= 1, out
