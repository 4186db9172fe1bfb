#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the REAL reference BLIS.

Run in the build container (needs /root/reference and oracle/_ref/libblis_ref.so
from oracle/build_ref.py):

    python tests/golden/make_golden.py

Writes (all small, committed):
  index_arith.json  bli_determine_blocksize / bli_thread_range_sub /
                    bli_thread_partition_2x2 results (bit-exact spec for the
                    tile scheduler and the multi-GPU block splits)
  packm.npz         micropanels packed by bli_??packm_struc_cxk for general and
                    triangular (lower/upper, unit, inverted diag, conj, ragged) panels
  gemm.npz          bli_?gemm outputs for the cases in `gemm_cases()`
  trsm.npz          bli_?trsm outputs for the cases in `trsm_cases()`
  gemm_md.npz       bli_gemm on objects of mixed datatypes (tests/ref_shim.c) for `gemm_md_cases()`
  strucmm.npz       bli_?hemm / symm / trmm3 / trmm outputs for `strucmm_cases()` (unstored triangle of A NaN-poisoned)
  gemmt.npz         bli_?gemmt / syrk / herk / syr2k / her2k outputs for `gemmt_cases()` (unstored triangle NaN-poisoned)
Inputs are not stored: tests rebuild them with tests/gen.py (integer-hash
generators, platform independent).
"""
from __future__ import annotations

import ctypes as C
import itertools
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(ROOT / "oracle"))

import gen  # noqa: E402
from refblis import (CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE, LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, RIGHT,  # noqa: E402
                     TRANSPOSE, UNIT_DIAG, UPPER, RefBlis)

i64, vp, ci = C.c_int64, C.c_void_p, C.c_int


# ----------------------------------------------------------------------------- case lists (shared with tests)
def index_cases():
    db = []
    for backward in (0, 1):
        for dim in list(range(1, 41)) + [100, 257, 1000, 4080, 16384, 32768]:
            for b_alg in (1, 4, 6, 8, 72, 128, 256, 4080):
                for b_max in (b_alg, b_alg + b_alg // 4):
                    if dim // b_alg <= 48:
                        db.append((backward, dim, b_alg, b_max))
    tr = []
    for n_way in range(1, 10):
        for n in (0, 1, 7, 8, 63, 64, 65, 100, 1000, 1023, 16384, 32768, 65536, 8192 + 5):
            for bf in (1, 4, 6, 8, 128):
                for low in (0, 1):
                    tr.append((n_way, n, bf, low))
    p2 = []
    for nt in range(1, 65):
        for w1, w2 in ((1000, 1000), (16384, 16384), (32768, 8192), (100, 4000), (65536, 65536), (3600, 3600), (7, 5000)):
            p2.append((nt, w1, w2))
    return db, tr, p2


def packm_cases():
    """(ch, tri, uplo, unit, conj, invdiag, panel_dim, panel_len, panel_dim_max, panel_len_max,
        panel_dim_off, panel_len_off, kappa, order)"""
    cases = []
    for ch in "sdcz":
        mr = {"s": 4, "d": 6, "c": 3, "z": 4}[ch]
        kap = 1.0
        kapc = (0.5 - 2.0j) if ch in "cz" else 2.0
        for order in "cr":
            # general panels: full, ragged in the short dim, ragged+padded in the long dim, conj, kappa
            cases += [(ch, 0, 0xE0, 0, 0, 0, mr, 9, mr, 9, 0, 0, kap, order),
                      (ch, 0, 0xE0, 0, 1, 0, mr - 1, 7, mr, 8, 0, 0, kapc, order),
                      (ch, 0, 0xE0, 0, 0, 0, 1, 1, mr, 2, 0, 0, kap, order)]
            for uplo in (LOWER, UPPER):
                for unit, conj, inv in itertools.product((0, 1), (0, 1), (0, 1)):
                    # diagonal block in the middle of the panel (p10 | p11 | p12 all present)
                    cases.append((ch, 1, uplo, unit, conj, inv, mr, 3 * mr, mr, 3 * mr, mr, 0, kap, order))
                # diagonal at the start / at the end, ragged last panel with identity extension
                cases += [(ch, 1, uplo, 0, 0, 1, mr, 2 * mr, mr, 2 * mr, 0, 0, kap, order),
                          (ch, 1, uplo, 0, 0, 1, mr, 2 * mr, mr, 2 * mr, mr, 0, kap, order),
                          (ch, 1, uplo, 0, 1, 1, mr - 1, 2 * mr - 1, mr, 2 * mr, mr, 0, kapc, order),
                          (ch, 1, uplo, 1, 0, 1, 1, 1, mr, mr, 0, 0, kap, order)]
    return cases


def gemm_cases():
    """(ch, kind, m, n, k, transa, transb, oa, ob, oc, alpha, beta)"""
    cases = []
    for ch in "sdcz":
        cx = ch in "cz"
        al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))   # testsuite/src/test_gemm.c:213-214
        trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
        for ta, tb in itertools.product(trs, trs):
            cases.append((ch, "frac", 19, 23, 31, ta, tb, "c", "c", "c", al, be))
        cases += [
            (ch, "frac", 100, 100, 100, NO_TRANSPOSE, NO_TRANSPOSE, "c", "c", "c", al, be),   # input.general.fast size
            (ch, "frac", 60, 40, 50, NO_TRANSPOSE, NO_TRANSPOSE, "r", "r", "r", al, be),
            (ch, "frac", 37, 5, 300, TRANSPOSE, NO_TRANSPOSE, "r", "c", "g", al, be),         # k > KC, general-stride C
            (ch, "frac", 1, 64, 7, NO_TRANSPOSE, TRANSPOSE, "g", "r", "c", al, 0.0),          # beta == 0
            (ch, "frac", 65, 1, 1, NO_TRANSPOSE, NO_TRANSPOSE, "c", "g", "r", 1.0, 1.0),
            (ch, "frac", 8, 9, 10, NO_TRANSPOSE, NO_TRANSPOSE, "c", "c", "c", 0.0, be),       # alpha == 0
            (ch, "pow2", 48, 40, 64, NO_TRANSPOSE, NO_TRANSPOSE, "c", "c", "c", 2.0, 0.5),     # exact
            (ch, "pow2", 33, 17, 50, TRANSPOSE, TRANSPOSE, "r", "c", "r", -1.0, 1.0),          # exact
        ]
    return cases


def trsm_cases():
    """(ch, kind, m, n, side, uplo, trans, diag, oa, ob, alpha)"""
    cases = []
    for ch in "sdcz":
        cx = ch in "cz"
        al = (2.0 + 0.3j) if cx else 2.0                                # testsuite/src/test_trsm.c:209 uses 2.0
        trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
        for side, uplo, tr, dg in itertools.product((LEFT, RIGHT), (LOWER, UPPER), trs, (NONUNIT_DIAG, UNIT_DIAG)):
            cases.append((ch, "frac", 23, 11, side, uplo, tr, dg, "c", "c", al))
            cases.append((ch, "ints", 14, 9, side, uplo, tr, dg, "r", "c", 2.0))      # exact
        cases += [
            (ch, "frac", 100, 100, LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, "c", "c", 2.0),   # input.general.fast size
            (ch, "frac", 48, 64, RIGHT, UPPER, NO_TRANSPOSE, NONUNIT_DIAG, "r", "r", 2.0),
            (ch, "frac", 300, 13, LEFT, UPPER, TRANSPOSE, NONUNIT_DIAG, "c", "g", al),        # m > KC
            (ch, "frac", 1, 1, LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, "c", "c", al),
            (ch, "frac", 5, 40, LEFT, LOWER, NO_TRANSPOSE, UNIT_DIAG, "c", "c", 0.0),          # alpha == 0
        ]
    return cases


def gemmt_cases():
    """(ch, op, kind, m, k, uplo, transa, transb, oa, ob, oc, alpha, beta); op in gemmt/syrk/herk/syr2k/her2k.
    Scalars as in testsuite/src/test_{gemmt,syrk,herk,syr2k,her2k}.c (alpha 2.0-ish, beta 1.2-ish; herk's alpha and
    beta and her2k's beta are real)."""
    cases = []
    for ch in "sdcz":
        cx = ch in "cz"
        al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
        trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
        for op in ("gemmt", "syrk", "herk", "syr2k", "her2k"):
            a_, b_ = (2.0 if op == "herk" else al), (1.2 if op in ("herk", "her2k") else be)
            for uplo, ta in itertools.product((LOWER, UPPER), trs):
                tb = trs[(trs.index(ta) + 1) % len(trs)]
                cases.append((ch, op, "frac", 21, 13, uplo, ta, tb, "c", "c", "c", a_, b_))
            cases += [
                (ch, op, "frac", 100, 100, LOWER, NO_TRANSPOSE, NO_TRANSPOSE, "c", "c", "c", a_, b_),  # input.general.fast size
                (ch, op, "frac", 45, 300, UPPER, TRANSPOSE, NO_TRANSPOSE, "r", "c", "r", a_, b_),        # k > KC, row-stored C
                (ch, op, "frac", 33, 9, LOWER, NO_TRANSPOSE, TRANSPOSE, "g", "r", "g", a_, b_),          # general strides
                (ch, op, "frac", 17, 6, UPPER, NO_TRANSPOSE, NO_TRANSPOSE, "c", "c", "c", a_, 0.0),      # beta == 0
                (ch, op, "frac", 9, 5, LOWER, NO_TRANSPOSE, NO_TRANSPOSE, "c", "c", "c", 0.0, b_),       # alpha == 0
                (ch, op, "frac", 1, 4, LOWER, NO_TRANSPOSE, NO_TRANSPOSE, "c", "c", "c", a_, b_),
                (ch, op, "pow2", 40, 32, LOWER, NO_TRANSPOSE, TRANSPOSE, "c", "c", "c", 2.0, 0.5),       # exact
                (ch, op, "pow2", 29, 48, UPPER, TRANSPOSE, NO_TRANSPOSE, "r", "c", "r", -1.0, 1.0),      # exact
            ]
    return cases


def gemmt_inputs(case, idx):
    """A, B (None for syrk/herk) and C with the triangle that must not be touched filled with NaN."""
    ch, op, kind, m, k, uplo, ta, tb, oa, ob, oc, al, be = case
    am, ak = (k, m) if ta & TRANSPOSE else (m, k)
    a = gen.matrix(ch, am, ak, 19 * idx + 1, kind, oa, pad=3)
    b = None
    if op == "gemmt":
        bm, bn = (m, k) if tb & TRANSPOSE else (k, m)
        b = gen.matrix(ch, bm, bn, 19 * idx + 2, kind, ob, pad=1)
    elif op in ("syr2k", "her2k"):
        bm, bn = (k, m) if tb & TRANSPOSE else (m, k)
        b = gen.matrix(ch, bm, bn, 19 * idx + 2, kind, ob, pad=1)
    c = gen.matrix(ch, m, m, 19 * idx + 3, kind, oc, pad=2)
    gen.poison_unstored(c, uplo == LOWER)
    return a, b, c


def gemmt_run(impl, case, a, b, c):
    ch, op, kind, m, k, uplo, ta, tb, oa, ob, oc, al, be = case
    if op in ("syrk", "herk"):
        getattr(impl, op)(uplo, ta, al, a, be, c)
    else:
        getattr(impl, op)(uplo, ta, tb, al, a, b, be, c)


def strucmm_cases():
    """(ch, op, kind, m, n, side, uplo, transa, diag, transb, oa, ob, oc, alpha, beta); op in hemm/symm/trmm3/trmm.
    For hemm/symm transa carries conja only and diag is ignored; trmm ignores transb, ob/oc describe B, beta unused."""
    cases = []
    for ch in "sdcz":
        cx = ch in "cz"
        al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
        trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
        for op in ("hemm", "symm", "trmm3", "trmm"):
            tas = trs if op in ("trmm3", "trmm") else ((NO_TRANSPOSE, CONJ_NO_TRANSPOSE) if cx else (NO_TRANSPOSE,))
            for side, uplo, ta in itertools.product((LEFT, RIGHT), (LOWER, UPPER), tas):
                tb = trs[(trs.index(ta) + 1) % len(trs)] if ta in trs else NO_TRANSPOSE
                dg = UNIT_DIAG if (uplo == UPPER) == (side == LEFT) else NONUNIT_DIAG
                cases.append((ch, op, "frac", 23, 11, side, uplo, ta, dg, tb, "c", "c", "c", al, be))
            cases += [
                (ch, op, "frac", *((100, 100) if ch == "d" else (36, 44)), LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, NO_TRANSPOSE, "c", "c", "c", al, be),   # d: input.general.fast size
                (ch, op, "frac", 12, 300, RIGHT, UPPER, NO_TRANSPOSE, NONUNIT_DIAG, TRANSPOSE, "r", "c", "r", al, be),       # n > KC on the structured side
                (ch, op, "frac", 33, 9, LEFT, UPPER, NO_TRANSPOSE, UNIT_DIAG, NO_TRANSPOSE, "g", "r", "g", al, be),          # general strides
                (ch, op, "frac", 17, 6, RIGHT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, NO_TRANSPOSE, "c", "c", "c", al, 0.0),     # beta == 0
                (ch, op, "frac", 9, 5, LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, NO_TRANSPOSE, "c", "c", "c", 0.0, be),       # alpha == 0
                (ch, op, "frac", 1, 4, LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, NO_TRANSPOSE, "c", "c", "c", al, be),
                (ch, op, "pow2", 40, 32, LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, TRANSPOSE, "c", "c", "c", 2.0, 0.5),       # exact
                (ch, op, "pow2", 29, 48, RIGHT, UPPER, TRANSPOSE, UNIT_DIAG, NO_TRANSPOSE, "r", "c", "r", -1.0, 1.0),        # exact
            ]
    return cases


def strucmm_inputs(case, idx):
    """A (only the uplo triangle valid: the other one is NaN), B and C; for trmm C is None (B is updated in place)."""
    ch, op, kind, m, n, side, uplo, ta, dg, tb, oa, ob, oc, al, be = case
    ma = m if side == LEFT else n
    a = gen.matrix(ch, ma, ma, 23 * idx + 1, kind, oa, pad=1)
    gen.poison_unstored(a, uplo == LOWER)
    if op == "trmm":
        return a, gen.matrix(ch, m, n, 23 * idx + 2, kind, oc, pad=2), None
    bm, bn = (n, m) if tb & TRANSPOSE else (m, n)
    b = gen.matrix(ch, bm, bn, 23 * idx + 2, kind, ob, pad=1)
    c = gen.matrix(ch, m, n, 23 * idx + 3, kind, oc, pad=2)
    return a, b, c


def strucmm_run(impl, case, a, b, c):
    """Runs the case; returns the array that holds the result (C, or B for trmm)."""
    ch, op, kind, m, n, side, uplo, ta, dg, tb, oa, ob, oc, al, be = case
    if op in ("hemm", "symm"):
        getattr(impl, op)(side, uplo, ta & CONJ_NO_TRANSPOSE, tb, al, a, b, be, c)
    elif op == "trmm3":
        impl.trmm3(side, uplo, ta, dg, tb, al, a, b, be, c)
    else:
        impl.trmm(side, uplo, ta, dg, al, a, b)
        return b
    return c


def gemm_md_cases():
    """(cha, chb, chc, comp_prec, kind, m, n, k, transa, transb, oc, alpha, beta): every combination of the four storage
    datatypes for A, B, C and both computation precisions (docs/MixedDatatypes.md; testsuite/input.operations.mixed)."""
    cases = []
    for cha, chb, chc in itertools.product("sdcz", repeat=3):
        for cp in (0, 2):
            i = len(cases)
            ta = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE)[i % 4]
            tb = (NO_TRANSPOSE, CONJ_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE)[(i // 4) % 4]
            cases.append((cha, chb, chc, cp, "frac", 13, 9, 21, ta, tb, "cr"[i % 2], 2.0 + 0.2j, 1.2 + 0.5j))
            cases.append((cha, chb, chc, cp, "pow2", 10, 12, 16, tb, ta, "rc"[i % 2], 0.5 - 0.25j, 2.0 + 0.5j))
    for chs in (("s", "s", "d"), ("d", "s", "z"), ("c", "z", "s"), ("z", "d", "c")):
        cases.append((*chs, 2, "frac", 40, 30, 300, NO_TRANSPOSE, NO_TRANSPOSE, "c", 2.0, 1.2))      # k > KC
        cases.append((*chs, 0, "frac", 9, 7, 5, NO_TRANSPOSE, NO_TRANSPOSE, "c", 2.0 + 0.2j, 0.0))    # beta == 0
        cases.append((*chs, 0, "frac", 6, 5, 4, NO_TRANSPOSE, NO_TRANSPOSE, "c", 0.0, 1.2 + 0.5j))    # alpha == 0
    return cases


def gemm_md_inputs(case, idx):
    cha, chb, chc, cp, kind, m, n, k, ta, tb, oc, al, be = case
    am, ak = (k, m) if ta & TRANSPOSE else (m, k)
    bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
    a = gen.matrix(cha, am, ak, 29 * idx + 1, kind, "c", pad=1)
    b = gen.matrix(chb, bk, bn, 29 * idx + 2, kind, "r", pad=2)
    c = gen.matrix(chc, m, n, 29 * idx + 3, kind, oc, pad=1)
    return a, b, c


def md_tol(case):
    """Elementwise tolerance: that of the lowest precision involved."""
    from util import TOL
    return TOL["s"] if (case[3] == 0 or any(ch in "sc" for ch in case[:3])) else TOL["d"]


def gemm_inputs(case, idx):
    ch, kind, m, n, k, ta, tb, oa, ob, oc, al, be = case
    am, ak = (k, m) if ta & TRANSPOSE else (m, k)
    bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
    a = gen.matrix(ch, am, ak, 11 * idx + 1, kind, oa, pad=3)
    b = gen.matrix(ch, bk, bn, 11 * idx + 2, kind, ob, pad=1)
    c = gen.matrix(ch, m, n, 11 * idx + 3, kind, oc, pad=2)
    return a, b, c


def trsm_inputs(case, idx):
    ch, kind, m, n, side, uplo, tr, dg, oa, ob, al = case
    ma = m if side == LEFT else n
    a = gen.triangular(ch, ma, 13 * idx + 1, kind, oa)
    gen.poison_unstored(a, uplo == LOWER)
    b = gen.matrix(ch, m, n, 13 * idx + 2, "ints" if kind == "ints" else kind, ob, pad=1)
    return a, b


# ----------------------------------------------------------------------------- reference enum values
def ref_enums():
    from build_ref import _inc_dirs  # type: ignore
    src = r'''
#include <stdio.h>
#include "blis.h"
int main(void){ printf("{\"PACKED_PANELS\":%d,\"TRIANGULAR\":%d,\"GENERAL\":%d,\"UNIT\":%d,\"NONUNIT\":%d,\"CONJ\":%d}",
  (int)BLIS_PACKED_PANELS,(int)BLIS_TRIANGULAR,(int)BLIS_GENERAL,(int)BLIS_UNIT_DIAG,(int)BLIS_NONUNIT_DIAG,(int)BLIS_CONJUGATE); return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        c = Path(td) / "e.c"
        c.write_text(src)
        exe = Path(td) / "e"
        subprocess.run(["gcc", "-std=c99", "-D_POSIX_C_SOURCE=200112L", *[f"-I{d}" for d in _inc_dirs()], str(c),
                        str(ROOT / "oracle" / "_ref" / "libblis_ref.so"), "-lm", "-lpthread", "-o", str(exe)], check=True)
        return json.loads(subprocess.run([str(exe)], capture_output=True, text=True, check=True,
                                         env={"LD_LIBRARY_PATH": str(ROOT / "oracle" / "_ref")}).stdout)


def write_gemmt(ref):
    res = {}
    for idx, cs in enumerate(gemmt_cases()):
        a, b, c = gemmt_inputs(cs, idx)
        gemmt_run(ref, cs, a, b, c)
        m = cs[3]
        unstored = np.triu(np.ones((m, m), bool), 1) if cs[5] == LOWER else np.tril(np.ones((m, m), bool), -1)
        assert np.isnan(np.abs(c[unstored])).all() and np.isfinite(np.abs(c[~unstored])).all(), \
            ("reference touched the unstored triangle of C?", cs)
        res[f"c{idx}"] = np.ascontiguousarray(c)
    np.savez_compressed(HERE / "gemmt.npz", **res)
    return len(res)


def write_strucmm(ref):
    res = {}
    for idx, cs in enumerate(strucmm_cases()):
        a, b, c = strucmm_inputs(cs, idx)
        out = strucmm_run(ref, cs, a, b, c)
        assert np.isfinite(np.abs(out)).all(), ("reference read the unstored triangle of A?", cs)
        res[f"c{idx}"] = np.ascontiguousarray(out)
    np.savez_compressed(HERE / "strucmm.npz", **res)
    return len(res)


def write_gemm_md():
    from refblis import ref_gemm_md
    res = {}
    for idx, cs in enumerate(gemm_md_cases()):
        a, b, c = gemm_md_inputs(cs, idx)
        ref_gemm_md(cs[8], cs[9], cs[11], a, b, cs[12], c, cs[3])
        res[f"c{idx}"] = np.ascontiguousarray(c)
    np.savez_compressed(HERE / "gemm_md.npz", **res)
    return len(res)


def main():
    ref = RefBlis(threads=1)
    L = ref.lib
    print("reference sub-configuration:", ref.arch())
    if len(sys.argv) > 1 and sys.argv[1] in ("gemmt", "strucmm", "gemm_md"):   # add one family's fixtures without rewriting the others
        n = {"gemmt": lambda: write_gemmt(ref), "strucmm": lambda: write_strucmm(ref), "gemm_md": write_gemm_md}[sys.argv[1]]()
        man = json.loads((HERE / "MANIFEST.json").read_text()); man["n_" + sys.argv[1]] = n
        (HERE / "MANIFEST.json").write_text(json.dumps(man, indent=1))
        print(sys.argv[1] + ".npz written:", n, "cases")
        return

    # ---- index arithmetic
    db, tr, p2 = index_cases()
    out = {"determine_blocksize": [], "thread_range_sub": [], "thread_partition_2x2": []}
    for (bw, dim, b_alg, b_max) in db:
        seq, i = [], 0
        while i < dim:
            b = ref.determine_blocksize(bw, i, dim, b_alg, b_max)
            seq.append(b)
            i += b
        out["determine_blocksize"].append([bw, dim, b_alg, b_max, seq])
    for (n_way, n, bf, low) in tr:
        out["thread_range_sub"].append([n_way, n, bf, low, [list(ref.thread_range_sub(w, n_way, n, bf, low)) for w in range(n_way)]])
    for (nt, w1, w2) in p2:
        out["thread_partition_2x2"].append([nt, w1, w2, list(ref.thread_partition_2x2(nt, w1, w2))])
    (HERE / "index_arith.json").write_text(json.dumps(out, separators=(",", ":")))

    # ---- packm micropanels
    en = ref_enums()
    cntx = L.bli_gks_query_cntx()
    packed = {}
    for idx, cs in enumerate(packm_cases()):
        ch, tri, uplo, unit, conj, inv, pd, pl, pdm, plm, pdo, plo, kappa, order = cs
        fn = getattr(L, f"bli_{ch}{ch}packm_struc_cxk")
        fn.argtypes = [ci, ci, ci, ci, ci, C.c_bool, i64, i64, i64, i64, i64, i64, i64, vp, vp, i64, i64, vp, i64, vp, vp]
        fn.restype = None
        src = gen.matrix(ch, pd, pl, 17 * idx + 5, "frac", order, pad=2)
        kap = np.array([kappa], dtype=src.dtype)
        p = np.full(pdm * plm, 777.0, dtype=src.dtype)
        rs, cs_ = src.strides[0] // src.itemsize, src.strides[1] // src.itemsize
        fn(en["TRIANGULAR"] if tri else en["GENERAL"], en["UNIT"] if unit else en["NONUNIT"], uplo,
           en["CONJ"] if conj else 0, en["PACKED_PANELS"], bool(inv), pd, pl, pdm, plm, pdo, plo, 1,
           kap.ctypes.data, src.ctypes.data, rs, cs_, p.ctypes.data, pdm, None, cntx)
        packed[f"p{idx}"] = p
    np.savez_compressed(HERE / "packm.npz", **packed)

    # ---- gemm / trsm
    res = {}
    for idx, cs in enumerate(gemm_cases()):
        a, b, c = gemm_inputs(cs, idx)
        ref.gemm(cs[5], cs[6], cs[10], a, b, cs[11], c)
        res[f"c{idx}"] = np.ascontiguousarray(c)
    np.savez_compressed(HERE / "gemm.npz", **res)
    res = {}
    for idx, cs in enumerate(trsm_cases()):
        a, b = trsm_inputs(cs, idx)
        ref.trsm(cs[4], cs[5], cs[6], cs[7], cs[10], a, b)
        assert np.isfinite(np.abs(b)).all(), ("reference read the unstored triangle?", cs)
        res[f"x{idx}"] = np.ascontiguousarray(b)
    np.savez_compressed(HERE / "trsm.npz", **res)
    (HERE / "MANIFEST.json").write_text(json.dumps({
        "generated_by": "tests/golden/make_golden.py", "reference_version": "3.0-dev (so 4.0.0)",
        "sub_configuration": ref.arch(), "threads": 1,
        "n_packm": len(packm_cases()), "n_gemm": len(gemm_cases()), "n_trsm": len(trsm_cases()),
        "n_gemmt": write_gemmt(ref), "n_strucmm": write_strucmm(ref),
        "n_gemm_md": write_gemm_md()}, indent=1))
    print("golden fixtures written:", sorted(p.name for p in HERE.iterdir()))


if __name__ == "__main__":
    main()
