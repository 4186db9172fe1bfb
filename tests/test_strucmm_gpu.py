"""Parity of the CUDA hemm / symm / trmm3 / trmm (through the C ABI) with the reference.

SURVEY.md section 8f rank 2.  Checkers: golden fixtures produced by the real reference (tests/golden/strucmm.npz), the
oracle restatement, the real reference library when it travelled.  Bars: bit-exact on power-of-two inputs, elementwise
util.TOL otherwise.  The triangle of A that is not stored is NaN in every case, so a finite result also proves it is
never read (nor the diagonal when it is declared unit).
"""
import numpy as np
import pytest
import torch

import gen
import make_golden as G
from refblis import (CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE, LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, RIGHT, TRANSPOSE, UNIT_DIAG,
                     UPPER)
from util import NP2T, TOL, estr, rel_err, to_numpy, to_torch

pytestmark = pytest.mark.gpu
GOLD = G.HERE
OPS = ("hemm", "symm", "trmm3", "trmm")


def run_op(engine, case, a, b, c, device="cuda", pin=False):
    """One case (tuple layout of make_golden.strucmm_cases) on numpy inputs; returns the result (C, or B for trmm)."""
    ch, op, kind, m, n, side, uplo, ta, dg, tb, oa, ob, oc, al, be = case
    ta_ = to_torch(a, device, pin=pin)
    tb_ = to_torch(b, device, pin=pin)
    if op == "trmm":
        getattr(engine, f"bli_{ch}trmm")(side, uplo, ta, dg, m, n, al, ta_, *estr(a), tb_, *estr(b))
        out = tb_
    else:
        tc_ = to_torch(c, device, pin=pin)
        if op == "trmm3":
            getattr(engine, f"bli_{ch}trmm3")(side, uplo, ta, dg, tb, m, n, al, ta_, *estr(a), tb_, *estr(b), be, tc_, *estr(c))
        else:
            getattr(engine, f"bli_{ch}{op}")(side, uplo, ta & CONJ_NO_TRANSPOSE, tb, m, n, al, ta_, *estr(a), tb_, *estr(b), be, tc_, *estr(c))
        out = tc_
    if device == "cuda":
        torch.cuda.synchronize()
    return to_numpy(out)


def test_strucmm_golden_fixtures(engine):
    gold = np.load(GOLD / "strucmm.npz")
    for idx, cs in enumerate(G.strucmm_cases()):
        a, b, c = G.strucmm_inputs(cs, idx)
        got = run_op(engine, cs, a, b, c)
        want = gold[f"c{idx}"]
        if cs[2] == "pow2":
            assert np.array_equal(got, want), f"strucmm golden case {idx} {cs}: not bit-exact"
        else:
            assert rel_err(got, want) <= TOL[cs[0]], f"strucmm golden case {idx} {cs}: {rel_err(got, want)}"


SHAPES = [(1, 1), (2, 5), (127, 65), (128, 16), (129, 300), (257, 100), (300, 513), (515, 64), (64, 700)]


@pytest.mark.parametrize("ch", list("sdcz"))
def test_strucmm_vs_oracle_all_params(engine, oracle, ch):
    """Every op x side x uplo x trans/conj x diag combination over edge shapes and storage layouts."""
    cx = ch in "cz"
    trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
    idx = 1000
    for (m, n) in SHAPES:
        for op in OPS:
            tas = trs if op in ("trmm3", "trmm") else ((NO_TRANSPOSE, CONJ_NO_TRANSPOSE) if cx else (NO_TRANSPOSE,))
            for side in (LEFT, RIGHT):
                for uplo in (LOWER, UPPER):
                    for ti, ta in enumerate(tas):
                        tb = trs[(ti + m + n) % len(trs)]
                        dg = UNIT_DIAG if (ti + m) % 2 else NONUNIT_DIAG
                        layouts = (("c", "c", "c"), ("r", "r", "r"), ("g", "c", "r")) if m * n <= 40000 else (("c", "c", "c"),)
                        for (oa, ob, oc) in layouts:
                            idx += 1
                            cs = (ch, op, "frac", m, n, side, uplo, ta, dg, tb, oa, ob, oc, al, be)
                            a, b, c = G.strucmm_inputs(cs, idx)
                            if dg == UNIT_DIAG and op in ("trmm3", "trmm"):
                                a[np.diag_indices(a.shape[0])] = np.nan           # a unit diagonal is never read
                            wa, wb, wc = a.copy(order="K"), b.copy(order="K"), (c.copy(order="K") if c is not None else None)
                            want = G.strucmm_run(oracle, cs, wa, wb, wc)
                            got = run_op(engine, cs, a, b, c)
                            err = rel_err(got, want)
                            assert err <= TOL[ch], (cs, err)


@pytest.mark.parametrize("ch", list("sdcz"))
def test_strucmm_pow2_bit_exact_vs_oracle(engine, oracle, ch):
    idx = 5000
    for op in OPS:
        for (m, n, side, uplo, ta, dg, tb, oc) in ((257, 64, LEFT, LOWER, 0, NONUNIT_DIAG, 0, "c"), (130, 148, RIGHT, UPPER, 8, UNIT_DIAG, 0, "r"),
                                                    (384, 33, LEFT, UPPER, 8, NONUNIT_DIAG, 8, "c"), (65, 250, RIGHT, LOWER, 0, UNIT_DIAG, 8, "c")):
            idx += 1
            cs = (ch, op, "pow2", m, n, side, uplo, ta, dg, tb, "c", "c", oc, 2.0, 0.5)
            a, b, c = G.strucmm_inputs(cs, idx)
            wa, wb, wc = a.copy(order="K"), b.copy(order="K"), (c.copy(order="K") if c is not None else None)
            want = G.strucmm_run(oracle, cs, wa, wb, wc)
            got = run_op(engine, cs, a, b, c)
            assert np.array_equal(got, want), (cs, "not bit-exact")


@pytest.mark.parametrize("ch", list("sdcz"))
def test_strucmm_vs_real_reference(engine, ref, ch):
    """Against the real reference BLIS at the testsuite's size 1000."""
    cx = ch in "cz"
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
    idx = 7000
    for op in OPS:
        for (m, n, side, uplo, ta, tb) in ((1000, 1000, LEFT, LOWER, NO_TRANSPOSE, NO_TRANSPOSE),
                                           (769, 300, RIGHT, UPPER, CONJ_TRANSPOSE if cx else TRANSPOSE, TRANSPOSE)):
            idx += 1
            cs = (ch, op, "frac", m, n, side, uplo, ta, NONUNIT_DIAG, tb, "c", "c", "c", al, be)
            a, b, c = G.strucmm_inputs(cs, idx)
            wa, wb, wc = a.copy(order="K"), b.copy(order="K"), (c.copy(order="K") if c is not None else None)
            want = G.strucmm_run(ref, cs, wa, wb, wc)
            got = run_op(engine, cs, a, b, c)
            assert rel_err(got, want) <= TOL[ch] * 4, (cs, rel_err(got, want))


@pytest.mark.parametrize("ch", list("sdcz"))
def test_strucmm_host_operands(engine, oracle, ch):
    """Pageable and pinned host operands (trmm: B comes back in the caller's buffer)."""
    cx = ch in "cz"
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
    idx = 8000
    for op in OPS:
        for (oa, ob, oc, pin, alpha, side) in (("c", "c", "c", False, al, LEFT), ("r", "c", "g", False, al, RIGHT), ("c", "r", "r", True, al, LEFT),
                                               ("c", "c", "c", False, 0.0, RIGHT)):
            idx += 1
            cs = (ch, op, "frac", 211, 150, side, LOWER if idx % 2 else UPPER, NO_TRANSPOSE, NONUNIT_DIAG, TRANSPOSE, oa, ob, oc, alpha, be)
            a, b, c = G.strucmm_inputs(cs, idx)
            wa, wb, wc = a.copy(order="K"), b.copy(order="K"), (c.copy(order="K") if c is not None else None)
            want = G.strucmm_run(oracle, cs, wa, wb, wc)
            got = run_op(engine, cs, a, b, c, device="cpu", pin=pin)
            assert rel_err(got, want) <= TOL[ch], (cs, rel_err(got, want))


@pytest.mark.parametrize("ch,m,n", [("d", 8192, 2048), ("s", 8192, 2048), ("z", 4096, 1024), ("c", 4096, 1024)])
def test_trmm_large_k_range_skipping_is_exact(engine, ch, m, n):
    """At sizes no CPU checker reaches in seconds: trmm3 with the zero k range skipped must equal, bit for bit, the
    same call with the skipping switched off (the skipped products are exact zeros), and both must equal the engine's
    gemm on an explicitly zero-filled triangular matrix; trmm (in place) must equal trmm3 with beta = 0; symm must equal
    gemm on the mirrored matrix."""
    dt = NP2T[np.dtype(gen.NP_DT[ch])]
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(int(0xB200))
    rdt = torch.float32 if ch in "sc" else torch.float64

    def rnd(r, c_):
        x = torch.rand(c_, r, dtype=rdt, device=dev, generator=g) * 2 - 1
        if ch in "cz":
            x = torch.complex(x, torch.rand(c_, r, dtype=rdt, device=dev, generator=g) * 2 - 1)
        return (x / 32).to(dt).t()                       # column-major r x c_

    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if ch in "cz" else (2.0, 1.2))
    gemm = getattr(engine, f"bli_{ch}gemm")
    trmm3 = getattr(engine, f"bli_{ch}trmm3")
    for side, uplo in ((LEFT, LOWER), (RIGHT, UPPER), (LEFT, UPPER)):
        ma = m if side == LEFT else n
        a, b, c0 = rnd(ma, ma), rnd(m, n), rnd(m, n)
        a_tri = (torch.tril(a) if uplo == LOWER else torch.triu(a)).t().contiguous().t()      # explicit zeros, column-major
        ones = torch.ones(ma, ma, dtype=torch.bool, device=dev)
        a_nan = torch.where(torch.tril(ones) if uplo == LOWER else torch.triu(ones), a, torch.full_like(a, float("nan")))
        a_nan = a_nan.t().contiguous().t()                 # unstored triangle NaN (by index: a stored value may be exactly 0)
        full = c0.clone()
        if side == LEFT:
            gemm(0, 0, m, n, m, al, a_tri, 1, ma, b, 1, m, be, full, 1, m)
        else:
            gemm(0, 0, m, n, n, al, b, 1, m, a_tri, 1, ma, be, full, 1, m)
        c1 = c0.clone()
        trmm3(side, uplo, 0, NONUNIT_DIAG, 0, m, n, al, a_nan, 1, ma, b, 1, m, be, c1, 1, m)
        engine.set_option("ktri_skip", 0)
        try:
            c2 = c0.clone()
            trmm3(side, uplo, 0, NONUNIT_DIAG, 0, m, n, al, a_nan, 1, ma, b, 1, m, be, c2, 1, m)
        finally:
            engine.set_option("ktri_skip", 1)
        torch.cuda.synchronize()
        assert torch.equal(c1, c2), (ch, side, uplo, "k-range skipping changed the result")
        assert torch.equal(c1, full), (ch, side, uplo, "trmm3 differs from gemm on the zero-filled matrix")
        b2 = b.clone()
        getattr(engine, f"bli_{ch}trmm")(side, uplo, 0, NONUNIT_DIAG, m, n, al, a_nan, 1, ma, b2, 1, m)
        c3 = torch.full_like(c0, float("nan"))
        trmm3(side, uplo, 0, NONUNIT_DIAG, 0, m, n, al, a_nan, 1, ma, b, 1, m, 0.0, c3, 1, m)
        torch.cuda.synchronize()
        assert torch.equal(b2, c3), (ch, side, uplo, "trmm differs from trmm3 with beta = 0")
    # symm == gemm on the mirrored matrix
    a, b, c0 = rnd(m, m), rnd(m, n), rnd(m, n)
    a_sym = (torch.tril(a) + torch.tril(a, -1).t()).t().contiguous().t()
    full = c0.clone(); gemm(0, 0, m, n, m, al, a_sym, 1, m, b, 1, m, be, full, 1, m)
    a_low = torch.where(torch.tril(torch.ones(m, m, dtype=torch.bool, device=dev)), a, torch.full_like(a, float("nan"))).t().contiguous().t()
    c1 = c0.clone(); getattr(engine, f"bli_{ch}symm")(LEFT, LOWER, 0, 0, m, n, al, a_low, 1, m, b, 1, m, be, c1, 1, m)
    torch.cuda.synchronize()
    assert torch.equal(c1, full), (ch, "symm differs from gemm on the mirrored matrix")
