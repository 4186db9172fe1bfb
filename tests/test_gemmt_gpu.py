"""Parity of the CUDA gemmt family (gemmt, syrk, herk, syr2k, her2k through the C ABI) with the reference.

SURVEY.md section 8f rank 1.  Checkers: golden fixtures produced by the real reference
(tests/golden/gemmt.npz), the oracle restatement, and the real reference library when it travelled.
Bars: bit-exact on power-of-two inputs, elementwise util.TOL otherwise; the triangle of C that is not stored
must come back bit-for-bit untouched (it is NaN-poisoned in most cases so that any stray read shows up too);
at large sizes the stored triangle must equal the engine's own full gemm bit for bit.
"""
import numpy as np
import pytest
import torch

import gen
import make_golden as G
from refblis import CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE, LOWER, NO_TRANSPOSE, TRANSPOSE, UPPER
from util import NP2T, TOL, estr, rel_err, to_numpy, to_torch

pytestmark = pytest.mark.gpu
GOLD = G.HERE
OPS = ("gemmt", "syrk", "herk", "syr2k", "her2k")


def run_op(engine, case, a, b, c, device="cuda", pin=False):
    """Run one gemmt-family case (tuple layout of make_golden.gemmt_cases) on numpy inputs; returns C as numpy."""
    ch, op, kind, m, k, uplo, ta, tb, oa, ob, oc, al, be = case
    ta_ = to_torch(a, device, pin=pin)
    tb_ = to_torch(b, device, pin=pin) if b is not None else None
    tc_ = to_torch(c, device, pin=pin)
    fn = getattr(engine, f"bli_{ch}{op}")
    if op in ("syrk", "herk"):
        fn(uplo, ta, m, k, al, ta_, *estr(a), be, tc_, *estr(c))
    else:
        fn(uplo, ta, tb, m, k, al, ta_, *estr(a), tb_, *estr(b), be, tc_, *estr(c))
    if device == "cuda":
        torch.cuda.synchronize()
    return to_numpy(tc_)


def masks(m, uplo):
    unstored = np.triu(np.ones((m, m), bool), 1) if uplo == LOWER else np.tril(np.ones((m, m), bool), -1)
    return ~unstored, unstored


def check(case, got, want, c_in, exact=False, tol_mult=1.0):
    ch, op, m, uplo = case[0], case[1], case[3], case[5]
    stored, unstored = masks(m, uplo)
    assert np.array_equal(got[unstored], c_in[unstored], equal_nan=True), (case, "unstored triangle of C was written")
    if exact:
        assert np.array_equal(got[stored], want[stored]), (case, "not bit-exact")
    else:
        err = rel_err(got[stored], want[stored])
        assert err <= TOL[ch] * tol_mult, (case, err)
    if op in ("herk", "her2k") and ch in "cz":
        assert (np.diag(got).imag == 0).all(), (case, "imaginary diagonal not zeroed")


def test_gemmt_family_golden_fixtures(engine):
    gold = np.load(GOLD / "gemmt.npz")
    for idx, cs in enumerate(G.gemmt_cases()):
        a, b, c = G.gemmt_inputs(cs, idx)
        got = run_op(engine, cs, a, b, c)
        check(cs, got, gold[f"c{idx}"], c, exact=(cs[2] == "pow2"))


SHAPES = [(1, 1), (2, 5), (8, 4), (127, 65), (128, 16), (129, 17), (130, 300), (257, 100), (300, 513), (515, 64), (700, 33),
          (64, 1), (1000, 8)]


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemmt_family_vs_oracle_all_params(engine, oracle, ch):
    """Every op x uplo x trans/conj combination x storage combination over edge shapes (ragged tiles, diagonal
    crossing tile corners, m < tile, k == 1)."""
    cx = ch in "cz"
    trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
    idx = 1000
    for (m, k) in SHAPES:
        for op in OPS:
            a_, b_ = (2.0 if op == "herk" else al), (1.2 if op in ("herk", "her2k") else be)
            for uplo in (LOWER, UPPER):
                for ti, ta in enumerate(trs):
                    tb = trs[(ti + (m + k) % len(trs)) % len(trs)]
                    for (oa, ob, oc) in (("c", "c", "c"), ("r", "r", "r"), ("c", "r", "g"), ("g", "c", "r")):
                        if m * m * k > 1_500_000 and (oa, ob, oc) not in (("c", "c", "c"), ("r", "r", "r")):
                            continue
                        idx += 1
                        cs = (ch, op, "frac", m, k, uplo, ta, tb, oa, ob, oc, a_, b_)
                        a, b, c = G.gemmt_inputs(cs, idx)
                        want = c.copy(order="K")
                        G.gemmt_run(oracle, cs, a, b, want)
                        got = run_op(engine, cs, a, b, c)
                        check(cs, got, want, c)


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemmt_family_pow2_bit_exact_vs_oracle(engine, oracle, ch):
    """Power-of-two inputs: every product and partial sum is exact, so the GPU tile order and the KC-blocked
    order of the reference must agree in every bit of the stored triangle."""
    idx = 5000
    for op in OPS:
        for (m, k, uplo, ta, tb, oc) in ((257, 64, LOWER, 0, 0, "c"), (130, 48, UPPER, 8, 0, "r"), (384, 33, LOWER, 0, 8, "c"),
                                         (65, 50, UPPER, 8, 8, "c")):
            idx += 1
            cs = (ch, op, "pow2", m, k, uplo, ta, tb, "c", "c", oc, 2.0, 0.5)
            a, b, c = G.gemmt_inputs(cs, idx)
            want = c.copy(order="K")
            G.gemmt_run(oracle, cs, a, b, want)
            got = run_op(engine, cs, a, b, c)
            check(cs, got, want, c, exact=True)


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemmt_family_vs_real_reference(engine, ref, ch):
    """Against the real reference BLIS (optimized CPU kernels, multithreaded) at the testsuite's size 1000."""
    cx = ch in "cz"
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
    idx = 7000
    for op in OPS:
        a_, b_ = (2.0 if op == "herk" else al), (1.2 if op in ("herk", "her2k") else be)
        for (m, k, uplo, ta, tb) in ((1000, 1000, LOWER, NO_TRANSPOSE, NO_TRANSPOSE), (769, 300, UPPER, TRANSPOSE, CONJ_TRANSPOSE if cx else TRANSPOSE)):
            idx += 1
            cs = (ch, op, "frac", m, k, uplo, ta, tb, "c", "c", "c", a_, b_)
            a, b, c = G.gemmt_inputs(cs, idx)
            want = c.copy(order="K")
            G.gemmt_run(ref, cs, a, b, want)
            got = run_op(engine, cs, a, b, c)
            check(cs, got, want, c, tol_mult=4.0)


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemmt_family_host_operands(engine, oracle, ch):
    """Pageable and pinned host operands: staged in, and only... the whole array of C travels back, so the triangle
    that is not stored must return with its original bits."""
    cx = ch in "cz"
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
    idx = 8000
    for op in OPS:
        a_, b_ = (2.0 if op == "herk" else al), (1.2 if op in ("herk", "her2k") else be)
        for (oa, ob, oc, pin, beta) in (("c", "c", "c", False, b_), ("r", "c", "g", False, b_), ("c", "r", "r", True, b_), ("c", "c", "c", False, 0.0)):
            idx += 1
            cs = (ch, op, "frac", 211, 150, LOWER if idx % 2 else UPPER, NO_TRANSPOSE, TRANSPOSE, oa, ob, oc, a_, beta)
            a, b, c = G.gemmt_inputs(cs, idx)
            want = c.copy(order="K")
            G.gemmt_run(oracle, cs, a, b, want)
            got = run_op(engine, cs, a, b, c, device="cpu", pin=pin)
            check(cs, got, want, c)


def test_gemmt_trivial_and_special_cases(engine):
    """m == 0 is a no-op; k == 0 or alpha == 0 scale the stored triangle only; beta == 0 must not read C
    (NaN in the stored triangle disappears); bad uplo fails loudly."""
    from blis_b200._lib import EngineError
    dev = "cuda"
    for dt, ch in ((torch.float32, "s"), (torch.float64, "d"), (torch.complex64, "c"), (torch.complex128, "z")):
        m, k = 150, 40
        a = torch.ones(m, k, dtype=dt, device=dev)
        c = torch.full((m, m), 3.0, dtype=dt, device=dev)
        syrk = getattr(engine, f"bli_{ch}syrk")
        syrk(LOWER, 0, 0, k, 1.0, a, k, 1, 1.0, c, m, 1); torch.cuda.synchronize()
        assert bool((c == 3.0).all())
        syrk(LOWER, 0, m, 0, 1.0, a, k, 1, 0.5, c, m, 1); torch.cuda.synchronize()          # k == 0: tril(C) *= 0.5
        assert bool((torch.tril(c) == torch.tril(torch.full_like(c, 1.5))).all()) and bool((torch.triu(c, 1) == torch.triu(torch.full_like(c, 3.0), 1)).all())
        syrk(UPPER, 0, m, k, 0.0, a, k, 1, 2.0, c, m, 1); torch.cuda.synchronize()          # alpha == 0: triu(C) *= 2
        assert bool((torch.triu(c, 1) == torch.triu(torch.full_like(c, 6.0), 1)).all()) and bool((torch.diagonal(c) == 3.0).all())
        c.fill_(float("nan"))
        syrk(LOWER, 0, m, k, 1.0, a, k, 1, 0.0, c, m, 1); torch.cuda.synchronize()          # beta == 0: C never read
        assert bool((torch.tril(c) == torch.tril(torch.full_like(c, float(k)))).all())
        assert bool(torch.isnan(torch.triu(c, 1).abs()[torch.triu(torch.ones(m, m, dtype=torch.bool, device=dev), 1)]).all())
        with pytest.raises(EngineError):
            syrk(0xE0, 0, m, k, 1.0, a, k, 1, 0.0, c, m, 1)                                     # BLIS_DENSE is not a triangle


@pytest.mark.parametrize("ch,m,k", [("d", 8192, 1024), ("s", 8192, 1024), ("z", 4096, 512), ("c", 4096, 512)])
def test_gemmt_large_equals_gemm_on_stored_triangle(engine, ch, m, k):
    """At sizes no CPU checker reaches in seconds: syrk/gemmt on device-resident column-major operands must
    produce, on the stored triangle, exactly the bits of the engine's own gemm (same kernel, same k order),
    and must leave the other triangle untouched; syr2k == gemmt + gemmt by linearity of the accumulation."""
    dt = NP2T[np.dtype(gen.NP_DT[ch])]
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(int(0xB200))
    rdt = torch.float32 if ch in "sc" else torch.float64

    def rnd(r, c_):
        x = torch.rand(c_, r, dtype=rdt, device=dev, generator=g) * 2 - 1
        if ch in "cz":
            x = torch.complex(x, torch.rand(c_, r, dtype=rdt, device=dev, generator=g) * 2 - 1)
        return (x / 32).to(dt).t()                       # column-major r x c_

    a, b, c0 = rnd(m, k), rnd(k, m), rnd(m, m)
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if ch in "cz" else (2.0, 1.2))
    full = c0.clone()
    getattr(engine, f"bli_{ch}gemm")(0, 0, m, m, k, al, a, 1, m, b, 1, k, be, full, 1, m)
    for uplo in (LOWER, UPPER):
        c = c0.clone()
        getattr(engine, f"bli_{ch}gemmt")(uplo, 0, 0, m, k, al, a, 1, m, b, 1, k, be, c, 1, m)
        torch.cuda.synchronize()
        keep = torch.tril if uplo == LOWER else torch.triu
        drop = (lambda x: torch.triu(x, 1)) if uplo == LOWER else (lambda x: torch.tril(x, -1))
        assert torch.equal(keep(c), keep(full)), (ch, uplo, "stored triangle differs from gemm")
        assert torch.equal(drop(c), drop(c0)), (ch, uplo, "unstored triangle written")
    # syrk(A) == gemmt(A, A^T) bit for bit
    c1, c2 = c0.clone(), c0.clone()
    getattr(engine, f"bli_{ch}syrk")(LOWER, 0, m, k, al, a, 1, m, be, c1, 1, m)
    getattr(engine, f"bli_{ch}gemmt")(LOWER, 0, TRANSPOSE, m, k, al, a, 1, m, a, 1, m, be, c2, 1, m)
    torch.cuda.synchronize()
    assert torch.equal(c1, c2)
