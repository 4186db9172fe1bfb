"""Parity of the CUDA mixed-datatype gemm (b200_gemm_md through the C ABI) with the reference.

SURVEY.md section 8f rank 3.  All 4 x 4 x 4 storage-datatype combinations of A, B, C and both computation precisions
(docs/MixedDatatypes.md).  Checkers: golden fixtures produced by the real reference's object API (tests/ref_shim.c), the
oracle restatement, the live reference.  Bars: bit-exact on power-of-two inputs; otherwise the elementwise tolerance of
the lowest precision involved (util.TOL)."""
import itertools

import numpy as np
import pytest
import torch

import gen
import make_golden as G
from refblis import CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE, NO_TRANSPOSE, TRANSPOSE, oracle_gemm_md, ref_gemm_md
from util import estr, rel_err, to_numpy, to_torch

pytestmark = pytest.mark.gpu
GOLD = G.HERE
T32, T64 = torch.float32, torch.float64


def run_md(engine, case, a, b, c, device="cuda"):
    cha, chb, chc, cp, kind, m, n, k, ta, tb, oc, al, be = case
    from blis_b200 import api
    ao, bo, co = api.Obj(to_torch(a, device)), api.Obj(to_torch(b, device)), api.Obj(to_torch(c, device))
    api.bli_obj_set_conjtrans(ta, ao); api.bli_obj_set_conjtrans(tb, bo)
    api.bli_gemm_md(al, ao, bo, be, co, comp_prec=T32 if cp == 0 else T64)
    if device == "cuda":
        torch.cuda.synchronize()
    return to_numpy(co.buf)


def test_gemm_md_golden_fixtures(engine):
    gold = np.load(GOLD / "gemm_md.npz")
    for idx, cs in enumerate(G.gemm_md_cases()):
        a, b, c = G.gemm_md_inputs(cs, idx)
        got = run_md(engine, cs, a, b, c)
        want = gold[f"c{idx}"]
        if cs[4] == "pow2":
            assert float(np.abs(got - want).max()) < 1e-300, f"gemm_md golden case {idx} {cs}: not bit-exact"
        else:
            assert rel_err(got, want) <= G.md_tol(cs), f"gemm_md golden case {idx} {cs}: {rel_err(got, want)}"


@pytest.mark.parametrize("chc", list("sdcz"))
def test_gemm_md_vs_oracle_shapes(engine, oracle, chc):
    """Ragged multi-tile shapes, every A/B datatype, both computation precisions, all transpositions."""
    trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE)
    idx = 3000
    for cha, chb in itertools.product("sdcz", repeat=2):
        for cp in (0, 2):
            for (m, n, k) in ((129, 67, 130), (1, 300, 17), (260, 5, 513)):
                idx += 1
                ta, tb = trs[idx % 4], trs[(idx // 4) % 4]
                cs = (cha, chb, chc, cp, "frac", m, n, k, ta, tb, "crg"[idx % 3], 2.0 + 0.2j, 1.2 + 0.5j)
                a, b, c = G.gemm_md_inputs(cs, idx)
                want = c.copy(order="K")
                oracle_gemm_md(oracle, ta, tb, cs[11], a, b, cs[12], want, cp)
                got = run_md(engine, cs, a, b, c)
                assert rel_err(got, want) <= G.md_tol(cs), (cs, rel_err(got, want))


def test_gemm_md_pow2_bit_exact_and_host_operands(engine, oracle):
    idx = 6000
    for cha, chb, chc in itertools.product("sdcz", repeat=3):
        idx += 1
        cp = 2 * (idx % 2)
        cs = (cha, chb, chc, cp, "pow2", 140, 70, 64, (0, 8, 16, 24)[idx % 4], (0, 24, 8, 16)[(idx // 2) % 4], "cr"[idx % 2], 0.5 - 0.25j, 2.0 + 0.5j)
        a, b, c = G.gemm_md_inputs(cs, idx)
        want = c.copy(order="K")
        oracle_gemm_md(oracle, cs[8], cs[9], cs[11], a, b, cs[12], want, cp)
        got = run_md(engine, cs, a, b, c, device="cuda" if idx % 3 else "cpu")
        assert np.array_equal(got, want), (cs, "not bit-exact")


def test_gemm_md_vs_real_reference_and_specials(engine, ref):
    """Testsuite-sized mixed problems against the live reference; alpha == 0 and beta == 0 semantics."""
    from blis_b200 import api
    idx = 9000
    for (cha, chb, chc, cp) in (("s", "s", "d", 2), ("d", "d", "s", 2), ("c", "d", "z", 0), ("z", "c", "d", 2), ("d", "z", "c", 2), ("s", "z", "z", 0)):
        idx += 1
        cs = (cha, chb, chc, cp, "frac", 600, 500, 700, NO_TRANSPOSE, TRANSPOSE, "c", 2.0 + 0.2j, 1.2 + 0.5j)
        a, b, c = G.gemm_md_inputs(cs, idx)
        want = c.copy(order="K")
        ref_gemm_md(cs[8], cs[9], cs[11], a, b, cs[12], want, cp)
        got = run_md(engine, cs, a, b, c)
        assert rel_err(got, want) <= G.md_tol(cs) * 4, (cs, rel_err(got, want))
    # beta == 0 must not read C, alpha == 0 must not read A*B
    a = torch.full((50, 40), float("nan"), dtype=torch.float32, device="cuda"); b = torch.ones(40, 30, dtype=torch.complex128, device="cuda")
    c = torch.full((50, 30), 2.0, dtype=torch.complex64, device="cuda")
    api.bli_gemm_md(0.0, api.Obj(a), api.Obj(b), 0.5, api.Obj(c)); torch.cuda.synchronize()
    assert bool((c == 1.0).all())
    a.fill_(1.0); c.fill_(float("nan"))
    api.bli_gemm_md(1.0, api.Obj(a), api.Obj(b), 0.0, api.Obj(c)); torch.cuda.synchronize()
    assert bool((c == 40.0).all())
    # the object API dispatches mixed operands by itself
    c64 = torch.zeros(50, 30, dtype=torch.float64, device="cuda")
    api.bli_gemm(1.0, api.Obj(a), api.Obj(b), 0.0, api.Obj(c64)); torch.cuda.synchronize()
    assert bool((c64 == 40.0).all())
