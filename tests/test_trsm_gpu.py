"""Parity of the CUDA trsm (through the C ABI) with the reference: golden
fixtures (real reference outputs), the oracle, the real reference library, and
the testsuite's residual at BASELINE size (testsuite/src/test_trsm.c:362-381).
Integer-valued systems must be solved bit-exactly."""
import numpy as np
import pytest
import torch

import gen
import make_golden as G
from refblis import (CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE, LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, RIGHT, TRANSPOSE,
                     UNIT_DIAG, UPPER)
from util import NP2T, TOL, estr, rel_err, to_numpy, to_torch

pytestmark = pytest.mark.gpu
TRSM = {"s": "bli_strsm", "d": "bli_dtrsm", "c": "bli_ctrsm", "z": "bli_ztrsm"}


def run_trsm(engine, ch, side, uplo, tr, dg, alpha, a, b, device="cuda"):
    ta_, tb_ = to_torch(a, device), to_torch(b, device)
    m, n = b.shape
    getattr(engine, TRSM[ch])(side, uplo, tr, dg, m, n, alpha, ta_, *estr(a), tb_, *estr(b))
    if device == "cuda":
        torch.cuda.synchronize()
    return to_numpy(tb_)


def test_trsm_golden_fixtures(engine):
    gold = np.load(G.HERE / "trsm.npz")
    for idx, cs in enumerate(G.trsm_cases()):
        ch, kind = cs[0], cs[1]
        a, b = G.trsm_inputs(cs, idx)              # unstored triangle of A is NaN-poisoned
        got = run_trsm(engine, ch, cs[4], cs[5], cs[6], cs[7], cs[10], a, b)
        want = gold[f"x{idx}"]
        if kind == "ints":
            assert np.array_equal(got, want), f"trsm golden case {idx} {cs}: not bit-exact"
        else:
            assert rel_err(got, want) <= 20 * TOL[ch], f"trsm golden case {idx} {cs}: {rel_err(got, want)}"


@pytest.mark.parametrize("ch", list("sdcz"))
def test_trsm_vs_oracle_all_params(engine, oracle, ch):
    cx = ch in "cz"
    trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
    al = (2.0 + 0.3j) if cx else 2.0
    seed = 0
    for (m, n) in ((1, 1), (5, 9), (64, 64), (65, 63), (100, 37), (257, 130), (33, 700), (700, 33)):
        for side in (LEFT, RIGHT):
            for uplo in (LOWER, UPPER):
                for tr in trs:
                    for dg in (NONUNIT_DIAG, UNIT_DIAG):
                        for (oa, ob) in (("c", "c"), ("r", "r"), ("c", "g")):
                            if m * n > 20000 and (oa, ob) != ("c", "c"):
                                continue
                            seed += 1
                            ma = m if side == LEFT else n
                            a = gen.triangular(ch, ma, seed, "frac", oa)
                            gen.poison_unstored(a, uplo == LOWER)
                            b = gen.matrix(ch, m, n, seed + 7000, "frac", ob, pad=1)
                            want = b.copy(order="K")
                            oracle.trsm(side, uplo, tr, dg, al, a, want)
                            got = run_trsm(engine, ch, side, uplo, tr, dg, al, a, b)
                            assert rel_err(got, want) <= 20 * TOL[ch], (ch, m, n, side, uplo, tr, dg, oa, ob, rel_err(got, want))


@pytest.mark.parametrize("ch", list("sdcz"))
def test_trsm_integer_systems_bit_exact(engine, oracle, ch):
    """Unit-ish integer triangular systems have exact integer solutions: block
    recursion + gemm updates on the GPU must reproduce the reference's bits."""
    for idx, (m, n, side, uplo, tr) in enumerate(((20, 33, LEFT, LOWER, 0), (18, 70, LEFT, UPPER, 8),
                                                  (40, 16, RIGHT, LOWER, 0), (24, 24, RIGHT, UPPER, 8))):
        ma = m if side == LEFT else n
        a = gen.triangular(ch, ma, 500 + idx, "ints")
        # keep growth tame: strictly-triangular entries in {-1,0,1}
        a[...] = np.clip(a.real, -1, 1) if ch in "sd" else np.clip(a.real, -1, 1) + 0j
        d = np.arange(ma); a[d, d] = 1 - 2 * (d % 2)
        b = gen.matrix(ch, m, n, 600 + idx, "ints")
        want = b.copy(order="K"); oracle.trsm(side, uplo, tr, NONUNIT_DIAG, 2.0, a, want)
        got = run_trsm(engine, ch, side, uplo, tr, NONUNIT_DIAG, 2.0, a, b)
        assert np.isfinite(np.abs(want)).all()
        assert np.array_equal(got, want), (ch, m, n, side, uplo, tr)


@pytest.mark.parametrize("ch", list("sdcz"))
def test_trsm_vs_real_reference(engine, ref, ch):
    al = (2.0 + 0.3j) if ch in "cz" else 2.0
    for idx, (m, n, side, uplo, tr) in enumerate(((1000, 1000, LEFT, LOWER, 0), (700, 300, LEFT, UPPER, 0),
                                                  (300, 900, RIGHT, LOWER, 8), (513, 257, RIGHT, UPPER, 0))):
        ma = m if side == LEFT else n
        a = gen.triangular(ch, ma, 70 + idx, "frac"); b = gen.matrix(ch, m, n, 80 + idx, "frac")
        want = b.copy(order="K"); ref.trsm(side, uplo, tr, NONUNIT_DIAG, al, a, want)
        got = run_trsm(engine, ch, side, uplo, tr, NONUNIT_DIAG, al, a, b)
        assert rel_err(got, want) <= 50 * TOL[ch], (ch, m, n, side, uplo, tr, rel_err(got, want))


def test_trsm_special_cases(engine, oracle):
    dev = "cuda"
    a = torch.eye(6, dtype=torch.float64, device=dev) * 2
    b = torch.full((6, 4), 3.0, dtype=torch.float64, device=dev)
    engine.bli_dtrsm(LEFT, LOWER, 0, NONUNIT_DIAG, 0, 4, 1.0, a, 6, 1, b, 4, 1)          # m == 0
    engine.bli_dtrsm(LEFT, LOWER, 0, NONUNIT_DIAG, 6, 0, 1.0, a, 6, 1, b, 4, 1)          # n == 0
    torch.cuda.synchronize(); assert bool((b == 3.0).all())
    engine.bli_dtrsm(LEFT, LOWER, 0, NONUNIT_DIAG, 6, 4, 1.0, a, 6, 1, b, 4, 1)
    torch.cuda.synchronize(); assert bool((b == 1.5).all())
    b.fill_(float("nan"))
    engine.bli_dtrsm(LEFT, LOWER, 0, NONUNIT_DIAG, 6, 4, 0.0, a, 6, 1, b, 4, 1)          # alpha == 0: B := 0
    torch.cuda.synchronize(); assert bool((b == 0).all())
    # host operands
    an = gen.triangular("d", 150, 5, "frac", "r"); bn = gen.matrix("d", 150, 40, 6, "frac", "c", pad=3)
    want = bn.copy(order="K"); oracle.trsm(LEFT, UPPER, TRANSPOSE, UNIT_DIAG, 2.0, an, want)
    ta_, tb_ = to_torch(an, "cpu"), to_torch(bn, "cpu")
    engine.bli_dtrsm(LEFT, UPPER, TRANSPOSE, UNIT_DIAG, 150, 40, 2.0, ta_, *estr(an), tb_, *estr(bn))
    assert rel_err(to_numpy(tb_), want) <= 20 * TOL["d"]
    # object + BLAS layers
    from blis_b200 import api
    ta_, tb_ = to_torch(gen.triangular("d", 150, 5, "frac", "c")), to_torch(bn)
    o = api.Obj(ta_); api.bli_obj_set_uplo(LOWER, o)
    tb2 = tb_.clone(memory_format=torch.preserve_format)
    api.bli_trsm(LEFT, 2.0, o, api.Obj(tb_))
    api.dtrsm_("L", "L", "N", "N", 150, 40, 2.0, ta_, 150, tb2, tb2.stride(1))
    torch.cuda.synchronize()
    assert torch.equal(tb_, tb2)


@pytest.mark.parametrize("pin", [False, True])
def test_trsm_host_a_only_stored_triangle_travels(engine, oracle, pin):
    """Host-resident A with m >= 2048: the engine uploads column panels cut at the diagonal (stage_tri_to_device), so the
    other triangle -- NaN here -- neither travels nor is read; pageable and pinned, column- and row-stored, every
    uplo/trans (the effective triangle flips with trans), a ragged last panel."""
    m, n, seed = 2100, 72, 40
    b0 = gen.matrix("d", m, n, 7, "frac", "c")
    for oa in ("c", "r"):
        for uplo in (LOWER, UPPER):
            for tr in (NO_TRANSPOSE, TRANSPOSE):
                seed += 1
                a = gen.triangular("d", m, seed, "frac", oa)
                gen.poison_unstored(a, uplo == LOWER)
                want = b0.copy(order="K")
                oracle.trsm(LEFT, uplo, tr, NONUNIT_DIAG, 2.0, a, want)
                ta_, tb_ = to_torch(a, "cpu", pin=pin), to_torch(b0, "cpu", pin=pin)
                engine.bli_dtrsm(LEFT, uplo, tr, NONUNIT_DIAG, m, n, 2.0, ta_, *estr(a), tb_, *estr(b0))
                err = rel_err(to_numpy(tb_), want)
                assert err <= 20 * TOL["d"], (oa, uplo, tr, pin, err)
    # complex, conjugate-transposed, right side: A is n x n
    a = gen.triangular("z", 2050, 77, "frac", "c"); gen.poison_unstored(a, False)
    b = gen.matrix("z", 40, 2050, 78, "frac", "c")
    want = b.copy(order="K"); oracle.trsm(RIGHT, UPPER, CONJ_TRANSPOSE, UNIT_DIAG, 2.0 + 0.3j, a, want)
    ta_, tb_ = to_torch(a, "cpu", pin=pin), to_torch(b, "cpu", pin=pin)
    engine.bli_ztrsm(RIGHT, UPPER, CONJ_TRANSPOSE, UNIT_DIAG, 40, 2050, 2.0 + 0.3j, ta_, *estr(a), tb_, *estr(b))
    assert rel_err(to_numpy(tb_), want) <= 20 * TOL["z"]


def test_trsm_pinned_host_operands_pipelined(engine, ref):
    """Large solves with pinned host operands take the pipelined path (host_trsm.cuh: trsm_host_rowpipe, the engine's own
    block rows: 256 at this size): B travels in row blocks, A in the order the solve reads it (one event per launch), X
    comes back block by block.  Against the real reference library; the unstored triangle of A is NaN; column- and
    row-stored A, every uplo/trans, A already on the device, the right-sided form on a row-stored B; and the result must
    equal the sequential path's (trsm_host_pipe = 0) to rounding (the updates are grouped differently)."""
    m, n, seed = 4608, 2200, 140
    b0 = gen.matrix("d", m, n, 17, "frac", "c")
    n0 = engine.launch_count()
    for oa in ("c", "r"):
        for uplo in (LOWER, UPPER):
            for tr in (NO_TRANSPOSE, TRANSPOSE):
                seed += 1
                a = gen.triangular("d", m, seed, "frac", oa)
                gen.poison_unstored(a, uplo == LOWER)
                want = b0.copy(order="K")
                ref.trsm(LEFT, uplo, tr, NONUNIT_DIAG, 2.0, a, want)
                ta_, tb_ = to_torch(a, "cpu", pin=True), to_torch(b0, "cpu", pin=True)
                engine.bli_dtrsm(LEFT, uplo, tr, NONUNIT_DIAG, m, n, 2.0, ta_, *estr(a), tb_, *estr(b0))
                got = to_numpy(tb_)
                err = rel_err(got, want)
                assert err <= 20 * TOL["d"], (oa, uplo, tr, err)
                if oa == "c" and tr == NO_TRANSPOSE:
                    # the sequential path on the same operands, and A resident on the device with B on the host
                    engine.set_option("trsm_host_pipe", 0)
                    try:
                        tb2 = to_torch(b0, "cpu", pin=True)
                        engine.bli_dtrsm(LEFT, uplo, tr, NONUNIT_DIAG, m, n, 2.0, ta_, *estr(a), tb2, *estr(b0))
                    finally:
                        engine.set_option("trsm_host_pipe", 1)
                    assert rel_err(to_numpy(tb2), got) <= 20 * TOL["d"], (uplo, rel_err(to_numpy(tb2), got))
                    tb3 = to_torch(b0, "cpu", pin=True)
                    engine.bli_dtrsm(LEFT, uplo, tr, NONUNIT_DIAG, m, n, 2.0, ta_.cuda(), *estr(a), tb3, *estr(b0))
                    assert rel_err(to_numpy(tb3), want) <= 20 * TOL["d"], (uplo, "device A")
    assert engine.launch_count() > n0
    # right side, row-stored B (the transposed problem is column-stored): X * A^T = alpha * B, unit diagonal
    a = gen.triangular("d", m, 991, "frac", "c"); gen.poison_unstored(a, True)
    b = gen.matrix("d", 1100, m, 992, "frac", "r")
    want = b.copy(order="K"); ref.trsm(RIGHT, LOWER, TRANSPOSE, UNIT_DIAG, 2.0, a, want)
    ta_, tb_ = to_torch(a, "cpu", pin=True), to_torch(b, "cpu", pin=True)
    engine.bli_dtrsm(RIGHT, LOWER, TRANSPOSE, UNIT_DIAG, 1100, m, 2.0, ta_, *estr(a), tb_, *estr(b))
    assert rel_err(to_numpy(tb_), want) <= 20 * TOL["d"]


def test_trsm_pinned_host_operands_row_blocks(engine, ref):
    """Tall solves with pinned host operands run the row-block (left-looking) pipeline (host_trsm.cuh: trsm_host_rowpipe):
    B travels in row blocks, each block row receives all its updates in one gemm reading A[j, 0:j] and is then solved by
    the recursion, X goes home block by block.  Forced at a size the reference solves in seconds (trsm_host_rb = 1024 on
    m = 4608: four whole block rows and a ragged one; 1280 on m = 5000: nothing divides).  Against the real reference
    library, the unstored triangle of A NaN; column- and row-stored A, every uplo/trans, unit diagonal, A resident on the
    device, the right-sided form; an integer-valued system must come back bit for bit (the arithmetic per element is the
    reference's sequence whatever the blocking)."""
    n0 = engine.launch_count()
    try:
        for m, n, rb in ((4608, 2200, 1024), (5000, 1100, 1280)):
            engine.set_option("trsm_host_rb", rb)
            b0 = gen.matrix("d", m, n, 23, "frac", "c")
            seed = 240 + m
            for oa in ("c", "r"):
                for uplo in (LOWER, UPPER):
                    for tr in (NO_TRANSPOSE, TRANSPOSE):
                        seed += 1
                        diag = UNIT_DIAG if (seed % 3 == 0) else NONUNIT_DIAG
                        a = gen.triangular("d", m, seed, "frac", oa)
                        gen.poison_unstored(a, uplo == LOWER)
                        want = b0.copy(order="K")
                        ref.trsm(LEFT, uplo, tr, diag, 2.0, a, want)
                        ta_, tb_ = to_torch(a, "cpu", pin=True), to_torch(b0, "cpu", pin=True)
                        engine.bli_dtrsm(LEFT, uplo, tr, diag, m, n, 2.0, ta_, *estr(a), tb_, *estr(b0))
                        err = rel_err(to_numpy(tb_), want)
                        assert err <= 20 * TOL["d"], (m, oa, uplo, tr, err)
                        if oa == "c" and tr == NO_TRANSPOSE:
                            tb3 = to_torch(b0, "cpu", pin=True)
                            engine.bli_dtrsm(LEFT, uplo, tr, diag, m, n, 2.0, ta_.cuda(), *estr(a), tb3, *estr(b0))
                            assert rel_err(to_numpy(tb3), want) <= 20 * TOL["d"], (m, uplo, "device A")
        # PAGEABLE operands (what a legacy dtrsm_ caller passes) ride the same pipeline through the pinned ring, X unpacked one
        # block row late: pageable A and B, pageable B with page-locked A, row-stored pageable A; the engine's own block rows
        engine.set_option("trsm_host_rb", 0)
        m, n = 4608, 2200
        b0 = gen.matrix("d", m, n, 29, "frac", "c")
        for oa, uplo, tr, pin_a in (("c", LOWER, NO_TRANSPOSE, False), ("c", UPPER, NO_TRANSPOSE, True), ("r", LOWER, TRANSPOSE, False),
                                    ("r", UPPER, NO_TRANSPOSE, False)):
            a = gen.triangular("d", m, 700 + uplo + tr + pin_a, "frac", oa)
            gen.poison_unstored(a, uplo == LOWER)
            want = b0.copy(order="K")
            ref.trsm(LEFT, uplo, tr, NONUNIT_DIAG, 2.0, a, want)
            ta_ = to_torch(a, "cpu", pin=True) if pin_a else torch.from_numpy(a.copy(order="K"))
            bh = b0.copy(order="K"); tb_ = torch.from_numpy(bh)
            assert not tb_.is_pinned() and tb_.stride() == (1, m)
            engine.bli_dtrsm(LEFT, uplo, tr, NONUNIT_DIAG, m, n, 2.0, ta_, *estr(a), tb_, *estr(b0))
            assert rel_err(bh, want) <= 20 * TOL["d"], ("pageable", oa, uplo, tr, rel_err(bh, want))
        # another datatype through the same pipeline (32-row leaves, conjugated A)
        engine.set_option("trsm_host_rb", 1000)
        m, n = 4100, 1030
        a = gen.triangular("z", m, 995, "frac", "c"); gen.poison_unstored(a, False)
        b = gen.matrix("z", m, n, 996, "frac", "c")
        want = b.copy(order="K"); ref.trsm(LEFT, UPPER, CONJ_NO_TRANSPOSE, NONUNIT_DIAG, 2.0 - 1.0j, a, want)
        ta_, tb_ = to_torch(a, "cpu", pin=True), to_torch(b, "cpu", pin=True)
        engine.bli_ztrsm(LEFT, UPPER, CONJ_NO_TRANSPOSE, NONUNIT_DIAG, m, n, 2.0 - 1.0j, ta_, *estr(a), tb_, *estr(b))
        assert rel_err(to_numpy(tb_), want) <= 20 * TOL["z"]
        # right side, row-stored B (the transposed problem is column-stored): X * A^T = alpha * B
        engine.set_option("trsm_host_rb", 1024)
        m = 4608
        a = gen.triangular("d", m, 993, "frac", "c"); gen.poison_unstored(a, True)
        b = gen.matrix("d", 1100, m, 994, "frac", "r")
        want = b.copy(order="K"); ref.trsm(RIGHT, LOWER, TRANSPOSE, NONUNIT_DIAG, 2.0, a, want)
        ta_, tb_ = to_torch(a, "cpu", pin=True), to_torch(b, "cpu", pin=True)
        engine.bli_dtrsm(RIGHT, LOWER, TRANSPOSE, NONUNIT_DIAG, 1100, m, 2.0, ta_, *estr(a), tb_, *estr(b))
        assert rel_err(to_numpy(tb_), want) <= 20 * TOL["d"]
        # exact system: unit lower triangle of small integers, integer right-hand sides, alpha = 1 -> every intermediate
        # value is an integer far below 2^53 for this construction (strictly lower part sparse: one entry per row)
        rng = np.random.default_rng(5)
        a = np.asfortranarray(np.eye(m))
        rows = np.arange(1, m); a[rows, rng.integers(0, rows)] = rng.integers(-1, 2, m - 1)
        x = np.asfortranarray(rng.integers(-4, 5, (m, 1100)).astype(np.float64))
        b = np.asfortranarray(a @ x)
        assert np.abs(b).max() < 2.0 ** 40
        ta_, tb_ = to_torch(a, "cpu", pin=True), to_torch(b, "cpu", pin=True)
        engine.bli_dtrsm(LEFT, LOWER, NO_TRANSPOSE, UNIT_DIAG, m, 1100, 1.0, ta_, *estr(a), tb_, *estr(b))
        assert np.array_equal(to_numpy(tb_), x)
    finally:
        engine.set_option("trsm_host_rb", 0)
    assert engine.launch_count() > n0


def test_trsm_full_size_testsuite_residual(engine):
    """BASELINE config #4: dtrsm left/lower/notrans/nonunit m=32768 n=8192.
    resid = || B t - alpha inv(A) (B0 t) ||  ==  || A (X t) - alpha B0 t || scaled, via a triangular
    mat-vec instead of trsv (testsuite/src/test_trsm.c:362-381); pass threshold 1e-14 (:44-47).
    Also encode->decode: A @ X reproduces alpha*B0 (round trip through torch's fp64 trmm-free matmul)."""
    dev = "cuda"
    m, n = 32768, 8192
    g = torch.Generator(device=dev); g.manual_seed(int(0xB200))
    a = (torch.rand(m, m, dtype=torch.float64, device=dev, generator=g) * 2 - 1)
    a = a / float(2 ** np.ceil(np.log2(float(a.abs().sum(dim=1).max()))))     # mobj_randomize normalisation
    a.diagonal().add_(2.0)                                                   # test_libblis.c:2583-2589
    a = torch.tril(a).t().contiguous().t()                                    # column-major lower
    b = (torch.rand(n, m, dtype=torch.float64, device=dev, generator=g) * 2 - 1).t()
    b = b / float(2 ** np.ceil(np.log2(float(b.abs().sum(dim=0).max()))))
    b0 = b.clone(memory_format=torch.preserve_format)
    engine.bli_dtrsm(LEFT, LOWER, NO_TRANSPOSE, NONUNIT_DIAG, m, n, 2.0, a, 1, m, b, 1, m)
    torch.cuda.synchronize()
    t = ((torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1) / n)
    xt, bt = b @ t, 2.0 * (b0 @ t)
    # bring both sides to the reference's form: w = inv(A)*bt by a host-checked triangular solve
    w = torch.linalg.solve_triangular(a, bt.unsqueeze(1), upper=False).squeeze(1)
    resid = float(torch.linalg.vector_norm(xt - w))
    assert resid <= 1e-14, resid
    back = float(torch.linalg.vector_norm(a @ xt - bt))
    assert back <= 1e-13, back


def test_trsm_fused_panel_kernel(engine, oracle):
    """dtrsm's fused 256-row diagonal-panel kernel (trsm_panel.cuh: in-panel DMMA updates + warp-shuffle substitution),
    proven by name to be the kernel that ran: ragged panels and ragged column tiles, every side/uplo/trans/diag, row-,
    column- and general-stride B, row-/column-stored A, unstored triangle NaN-poisoned -- against the oracle; same result
    (to tolerance) as the 64-row block-solve path it replaces."""
    seed = 9000
    for (m, n) in ((65, 8), (130, 70), (256, 64), (255, 129), (300, 65), (513, 200), (1000, 96)):
        for side in (LEFT, RIGHT):
            for uplo in (LOWER, UPPER):
                for tr in (NO_TRANSPOSE, TRANSPOSE):
                    for dg in (NONUNIT_DIAG, UNIT_DIAG):
                        for (oa, ob) in (("c", "c"), ("r", "r"), ("c", "g")):
                            if m * n > 30000 and ((oa, ob) != ("c", "c") or dg == UNIT_DIAG):
                                continue
                            seed += 1
                            mm, nn = (m, n) if side == LEFT else (n, m)
                            a = gen.triangular("d", m, seed, "frac", oa)
                            gen.poison_unstored(a, uplo == LOWER)
                            b = gen.matrix("d", mm, nn, seed + 7000, "frac", ob, pad=1)
                            want = b.copy(order="K")
                            oracle.trsm(side, uplo, tr, dg, 2.0, a, want)
                            engine.kernel_stats(reset=True)
                            got = run_trsm(engine, "d", side, uplo, tr, dg, 2.0, a, b)
                            ks = engine.kernel_stats()
                            assert any(k.startswith("trsm_panel_kernel") for k in ks), ks
                            assert rel_err(got, want) <= 20 * TOL["d"], (m, n, side, uplo, tr, dg, oa, ob, rel_err(got, want))
                            engine.set_option("trsm_fused", 0)
                            try:
                                got2 = run_trsm(engine, "d", side, uplo, tr, dg, 2.0, a, b)
                            finally:
                                engine.set_option("trsm_fused", 1)
                            assert rel_err(got, got2) <= 20 * TOL["d"]


def test_trsm_fused_panel_integer_system_bit_exact(engine, oracle):
    """Integer-valued banded systems (sub-diagonals 1 and 67, entries +-1, diagonal +-1): every intermediate is an exact
    integer, so the fused kernel (tensor-pipe updates, shuffled substitution, any summation order) must reproduce the
    reference's bits, lower and upper, across panel and block boundaries."""
    m, n = 600, 80
    i = np.arange(m)
    for idx, (uplo, tr) in enumerate(((LOWER, NO_TRANSPOSE), (UPPER, NO_TRANSPOSE), (LOWER, TRANSPOSE), (UPPER, TRANSPOSE))):
        a = np.zeros((m, m), order="F")
        a[i, i] = 1.0 - 2.0 * (i % 2)
        a[i[1:], i[:-1]] = 1.0 - 2.0 * ((i[1:] // 3) % 2)
        a[i[67:], i[:-67]] = 1.0 - 2.0 * ((i[67:] // 5) % 2)
        if uplo == UPPER:
            a = np.asfortranarray(a.T)
        gen.poison_unstored(a, uplo == LOWER)
        b = gen.matrix("d", m, n, 650 + idx, "ints")
        want = b.copy(order="K"); oracle.trsm(LEFT, uplo, tr, NONUNIT_DIAG, 2.0, a, want)
        assert np.isfinite(want).all() and np.abs(want).max() < 2.0 ** 50
        engine.kernel_stats(reset=True)
        got = run_trsm(engine, "d", LEFT, uplo, tr, NONUNIT_DIAG, 2.0, a, b)
        assert any(k.startswith("trsm_panel_kernel") for k in engine.kernel_stats())
        assert np.array_equal(got, want), (uplo, tr, np.abs(got - want).max())
