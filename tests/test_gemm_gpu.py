"""Parity of the CUDA gemm (through the C ABI) with the reference.

Checkers: the committed golden fixtures (real reference outputs), the oracle
restatement, and the real reference library when it travelled (oracle/_ref).
Bars: bit-exact on power-of-two inputs (any summation order must agree);
elementwise TOL otherwise (util.TOL: 1e-12 for d/z, 5e-5 for s/c, relative to
max(1,|ref|max)); at BASELINE sizes the testsuite's own residual
(testsuite/src/test_gemm.c:393-401) with its pass thresholds (:44-47).
"""
import numpy as np
import pytest
import torch

import gen
import make_golden as G
from refblis import CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE, NO_TRANSPOSE, TRANSPOSE
from util import NP2T, TOL, estr, rel_err, to_numpy, to_torch

pytestmark = pytest.mark.gpu
GOLD = G.HERE
GEMM = {"s": "bli_sgemm", "d": "bli_dgemm", "c": "bli_cgemm", "z": "bli_zgemm"}


def run_gemm(engine, ch, ta, tb, alpha, a, b, beta, c, device="cuda"):
    """a, b, c: numpy arrays (any strides).  Returns C as numpy."""
    ta_, tb_, tc_ = to_torch(a, device), to_torch(b, device), to_torch(c, device)
    m, n = c.shape
    k = a.shape[0] if (ta & TRANSPOSE) else a.shape[1]
    getattr(engine, GEMM[ch])(ta, tb, m, n, k, alpha, ta_, *estr(a), tb_, *estr(b), beta, tc_, *estr(c))
    if device == "cuda":
        torch.cuda.synchronize()
    return to_numpy(tc_)


def test_gemm_golden_fixtures(engine):
    gold = np.load(GOLD / "gemm.npz")
    for idx, cs in enumerate(G.gemm_cases()):
        ch, kind = cs[0], cs[1]
        a, b, c = G.gemm_inputs(cs, idx)
        got = run_gemm(engine, ch, cs[5], cs[6], cs[10], a, b, cs[11], c)
        want = gold[f"c{idx}"]
        if kind == "pow2":
            assert np.array_equal(got, want), f"gemm golden case {idx} {cs}: not bit-exact"
        else:
            assert rel_err(got, want) <= TOL[ch], f"gemm golden case {idx} {cs}: {rel_err(got, want)}"


SHAPES = [(1, 1, 1), (2, 3, 4), (8, 8, 4), (127, 129, 65), (128, 128, 16), (129, 127, 17), (256, 384, 100),
          (300, 77, 513), (1, 700, 33), (515, 1, 64), (64, 1000, 1), (33, 65, 1025)]


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_vs_oracle_all_params(engine, oracle, ch):
    """Every trans/conj combination x storage combination x edge shapes."""
    cx = ch in "cz"
    trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
    seed = 100
    for (m, n, k) in SHAPES:
        for ta in trs:
            for tb in trs:
                for (oa, ob, oc) in (("c", "c", "c"), ("r", "r", "r"), ("c", "r", "g"), ("g", "c", "r")):
                    if m * n * k > 2_000_000 and (oa, ob, oc) != ("c", "c", "c"):
                        continue
                    seed += 1
                    am, ak = (k, m) if ta & TRANSPOSE else (m, k)
                    bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
                    a = gen.matrix(ch, am, ak, seed, "frac", oa, pad=1)
                    b = gen.matrix(ch, bk, bn, seed + 5000, "frac", ob, pad=2)
                    c = gen.matrix(ch, m, n, seed + 9000, "frac", oc, pad=3)
                    want = c.copy(order="K")
                    oracle.gemm(ta, tb, al, a, b, be, want)
                    got = run_gemm(engine, ch, ta, tb, al, a, b, be, c)
                    assert rel_err(got, want) <= TOL[ch], (ch, m, n, k, ta, tb, oa, ob, oc, rel_err(got, want))


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_pow2_bit_exact_vs_oracle(engine, oracle, ch):
    """Power-of-two inputs: products and sums are exact, so the tile order of the
    GPU kernel and the KC-blocked order of the reference must give identical bits."""
    for idx, (m, n, k, ta, tb, oc) in enumerate(((257, 131, 64, 0, 0, "c"), (130, 260, 48, 8, 0, "r"),
                                                 (64, 64, 64, 0, 8, "c"), (513, 9, 33, 8, 8, "c"))):
        am, ak = (k, m) if ta & TRANSPOSE else (m, k)
        bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
        a = gen.matrix(ch, am, ak, 700 + idx, "pow2"); b = gen.matrix(ch, bk, bn, 800 + idx, "pow2")
        c = gen.matrix(ch, m, n, 900 + idx, "pow2", oc)
        want = c.copy(order="K")
        oracle.gemm(ta, tb, 2.0, a, b, 0.5, want)
        got = run_gemm(engine, ch, ta, tb, 2.0, a, b, 0.5, c)
        assert np.array_equal(got, want), (ch, m, n, k, ta, tb)


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_vs_real_reference(engine, ref, ch):
    """Against the real reference BLIS (optimized CPU kernels) on the same inputs."""
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if ch in "cz" else (2.0, 1.2))
    for idx, (m, n, k) in enumerate(((1000, 1000, 1000), (511, 769, 300))):      # config #1 size: dgemm 1000^3
        a = gen.matrix(ch, m, k, 40 + idx, "frac"); b = gen.matrix(ch, k, n, 50 + idx, "frac")
        c = gen.matrix(ch, m, n, 60 + idx, "frac")
        want = c.copy(order="K")
        ref.gemm(0, 0, al, a, b, be, want)
        got = run_gemm(engine, ch, 0, 0, al, a, b, be, c)
        assert rel_err(got, want) <= TOL[ch] * 4, (ch, m, n, k, rel_err(got, want))


def test_gemm_trivial_and_special_cases(engine):
    """Empty dims, k == 0, alpha == 0 (C := beta*C), beta == 0 must not read C
    (docs/KernelsHowTo.md:342; frame/3/bli_l3_util.c:40-65)."""
    dev = "cuda"
    a = torch.ones(5, 4, dtype=torch.float64, device=dev); b = torch.ones(4, 3, dtype=torch.float64, device=dev)
    c = torch.full((5, 3), 3.0, dtype=torch.float64, device=dev)
    engine.bli_dgemm(0, 0, 0, 3, 4, 1.0, a, 4, 1, b, 3, 1, 1.0, c, 3, 1)         # m == 0: no-op
    engine.bli_dgemm(0, 0, 5, 0, 4, 1.0, a, 4, 1, b, 3, 1, 1.0, c, 3, 1)         # n == 0: no-op
    torch.cuda.synchronize(); assert bool((c == 3.0).all())
    engine.bli_dgemm(0, 0, 5, 3, 0, 1.0, a, 4, 1, b, 3, 1, 0.5, c, 3, 1)         # k == 0: C := beta*C
    torch.cuda.synchronize(); assert bool((c == 1.5).all())
    engine.bli_dgemm(0, 0, 5, 3, 4, 0.0, a, 4, 1, b, 3, 1, 2.0, c, 3, 1)         # alpha == 0
    torch.cuda.synchronize(); assert bool((c == 3.0).all())
    for dt, fn in ((torch.float32, engine.bli_sgemm), (torch.float64, engine.bli_dgemm),
                   (torch.complex64, engine.bli_cgemm), (torch.complex128, engine.bli_zgemm)):
        a = torch.ones(70, 50, dtype=dt, device=dev); b = torch.ones(50, 90, dtype=dt, device=dev)
        c = torch.full((70, 90), float("nan"), dtype=dt, device=dev)
        fn(0, 0, 70, 90, 50, 1.0, a, 50, 1, b, 90, 1, 0.0, c, 90, 1)
        torch.cuda.synchronize()
        assert bool((c == 50).all()), dt
        c.fill_(float("nan"))
        fn(0, 0, 70, 90, 50, 0.0, a, 50, 1, b, 90, 1, 0.0, c, 90, 1)               # alpha == beta == 0: C := 0
        torch.cuda.synchronize()
        assert bool((c == 0).all()), dt


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_host_operands_pinned_staging(engine, oracle, ch):
    """Host pointers (pageable and pinned, column/row/general storage) go through
    the engine's pinned staging and come back in the caller's buffer."""
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if ch in "cz" else (2.0, 1.2))
    for idx, (oa, ob, oc, pin) in enumerate((("c", "c", "c", False), ("r", "c", "r", False), ("g", "r", "g", False), ("c", "c", "c", True))):
        m, n, k = 211, 97, 150
        a = gen.matrix(ch, m, k, 300 + idx, "frac", oa, pad=2); b = gen.matrix(ch, k, n, 310 + idx, "frac", ob)
        c = gen.matrix(ch, m, n, 320 + idx, "frac", oc, pad=1)
        want = c.copy(order="K"); oracle.gemm(0, 0, al, a, b, be, want)
        ta_, tb_, tc_ = (to_torch(x, "cpu", pin=pin) for x in (a, b, c))
        assert tc_.is_pinned() == pin
        getattr(engine, GEMM[ch])(0, 0, m, n, k, al, ta_, *estr(a), tb_, *estr(b), be, tc_, *estr(c))
        assert rel_err(to_numpy(tc_), want) <= TOL[ch], (ch, oa, ob, oc, pin)


def test_gemm_object_and_blas_layers(engine, oracle):
    """The object API (bli_gemm on Obj) and the BLAS layer (dgemm_) reach the same kernel."""
    from blis_b200 import api
    m, n, k = 150, 130, 70
    a = gen.matrix("d", k, m, 1, "frac"); b = gen.matrix("d", k, n, 2, "frac"); c = gen.matrix("d", m, n, 3, "frac")
    want = c.copy(order="K"); oracle.gemm(TRANSPOSE, 0, 2.0, a, b, 1.2, want)
    ta_, tb_, tc_ = to_torch(a), to_torch(b), to_torch(c)
    ao = api.Obj(ta_); api.bli_obj_set_conjtrans(TRANSPOSE, ao)
    api.bli_gemm(2.0, ao, api.Obj(tb_), 1.2, api.Obj(tc_))
    torch.cuda.synchronize()
    assert rel_err(to_numpy(tc_), want) <= TOL["d"]
    tc2 = to_torch(c)
    api.dgemm_("T", "N", m, n, k, 2.0, ta_, k, tb_, k, 1.2, tc2, m)
    torch.cuda.synchronize()
    assert np.array_equal(to_numpy(tc2), to_numpy(tc_))


def _testsuite_resid(alpha, a, b, beta, c0, c):
    """resid = || C t - ( beta C0 t + alpha A (B t) ) ||_F, t random  (testsuite/src/test_gemm.c:393-401)."""
    n = c.shape[1]
    g = torch.Generator(device=c.device); g.manual_seed(7)
    t = torch.rand(n, dtype=torch.float64, device=c.device, generator=g) * 2 - 1
    t = (t / n).to(c.dtype)
    z = beta * (c0 @ t) + alpha * (a @ (b @ t))
    return float(torch.linalg.vector_norm(c @ t - z))


@pytest.mark.parametrize("ch,n", [("d", 16384), ("z", 8192), ("s", 16384), ("c", 8192)])
def test_gemm_full_size_testsuite_residual(engine, ch, n):
    """BASELINE config #2/#3 sizes, column-major, device-resident: the reference
    testsuite's own randomized residual and pass thresholds, plus a linearity
    property: gemm(alpha) + gemm(alpha) into the same C == gemm(2*alpha)."""
    dt = NP2T[np.dtype(gen.NP_DT[ch])]
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(int(0xB200))
    rdt = torch.float32 if ch in "sc" else torch.float64

    def rnd(m, k):
        x = torch.rand(k, m, dtype=rdt, device=dev, generator=g) * 2 - 1
        if ch in "cz":
            x = torch.complex(x, torch.rand(k, m, dtype=rdt, device=dev, generator=g) * 2 - 1)
        # libblis_test_mobj_randomize: normalise by the 1-norm rounded up to a power of two (test_libblis.c:2529-2565)
        nrm = float(x.abs().sum(dim=1).max())
        return (x / float(2 ** np.ceil(np.log2(nrm)))).t()

    a, b, c = rnd(n, n), rnd(n, n), rnd(n, n)
    c0 = c.clone()
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if ch in "cz" else (2.0, 1.2))
    getattr(engine, GEMM[ch])(0, 0, n, n, n, al, a, 1, n, b, 1, n, be, c, 1, n)
    torch.cuda.synchronize()
    resid = _testsuite_resid(al, a, b, be, c0, c)
    thresh = 1e-5 if ch in "sc" else 1e-14                     # testsuite/src/test_gemm.c:44-47 "pass"
    assert resid <= thresh, (ch, n, resid)
    # linearity (size independent): C1 = 0 + al*A*B twice accumulated == 2*al*A*B
    m2 = 4096
    c1 = torch.zeros(m2, m2, dtype=dt, device=dev).t(); c2 = torch.zeros(m2, m2, dtype=dt, device=dev).t()
    f = getattr(engine, GEMM[ch])
    f(0, 0, m2, m2, n, 1.0, a, 1, n, b, 1, n, 0.0, c1, 1, m2)
    f(0, 0, m2, m2, n, 1.0, a, 1, n, b, 1, n, 1.0, c1, 1, m2)
    f(0, 0, m2, m2, n, 2.0, a, 1, n, b, 1, n, 0.0, c2, 1, m2)
    torch.cuda.synchronize()
    assert bool(torch.equal(c1, c2)), "x + x != 2x: gemm is not deterministic/linear"


@pytest.mark.parametrize("pin", [False, True])
def test_gemm_host_operands_pipelined_blocks(engine, pin):
    """Large host problem: the engine pipelines column blocks of B/C (H2D, kernels, D2H on three
    streams).  Ragged last block, beta != 0 and beta == 0, pageable and pinned memory."""
    m, n, k = 1000, 1100 + 37, 2000
    a = gen.matrix("d", m, k, 71, "frac"); b = gen.matrix("d", k, n, 72, "frac", pad=3)
    for beta in (1.2, 0.0):
        c = gen.matrix("d", m, n, 73, "frac", pad=1)
        want = beta * c + 2.0 * (a @ b)
        ta_, tb_, tc_ = (to_torch(x, "cpu", pin=pin) for x in (a, b, c))
        if beta == 0.0:
            tc_.fill_(float("nan"))
        engine.bli_dgemm(0, 0, m, n, k, 2.0, ta_, *estr(a), tb_, *estr(b), beta, tc_, *estr(c))
        assert rel_err(to_numpy(tc_), want) <= TOL["d"], (pin, beta)
    # k >= 2048: a host-resident A also moves in k panels and the first column block is accumulated panel by panel
    # (ragged last panel, transposed A, row-stored B)
    m, n, k = 900, 1100, 2500
    a = gen.matrix("d", k, m, 74, "frac", pad=2); b = gen.matrix("d", k, n, 75, "frac", "r")
    for beta in (1.2, 0.0):
        c = gen.matrix("d", m, n, 76, "frac")
        want = beta * c + 2.0 * (a.T @ b)
        ta_, tb_, tc_ = (to_torch(x, "cpu", pin=pin) for x in (a, b, c))
        if beta == 0.0:
            tc_.fill_(float("nan"))
        engine.bli_dgemm(TRANSPOSE, 0, m, n, k, 2.0, ta_, *estr(a), tb_, *estr(b), beta, tc_, *estr(c))
        assert rel_err(to_numpy(tc_), want) <= TOL["d"], (pin, beta, "k panels")
    # a row-major caller (A, B, C row-stored): the engine solves the transposed problem, whose operands are column-stored,
    # so the same pipelines serve it (m = 1300 becomes the pipelined dimension); complex with a conjugated operand too
    m, n, k = 1300, 700, 2100
    for ch, ta, al, be in (("d", NO_TRANSPOSE, 2.0, 1.2), ("z", CONJ_NO_TRANSPOSE, 2.0 + 0.2j, 1.2 + 0.5j)):
        a = gen.matrix(ch, m, k, 77, "frac", "r", pad=1); b = gen.matrix(ch, k, n, 78, "frac", "r"); c = gen.matrix(ch, m, n, 79, "frac", "r", pad=2)
        want = be * c + al * ((a.conj() if ta == CONJ_NO_TRANSPOSE else a) @ b)
        ta_, tb_, tc_ = (to_torch(x, "cpu", pin=pin) for x in (a, b, c))
        getattr(engine, GEMM[ch])(ta, 0, m, n, k, al, ta_, *estr(a), tb_, *estr(b), be, tc_, *estr(c))
        assert rel_err(to_numpy(tc_), want) <= TOL[ch] * 4, (pin, ch, "row-major caller")


@pytest.mark.parametrize("kdim", [8200 + 40, 8192])
@pytest.mark.parametrize("pin", [False, True])
def test_gemm_host_operands_k_panel_pipeline(engine, pin, kdim):
    """Long k with host operands: the engine accumulates over k panels (A and B move panel by panel, the host C is staged
    separately and merged in the last round, finished column blocks leave under the next block's kernels).  Ragged last
    panel and last column block, transposed A, beta != 0 and beta == 0, pageable and pinned memory; same result with the
    k-panel pipeline switched off (column-block pipeline)."""
    m, n, k = 1800, 4100 + 28, kdim             # 8240: ragged last panel, the final phase takes one panel; 8192: the last two, as one k-panel launch
    rng = np.random.default_rng(5)
    a = np.asfortranarray(rng.uniform(-1, 1, (k, m))); b = np.asfortranarray(rng.uniform(-1, 1, (k, n)))
    ab = torch.from_numpy(a).t() @ torch.from_numpy(b)
    for beta in (1.2, 0.0):
        c = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
        want = (beta * torch.from_numpy(c) + 2.0 * ab).numpy()
        outs = []
        for kpipe in (1, 0):
            engine.set_option("host_kpipe", kpipe)
            try:
                ta_, tb_, tc_ = (to_torch(x, "cpu", pin=pin) for x in (a, b, c))
                if beta == 0.0:
                    tc_.fill_(float("nan"))
                n0 = engine.launch_count()
                engine.bli_dgemm(TRANSPOSE, 0, m, n, k, 2.0, ta_, *estr(a), tb_, *estr(b), beta, tc_, *estr(c))
                launches = engine.launch_count() - n0
            finally:
                engine.set_option("host_kpipe", 1)
            outs.append(to_numpy(tc_))
            assert rel_err(outs[-1], want) <= 4 * TOL["d"], (pin, beta, kpipe)
            if kpipe:
                # the schedule of gemm_host_kpipe (host_gemm.cuh): k in eighths rounded to 128, the first one cut at a quarter; the
                # final phase covers the last two panels when they are equally wide; 7 column blocks of n
                kb = max(512, -(-(-(-k // 8)) // 128) * 128)
                pk = [0] + ([kb // 4 // 128 * 128] if kb >= 1024 and kb // 4 // 128 * 128 >= 256 else [])
                while pk[-1] < k:
                    pk.append(min(k, pk[-1] + kb - (pk[-1] if len(pk) == 2 and pk[-1] < kb else 0)))
                npan = len(pk) - 1
                tail = 2 if npan >= 3 and pk[-1] - pk[-2] == pk[-2] - pk[-3] else 1
                assert (npan, tail) == ((9, 1) if k == 8240 else (9, 2))
                assert launches == (npan - tail) + 7 + (7 if beta != 0.0 else 0), ("k-panel pipeline: full rounds + 7 column blocks (+ 7 merges of C)", launches)


@pytest.mark.parametrize("ch", ["d", "z"])
def test_gemm_kpanels_accumulates_like_blk_var3(engine, oracle, ch):
    """b200_gemm_kpanels == the reference's pc loop: one rank-k update per panel with beta reset
    to one after the first (frame/3/gemm/bli_gemm_blk_var3.c:110-112), folded into one launch."""
    from blis_b200 import api
    m, n, k, npan = 210, 150, 96, 5
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if ch == "z" else (2.0, 1.2))
    for (oa, ob, oc) in (("c", "c", "c"), ("r", "c", "r")):
        a_p = [gen.matrix(ch, m, k, 400 + s, "frac", oa) for s in range(npan)]
        b_p = [gen.matrix(ch, k, n, 500 + s, "frac", ob) for s in range(npan)]
        c = gen.matrix(ch, m, n, 600, "frac", oc)
        want = c.copy(order="K")
        for s in range(npan):
            oracle.gemm(0, 0, al, a_p[s], b_p[s], be if s == 0 else 1.0, want)
        ta = [to_torch(x) for x in a_p]; tb = [to_torch(x) for x in b_p]; tc = to_torch(c)
        api.bli_gemm_kpanels(NP2T[np.dtype(gen.NP_DT[ch])], 0, 0, m, n, k, al, ta, *estr(a_p[0]), tb, *estr(b_p[0]), be, tc, *estr(c))
        torch.cuda.synchronize()
        assert rel_err(to_numpy(tc), want) <= TOL[ch], (ch, oa, ob, oc)


@pytest.mark.parametrize("tr", [(0, 0), (TRANSPOSE, 0), (0, TRANSPOSE), (TRANSPOSE, TRANSPOSE)])
def test_gemm_kpanels_tma_slots_of_one_buffer(engine, oracle, tr):
    """k panels that are slots of ONE strided buffer (the receive buffers of b200_dist_gemm) are served by the TMA kernel
    through 3-D tensor maps whose third coordinate is the slot (gemm_dmma_tma.cuh: tma_load_3d; gemm_launch.cuh:
    seg_slots / make_tmap3): all four staging orientations, slots in a scrambled order with gaps (and different orders
    for A and B), ragged m, n and per-panel k, beta != 0 with a short total k (CST: D staged through the ring) and
    beta == 0 on NaN-poisoned C, against the oracle's pc loop (one rank-k update per panel, beta then one:
    frame/3/gemm/bli_gemm_blk_var3.c:110-112).  The kernel name proves the path."""
    from blis_b200 import api
    ta, tb = tr
    engine.set_option("dgemm_cfg", 9)
    try:
        for (m, n, k, npan, be) in ((260, 388, 36, 5, 1.2), (516, 260, 28, 8, 0.0), (130, 140, 200, 7, 1.2), (388, 260, 68, 5, 0.0)):
            am, ak = (k, m) if ta else (m, k)
            bk, bn = (n, k) if tb else (k, n)
            slots_a = [9, 2, 5, 0, 7, 3, 11, 6][:npan]
            slots_b = [1, 4, 0, 8, 2, 6, 3, 10][:npan]
            abuf = torch.full((12, ak, am), float("nan"), dtype=torch.float64, device="cuda")      # slot = dense image of a column-major am x ak panel
            bbuf = torch.full((11, bn, bk), float("nan"), dtype=torch.float64, device="cuda")
            a_p = [gen.matrix("d", am, ak, 700 + s, "frac", "c") for s in range(npan)]
            b_p = [gen.matrix("d", bk, bn, 800 + s, "frac", "c") for s in range(npan)]
            for s in range(npan):
                abuf[slots_a[s]].t().copy_(to_torch(a_p[s])); bbuf[slots_b[s]].t().copy_(to_torch(b_p[s]))
            c = gen.matrix("d", m, n, 900, "frac", "c")
            want = c.copy(order="K")
            for s in range(npan):
                oracle.gemm(ta, tb, 2.0, a_p[s], b_p[s], be if s == 0 else 1.0, want)
            if be == 0.0:
                c[...] = np.nan
            tc = to_torch(c)
            api.bli_gemm_kpanels(torch.float64, ta, tb, m, n, k, 2.0, [abuf[q].t() for q in slots_a], 1, am, [bbuf[q].t() for q in slots_b], 1, bk,
                                 be, tc, *estr(c))
            torch.cuda.synchronize()
            kn = engine.last_kernel()
            assert kn.startswith("gemm_dmma_tma_kernel") and f"XK={int(not tb)},YK={int(bool(ta))}" in kn, kn
            assert ("CST=1" in kn) == (k * npan <= 256), (kn, k, npan)
            assert rel_err(to_numpy(tc), want) <= TOL["d"], (tr, m, n, k, npan, be, kn, rel_err(to_numpy(tc), want))
        # panels of unrelated allocations cannot be slots of one map: the cp.async kernel serves them (same result)
        ta_l = [torch.rand(64, 200, dtype=torch.float64, device="cuda") for _ in range(3)]      # column-major 200 x 64
        tb_l = [torch.rand(136, 64, dtype=torch.float64, device="cuda") for _ in range(3)]      # column-major 64 x 136
        ta_l[1] = torch.rand(64 * 200 + 1, dtype=torch.float64, device="cuda")[1:].view(64, 200)     # 8-byte offset: no 16-byte slot stride
        tcc = torch.zeros(136, 200, dtype=torch.float64, device="cuda")
        api.bli_gemm_kpanels(torch.float64, 0, 0, 200, 136, 64, 1.0, [x.t() for x in ta_l], 1, 200, [x.t() for x in tb_l], 1, 64, 0.0, tcc.t(), 1, 200)
        torch.cuda.synchronize()
        assert engine.last_kernel().startswith("gemm_dmma_ws_kernel"), engine.last_kernel()
        wantc = sum(b_ @ a_ for a_, b_ in zip(ta_l, tb_l))
        assert float((tcc - wantc).abs().max()) < 1e-11
    finally:
        engine.set_option("dgemm_cfg", -1)


def test_dgemm_ping_pong_small_k(engine, oracle):
    """The ping-pong dgemm kernel for small k (gemm_dmma_pp.cuh: the two q-halves of a tile, each with its own TMA ring,
    alternate on the tensor pipe, ordered by named barriers) against the oracle: all four staging orientations, ragged m/n/k (a CTA whose
    last tile is ragged, a problem with fewer tiles than CTAs, more tiles than CTAs), beta == 0 on NaN-poisoned C and
    beta != 0, row- and column-stored C; and bit-for-bit against the lockstep kernel (same k order per accumulator)."""
    engine.set_option("dgemm_cfg", 9); engine.set_option("dmma_pp", 1 << 20)
    seed = 7000
    try:
        for (m, n, k) in ((260, 132, 68), (128, 128, 16), (1540, 1412, 64), (2052, 3100, 132), (40, 24, 8)):
            for ta in (NO_TRANSPOSE, TRANSPOSE):
                for tb in (NO_TRANSPOSE, TRANSPOSE):
                    for oc, be in (("c", 1.2), ("r", 0.0), ("c", 0.0)):
                        seed += 1
                        am, ak = (k, m) if ta & TRANSPOSE else (m, k)
                        bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
                        a = gen.matrix("d", am, ak, seed, "frac", "c"); b = gen.matrix("d", bk, bn, seed + 5000, "frac", "c")
                        c = gen.matrix("d", m, n, seed + 9000, "frac", oc)
                        want = c.copy(order="K")
                        oracle.gemm(ta, tb, 2.0, a, b, be, want)
                        c0 = c.copy(order="K")
                        if be == 0.0:
                            c[...] = np.nan
                        got = run_gemm(engine, "d", ta, tb, 2.0, a, b, be, c)
                        kn = engine.last_kernel()
                        assert kn.startswith("gemm_dmma_pp_kernel"), kn
                        assert rel_err(got, want) <= TOL["d"], (m, n, k, ta, tb, oc, be, rel_err(got, want))
                        engine.set_option("dmma_pp", 0); engine.set_option("dmma_cst", 0)
                        try:
                            lock = run_gemm(engine, "d", ta, tb, 2.0, a, b, be, c0 if be != 0.0 else c)
                            assert engine.last_kernel().startswith("gemm_dmma_tma_kernel")
                        finally:
                            engine.set_option("dmma_pp", 1 << 20); engine.set_option("dmma_cst", 256)
                        assert np.array_equal(got, lock), (m, n, k, ta, tb, oc, be)
    finally:
        engine.set_option("dgemm_cfg", -1); engine.set_option("dmma_pp", 0); engine.set_option("dmma_cst", 256)


# Which option forces, and which kernel name proves, the TMA tensor-map kernel of a datatype.  Without forcing, the
# small-problem rules (gemm_d.cu: 4*t128 < 3*SMs, gemm_s.cu: 20*t128 < 11*SMs) send shapes of this size to the
# cp.async small-tile kernels, so the orientation / ragged-edge / CST logic of the TMA kernels would go untested.
TMA_FORCE = {"d": ("dgemm_cfg", 9, -1, "gemm_dmma_tma_kernel"), "s": ("sgemm_cfg", 3, -1, "gemm_ffma_tma_kernel"),
             "c": ("cgemm_cfg", 3, -1, "gemm_cfma_tma_kernel"), "z": ("zgemm_cfg", 2, 1, "gemm_zmma_tma_kernel")}


def _orientation(ta, tb, oc):
    """(XK, YK) of the kernel form for column-major A/B (host_gemm.cuh: column-stored C swaps the operands)."""
    a_kc, b_kc = bool(ta & TRANSPOSE), not (tb & TRANSPOSE)           # is k the contiguous index of op(A) rows / op(B) columns?
    return (int(b_kc), int(a_kc)) if oc == "c" else (int(a_kc), int(b_kc))


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_aligned_operands_tma_paths(engine, oracle, ch):
    """16-byte aligned operands on the TMA tensor-map kernels, FORCED through set_option and PROVEN by the name of the
    kernel that ran (b200_last_kernel): every trans/conj combination x C storage = all four (k-contiguous | p/q-contiguous)
    staging orientations, ragged tiles in m, n and k (TMA zero-fills out-of-bounds box elements), beta == 0 (C poisoned
    with NaN: must not be read) and beta != 0, k = 64 (CST: D staged through the ring) -- against the oracle."""
    cx = ch in "cz"
    trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
    al = (2.0 + 0.2j) if cx else 2.0
    key, forced, default, kname = TMA_FORCE[ch]
    seed, seen = 3000, set()
    engine.set_option(key, forced)
    try:
        for (m, n, k) in ((260, 132, 68), (128, 128, 32), (4, 8, 4), (516, 260, 100), (388, 516, 64), (132, 140, 300)):
            for ta in trs:
                for tb in trs:
                    for oc in "cr":
                        for be in ((1.2 + 0.5j) if cx else 1.2, 0.0):
                            seed += 1
                            am, ak = (k, m) if ta & TRANSPOSE else (m, k)
                            bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
                            a = gen.matrix(ch, am, ak, seed, "frac", "c"); b = gen.matrix(ch, bk, bn, seed + 5000, "frac", "c")
                            c = gen.matrix(ch, m, n, seed + 9000, "frac", oc)
                            want = c.copy(order="K")
                            oracle.gemm(ta, tb, al, a, b, be, want)
                            if be == 0.0:
                                c[...] = np.nan
                            got = run_gemm(engine, ch, ta, tb, al, a, b, be, c)
                            kn = engine.last_kernel()
                            # s/c: a k-contiguous Y of a large problem is transposed first; never at these sizes
                            assert kn.startswith(kname), (kn, ch, m, n, k, ta, tb, oc)
                            xk, yk = _orientation(ta, tb, oc)
                            assert f"XK={xk},YK={yk}" in kn, (kn, ta, tb, oc)
                            seen.add(kn)
                            assert rel_err(got, want) <= TOL[ch], (ch, m, n, k, ta, tb, oc, be, kn, rel_err(got, want))
    finally:
        engine.set_option(key, default)
    assert len({kn.split("TRI")[0] for kn in seen}) == 4, seen                     # all four orientations ran
    if ch in "dsc":
        assert any("CST=1" in kn for kn in seen) and any("CST=0" in kn for kn in seen), seen


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_tma_kernels_large_ragged_vs_reference(engine, ref, ch):
    """The DEFAULT dispatch at sizes above the small-tile thresholds: ragged m, n, k (1540 x 1412 x 132; k = 64; k = 1028
    just above the CST limit of 1024), every orientation, both C storages, beta == 0 and != 0 -- against the real reference
    library on the same inputs, with the kernel name asserted (the TMA kernels for s/d/c, the warp-specialised kernel for z)."""
    cx = ch in "cz"
    al = (2.0 + 0.2j) if cx else 2.0
    want_kernel = {"d": "gemm_dmma_tma_kernel", "s": "gemm_ffma_tma_kernel", "c": "gemm_cfma_tma_kernel", "z": "gemm_dmma_ws_kernel<double2"}[ch]
    seed, seen = 8000, set()
    for (m, n, k) in ((1540, 1412, 132), (1540, 1412, 64), (1412, 1540, 1028)):
        for ta in (NO_TRANSPOSE, CONJ_TRANSPOSE if cx else TRANSPOSE):
            for tb in (NO_TRANSPOSE, TRANSPOSE):
                for oc, be in (("c", (1.2 + 0.5j) if cx else 1.2), ("r", 0.0), ("r", 1.2)):
                    seed += 1
                    am, ak = (k, m) if ta & TRANSPOSE else (m, k)
                    bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
                    a = gen.matrix(ch, am, ak, seed, "frac", "c"); b = gen.matrix(ch, bk, bn, seed + 5000, "frac", "c")
                    c = gen.matrix(ch, m, n, seed + 9000, "frac", oc)
                    want = c.copy(order="K")
                    ref.gemm(ta, tb, al, a, b, be, want)
                    if be == 0.0:
                        c[...] = np.nan
                    got = run_gemm(engine, ch, ta, tb, al, a, b, be, c)
                    kn = engine.last_kernel()
                    assert kn.startswith(want_kernel), (kn, ch, m, n, k, ta, tb, oc)
                    seen.add(kn)
                    assert rel_err(got, want) <= TOL[ch] * 4, (ch, m, n, k, ta, tb, oc, be, kn, rel_err(got, want))
    if ch == "d":
        assert any("CST=1" in kn for kn in seen) and any("CST=0" in kn for kn in seen), seen


def test_dgemm_split_k_tail_mid_size(engine, ref):
    """Mid-size dgemm whose last wave of 128x128 tiles would leave SMs idle takes the split-k tail schedule of the TMA
    kernel (gemm_dmma_tma_kernel<...,SK=1>: tail tiles cut into k chunks, partial accumulators added in chunk order by the
    last unit).  Checked against the real reference library (elementwise), for bit-for-bit equality with the UNSPLIT kernel
    on power-of-two inputs (every summation order is exact there), for run-to-run reproducibility on ordinary inputs, and
    with ragged m, n, k, every staging orientation, both C storages, beta == 0 on a NaN-poisoned C."""
    seed, seen = 12000, set()
    cases = [(2048, 2048, 2048, NO_TRANSPOSE, NO_TRANSPOSE, "c", 1.2),       # 256 tiles on 148 SMs: 148 whole + 108 x 4 chunks
             (2040, 2000, 2100, TRANSPOSE, NO_TRANSPOSE, "r", 0.0),          # ragged in m, n and k
             (1412, 1540, 1028, NO_TRANSPOSE, TRANSPOSE, "c", 1.2),          # 156 tiles: 148 whole + 8 x 2
             (1664, 2304, 772, TRANSPOSE, TRANSPOSE, "c", 1.2)]              # 234 tiles: 148 whole + 86 x 3
    for (m, n, k, ta, tb, oc, be) in cases:
        seed += 1
        am, ak = (k, m) if ta & TRANSPOSE else (m, k)
        bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
        a = gen.matrix("d", am, ak, seed, "frac", "c"); b = gen.matrix("d", bk, bn, seed + 5000, "frac", "c")
        c = gen.matrix("d", m, n, seed + 9000, "frac", oc)
        want = c.copy(order="K")
        ref.gemm(ta, tb, 2.0, a, b, be, want)
        if be == 0.0:
            c[...] = np.nan
        got = run_gemm(engine, "d", ta, tb, 2.0, a, b, be, c)
        kn = engine.last_kernel()
        assert kn.startswith("gemm_dmma_tma_kernel") and "SK=1" in kn, (kn, m, n, k)
        seen.add(kn)
        assert rel_err(got, want) <= TOL["d"] * 4, (m, n, k, kn, rel_err(got, want))
        again = run_gemm(engine, "d", ta, tb, 2.0, a, b, be, c)
        assert np.array_equal(got, again), ("split-k result is not reproducible", m, n, k)
        # exact inputs: split and unsplit must agree bit for bit
        a2 = gen.matrix("d", am, ak, seed + 1, "pow2", "c"); b2 = gen.matrix("d", bk, bn, seed + 5001, "pow2", "c")
        c2 = gen.matrix("d", m, n, seed + 9001, "pow2", oc)
        split = run_gemm(engine, "d", ta, tb, 2.0, a2, b2, 1.0, c2)
        engine.set_option("dgemm_splitk", 0)
        try:
            whole = run_gemm(engine, "d", ta, tb, 2.0, a2, b2, 1.0, c2)
            kn0 = engine.last_kernel()
        finally:
            engine.set_option("dgemm_splitk", 1)
        assert "SK=1" not in kn0, kn0
        assert np.array_equal(split, whole), ("split-k differs from the unsplit kernel on exact inputs", m, n, k)
    assert len(seen) == 4, seen                                              # all four staging orientations


@pytest.mark.parametrize("ch", list("sdcz"))
@pytest.mark.parametrize("n", [4096, 16384])
def test_gemm_skinny_k64_baseline_shapes_testsuite_residual(engine, ch, n):
    """BASELINE configs[2] skinny shapes (n, n, 64), column-major, device resident: the reference testsuite's randomized
    residual with its pass thresholds (testsuite/src/test_gemm.c:393-401, :44-47); the kernel that serves the shape is
    recorded (d/s/c: the CST variant of the TMA kernels)."""
    if ch in "cz" and n == 16384:
        n = 8192                                           # 16384^2 complex128 x 3 copies + temporaries: keep the test light
    dev, k = "cuda", 64
    g = torch.Generator(device=dev); g.manual_seed(int(0xB200) + n)
    rdt = torch.float32 if ch in "sc" else torch.float64

    def rnd(rows, cols):
        x = torch.rand(cols, rows, dtype=rdt, device=dev, generator=g) * 2 - 1
        if ch in "cz":
            x = torch.complex(x, torch.rand(cols, rows, dtype=rdt, device=dev, generator=g) * 2 - 1)
        nrm = float(x.abs().sum(dim=1).max())
        return (x / float(2 ** np.ceil(np.log2(nrm)))).t()           # column-major rows x cols

    a, b, c = rnd(n, k), rnd(k, n), rnd(n, n)
    c0 = c.clone()
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if ch in "cz" else (2.0, 1.2))
    getattr(engine, GEMM[ch])(0, 0, n, n, k, al, a, 1, n, b, 1, k, be, c, 1, n)
    torch.cuda.synchronize()
    kn = engine.last_kernel()
    if ch in "dsc":
        assert "tma_kernel" in kn and "CST=1" in kn, kn
    resid = _testsuite_resid(al, a, b, be, c0, c)
    thresh = 1e-5 if ch in "sc" else 1e-14
    assert resid <= thresh, (ch, n, kn, resid)
    # beta == 0 must not read C (NaN poisoned) and must equal the beta != 0 result on a zero C
    c1 = torch.full((n, n), float("nan"), dtype=c.dtype, device=dev).t()
    c2 = torch.zeros(n, n, dtype=c.dtype, device=dev).t()
    getattr(engine, GEMM[ch])(0, 0, n, n, k, al, a, 1, n, b, 1, k, 0.0, c1, 1, n)
    getattr(engine, GEMM[ch])(0, 0, n, n, k, al, a, 1, n, b, 1, k, 1.0, c2, 1, n)
    torch.cuda.synchronize()
    assert bool(torch.equal(c1, c2)), (ch, n, "beta == 0 differs from accumulation into zeros")


@pytest.mark.parametrize("ch", list("sc"))
def test_gemm_fp32_k_contiguous_y_is_transposed_once(engine, ref, ch):
    """s/c with a k-contiguous second kernel operand (column-major "TN"/"TT", row-major "NT"...) and a problem large
    enough: the engine transposes that operand once into a q-contiguous temporary and runs the fast orientation.
    Same results (within TOL) with the transposition switched off, and against the real reference; ragged sizes."""
    cx = ch == "c"
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
    for idx, (m, n, k, ta, tb, oc) in enumerate(((1028, 1036, 1004, TRANSPOSE, NO_TRANSPOSE, "c"),
                                                 (1100, 900, 1200, CONJ_TRANSPOSE if cx else TRANSPOSE, TRANSPOSE, "c"),
                                                 (1024, 1024, 1024, NO_TRANSPOSE, NO_TRANSPOSE, "r"))):    # row-major C: Y = B
        am, ak = (k, m) if ta & TRANSPOSE else (m, k)
        bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
        a = gen.matrix(ch, am, ak, 4000 + idx, "frac"); b = gen.matrix(ch, bk, bn, 4100 + idx, "frac")
        c = gen.matrix(ch, m, n, 4200 + idx, "frac", oc)
        want = c.copy(order="K")
        ref.gemm(ta, tb, al, a, b, be, want)
        n0 = engine.launch_count()
        got = run_gemm(engine, ch, ta, tb, al, a, b, be, c)
        assert engine.launch_count() - n0 == 2, ("expected one transposition + one gemm kernel", idx)
        assert rel_err(got, want) <= TOL[ch] * 4, (ch, m, n, k, rel_err(got, want))
        engine.set_option("transpose_y", 0)
        try:
            n0 = engine.launch_count()
            got2 = run_gemm(engine, ch, ta, tb, al, a, b, be, c)
            assert engine.launch_count() - n0 == 1
        finally:
            engine.set_option("transpose_y", 1)
        assert rel_err(got2, want) <= TOL[ch] * 4


def test_zgemm_tma_kernel_opt_in(engine, oracle):
    """The TMA variant of the zgemm kernel (zgemm_cfg = 2; FLOAT64-typed tensor maps with two elements per complex number)
    against the oracle for every transposition/conjugation, ragged tiles in m, n, k, both C storages, and bit-for-bit
    against the default warp-specialised cp.async kernel (same fragment values, same accumulation order per accumulator
    up to the k permutation inside a stage, so only compared within tolerance), plus a triangular (herk) call."""
    trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE)
    al, be = 2.0 + 0.2j, 1.2 + 0.5j
    seed = 7000
    engine.set_option("zgemm_cfg", 2)
    try:
        for (m, n, k) in ((260, 132, 68), (64, 128, 8), (4, 8, 4), (516, 260, 100), (129, 200, 1)):
            for ta in trs:
                for tb in trs:
                    for oc in "cr":
                        seed += 1
                        am, ak = (k, m) if ta & TRANSPOSE else (m, k)
                        bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
                        a = gen.matrix("z", am, ak, seed, "frac", "c"); b = gen.matrix("z", bk, bn, seed + 5000, "frac", "c")
                        c = gen.matrix("z", m, n, seed + 9000, "frac", oc)
                        want = c.copy(order="K")
                        oracle.gemm(ta, tb, al, a, b, be, want)
                        got = run_gemm(engine, "z", ta, tb, al, a, b, be, c)
                        assert rel_err(got, want) <= TOL["z"], (m, n, k, ta, tb, oc, rel_err(got, want))
        a = gen.matrix("z", 300, 90, 1, "frac"); c = gen.matrix("z", 300, 300, 2, "frac")
        want = c.copy(order="K"); oracle.herk(0xC0, 0, 2.0, a, 1.2, want)
        ta_, tc_ = to_torch(a), to_torch(c)
        engine.bli_zherk(0xC0, 0, 300, 90, 2.0, ta_, *estr(a), 1.2, tc_, *estr(c)); torch.cuda.synchronize()
        assert rel_err(to_numpy(tc_), want) <= TOL["z"]
    finally:
        engine.set_option("zgemm_cfg", 1)


def test_gemm_concurrent_host_threads(engine):
    """BLIS is re-entrant (SURVEY 8b 'Threading'): several application threads call gemm at the same time,
    each on its own CUDA stream; the engine's shared state (tile-scheduler counters, workspace pool, staging
    ring) must not interfere."""
    import threading
    results = {}

    def work(tid):
        torch.cuda.set_device(0)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            g = torch.Generator(device="cuda"); g.manual_seed(100 + tid)
            n = 768 + 64 * tid
            a = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g).t()
            b = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g).t()
            ok = True
            for _ in range(20):
                c = torch.zeros(n, n, dtype=torch.float64, device="cuda").t()
                engine.bli_dgemm(0, 0, n, n, n, 1.0, a, 1, n, b, 1, n, 0.0, c, 1, n)
                s.synchronize()
                ok = ok and bool(torch.allclose(c, a @ b, rtol=1e-12, atol=1e-9))
            results[tid] = ok

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts: t.start()
    for t in ts: t.join()
    assert results == {0: True, 1: True, 2: True, 3: True}, results


def test_gemm_tile_counters_are_per_stream(engine):
    """A long kernel on one stream and more than 64 launches on another must not share a {tile, done} scheduler pair
    (round-1 advisor finding: the 64-slot ring wrapped).  Pairs are now owned by streams (context.cu: sched_slot)."""
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(5)
    nb, ns = 6144, 512
    a = torch.rand(nb, nb, dtype=torch.float64, device=dev, generator=g).t()
    b = torch.rand(nb, nb, dtype=torch.float64, device=dev, generator=g).t()
    sa = torch.rand(ns, ns, dtype=torch.float64, device=dev, generator=g).t()
    sb = torch.rand(ns, ns, dtype=torch.float64, device=dev, generator=g).t()
    want_big, want_small = a @ b, sa @ sb
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        c_big = torch.zeros(nb, nb, dtype=torch.float64, device=dev).t()
        c_small = [torch.zeros(ns, ns, dtype=torch.float64, device=dev).t() for _ in range(150)]
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            engine.bli_dgemm(0, 0, nb, nb, nb, 1.0, a, 1, nb, b, 1, nb, 0.0, c_big, 1, nb)          # ~13 ms
        with torch.cuda.stream(s2):
            for c in c_small:                                                                        # 150 launches meanwhile
                engine.bli_dgemm(0, 0, ns, ns, ns, 1.0, sa, 1, ns, sb, 1, ns, 0.0, c, 1, ns)
        torch.cuda.synchronize()
        assert bool(torch.allclose(c_big, want_big, rtol=1e-12, atol=1e-9))
        for c in c_small:
            assert bool(torch.allclose(c, want_small, rtol=1e-12, atol=1e-9))
