"""Batched gemm (b200_gemm_batch) against the oracle, problem by problem: the reference's
?gemm_batch_ (frame/compat/extra/bla_gemm_batch.c) is a loop of bli_?gemm calls, so the per-problem checker is the gemm
oracle.  Device problems run concurrently on the engine's stream pool and must be ordered after earlier work on the
caller's stream and before later work."""
import numpy as np
import pytest
import torch

import gen
from refblis import CONJ_TRANSPOSE, NO_TRANSPOSE, TRANSPOSE
from util import NP2T, TOL, rel_err, to_numpy, to_torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_batch_groups_vs_oracle(engine, oracle, ch):
    cx = ch in "cz"
    dt = NP2T[np.dtype(gen.NP_DT[ch])]
    shapes = [(64, 48, 32, NO_TRANSPOSE, NO_TRANSPOSE, 5), (130, 7, 65, TRANSPOSE, NO_TRANSPOSE, 3), (1, 200, 9, NO_TRANSPOSE, CONJ_TRANSPOSE if cx else TRANSPOSE, 4),
              (257, 129, 40, NO_TRANSPOSE, NO_TRANSPOSE, 2), (16, 16, 0, NO_TRANSPOSE, NO_TRANSPOSE, 2),
              (300, 260, 200, TRANSPOSE, NO_TRANSPOSE, 3)]        # above the grouped kernel's size limit: stream pool + tiled kernels
    groups, wants, seed = [], [], 100
    for gi, (m, n, k, ta, tb, cnt) in enumerate(shapes):
        al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if cx else (2.0, 1.2))
        if gi == 1:
            be = 0.0
        g = dict(transa=ta, transb=tb, m=m, n=n, k=k, alpha=al, beta=be, a=[], b=[], c=[])
        for j in range(cnt):
            seed += 1
            am, ak = (k, m) if ta & TRANSPOSE else (m, k)
            bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
            a = gen.matrix(ch, max(am, 1), max(ak, 1), seed, "frac")[:am, :ak]; b = gen.matrix(ch, max(bk, 1), max(bn, 1), seed + 500, "frac")[:bk, :bn]
            c = gen.matrix(ch, m, n, seed + 900, "frac")
            want = c.copy(order="K")
            oracle.gemm(ta, tb, al, np.asfortranarray(a) if a.size else np.zeros((am, ak), a.dtype, order="F"),
                        np.asfortranarray(b) if b.size else np.zeros((bk, bn), b.dtype, order="F"), be, want)
            wants.append(want)
            g["a"].append(to_torch(np.asfortranarray(a)) if a.size else torch.zeros(am, ak, dtype=dt, device="cuda").t().contiguous().t())
            g["b"].append(to_torch(np.asfortranarray(b)) if b.size else torch.zeros(bk, bn, dtype=dt, device="cuda").t().contiguous().t())
            g["c"].append(to_torch(c))
        groups.append(g)
    # strides are per group: make every tensor of a group share the first one's strides
    for g in groups:
        for key in "abc":
            assert all(t.stride() == g[key][0].stride() for t in g[key])
    engine.gemm_batch(dt, groups)
    torch.cuda.synchronize()
    i = 0
    for g in groups:
        for t in g["c"]:
            assert rel_err(to_numpy(t), wants[i]) <= TOL[ch], (ch, g["m"], g["n"], g["k"], i)
            i += 1


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_batch_small_problems_share_one_launch(engine, oracle, ch):
    """Many small device-resident problems (the regime of test/test_gemm_batch.c) are served by ONE launch of the grouped
    kernel (gemm_grouped.cuh), proven by the launch counter and the kernel name; every problem against the oracle:
    all transposition / conjugation combinations, row- and column-stored C, ragged sizes below and above the 32 x 32 tile,
    beta == 0 on NaN-poisoned C, alpha == 0 with NaN in A (must not be read), k == 0.  With batch_grouped = 0 the same
    batch takes one launch per problem and must give the same bits (same k order per element) on exact inputs."""
    cx = ch in "cz"
    dt = NP2T[np.dtype(gen.NP_DT[ch])]
    trs = (NO_TRANSPOSE, TRANSPOSE, CONJ_TRANSPOSE) if cx else (NO_TRANSPOSE, TRANSPOSE)
    shapes = [(8, 8, 8), (33, 31, 17), (64, 64, 64), (100, 3, 70), (5, 127, 40), (96, 96, 128), (32, 32, 0), (48, 40, 33)]
    groups, wants, seed = [], [], 700
    for gi, (m, n, k) in enumerate(shapes):
        ta, tb = trs[gi % len(trs)], trs[(gi // 2) % len(trs)]
        al, be = ((2.0 + 1.0j, 0.5 - 0.25j) if cx else (2.0, 0.5))        # exact scalars: every rounding order gives the same bits
        if gi == 2:
            be = 0.0
        if gi == 3:
            al = 0.0
        oc = "r" if gi % 3 == 1 else "c"
        g = dict(transa=ta, transb=tb, m=m, n=n, k=k, alpha=al, beta=be, a=[], b=[], c=[])
        for j in range(40):
            seed += 1
            am, ak = (k, m) if ta & TRANSPOSE else (m, k)
            bk, bn = (n, k) if tb & TRANSPOSE else (k, n)
            a = gen.matrix(ch, max(am, 1), max(ak, 1), seed, "pow2")[:am, :ak]; b = gen.matrix(ch, max(bk, 1), max(bn, 1), seed + 500, "pow2")[:bk, :bn]
            c = gen.matrix(ch, m, n, seed + 900, "pow2", oc)
            want = c.copy(order="K")
            oracle.gemm(ta, tb, al, np.asfortranarray(a) if a.size else np.zeros((am, ak), a.dtype, order="F"),
                        np.asfortranarray(b) if b.size else np.zeros((bk, bn), b.dtype, order="F"), be, want)
            wants.append(want)
            if gi == 3:
                a = a.copy(order="F"); a[...] = np.nan
            if gi == 2:
                c[...] = np.nan
            g["a"].append(to_torch(np.asfortranarray(a)) if a.size else torch.zeros(am, ak, dtype=dt, device="cuda").t().contiguous().t())
            g["b"].append(to_torch(np.asfortranarray(b)) if b.size else torch.zeros(bk, bn, dtype=dt, device="cuda").t().contiguous().t())
            g["c"].append(to_torch(c))
        groups.append(g)
    c0 = [[t.clone() for t in g["c"]] for g in groups]
    torch.cuda.synchronize()
    n0 = engine.launch_count()
    engine.gemm_batch(dt, groups)
    torch.cuda.synchronize()
    assert engine.launch_count() - n0 == 1, engine.launch_count() - n0
    assert engine.last_kernel().startswith("gemm_grouped_kernel"), engine.last_kernel()
    got, i = [], 0
    for g in groups:
        for t in g["c"]:
            got.append(to_numpy(t))
            assert np.array_equal(got[-1], wants[i]), (ch, g["m"], g["n"], g["k"], i, rel_err(got[-1], wants[i]))   # exact inputs
            i += 1
    # the same batch without the grouped kernel: one launch (or more) per problem, same bits
    for g, cs in zip(groups, c0):
        g["c"] = cs
    engine.set_option("batch_grouped", 0)
    try:
        n0 = engine.launch_count()
        engine.gemm_batch(dt, groups)
        torch.cuda.synchronize()
        assert engine.launch_count() - n0 >= 40 * (len(shapes) - 1)
    finally:
        engine.set_option("batch_grouped", 1)
    i = 0
    for g in groups:
        for t in g["c"]:
            assert np.array_equal(to_numpy(t), got[i]), (ch, g["m"], g["n"], g["k"], i)
            i += 1


def test_gemm_batch_stream_ordering_and_host_operands(engine, oracle):
    """Work queued before the batch on the caller's stream is visible to it, work queued after it sees its results;
    a group with host operands takes the staged path."""
    dev = "cuda"
    n = 96
    a = [torch.randn(n, n, dtype=torch.float64, device=dev).t() for _ in range(12)]
    b = [torch.randn(n, n, dtype=torch.float64, device=dev).t() for _ in range(12)]
    c = [torch.zeros(n, n, dtype=torch.float64, device=dev).t() for _ in range(12)]
    for t in c:
        t.fill_(1.0)                                   # queued before the batch on torch's current stream
    g = dict(transa=0, transb=0, m=n, n=n, k=n, alpha=1.0, beta=2.0, a=a, b=b, c=c)
    engine.gemm_batch(torch.float64, [g])
    outs = [t * 1.0 for t in c]                        # queued after the batch
    torch.cuda.synchronize()
    for x, y, o in zip(a, b, outs):
        assert torch.allclose(o, 2.0 + x @ y, rtol=0, atol=1e-10)
    # host operands
    ah = gen.matrix("d", 70, 30, 1, "frac"); bh = gen.matrix("d", 30, 50, 2, "frac"); ch_ = gen.matrix("d", 70, 50, 3, "frac")
    want = ch_.copy(order="K"); oracle.gemm(0, 0, 2.0, ah, bh, 1.2, want)
    th = [torch.from_numpy(np.ascontiguousarray(x.T)).t() for x in (ah, bh, ch_)]
    engine.gemm_batch(torch.float64, [dict(transa=0, transb=0, m=70, n=50, k=30, alpha=2.0, beta=1.2, a=[th[0]], b=[th[1]], c=[th[2]])])
    assert rel_err(th[2].numpy(), want) <= TOL["d"]
