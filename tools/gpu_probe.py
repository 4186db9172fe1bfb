"""GPU probe: pipe peaks, cuBLAS reference points, correctness spot checks and a
tile-config sweep.  Development tool (run under gpurun); results land in
gpurun_out/probe.json.  Not part of the product path.
"""
from __future__ import annotations

import json
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blis_b200 import api as A  # noqa: E402
from blis_b200 import (BLIS_CONJ_NO_TRANSPOSE, BLIS_CONJ_TRANSPOSE, BLIS_LEFT, BLIS_LOWER,  # noqa: E402
                       BLIS_NO_TRANSPOSE, BLIS_NONUNIT_DIAG, BLIS_RIGHT, BLIS_TRANSPOSE,
                       BLIS_UNIT_DIAG, BLIS_UPPER)

OUT = {}
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def section(name):
    def deco(fn):
        t0 = time.time()
        try:
            OUT[name] = fn()
        except Exception as e:  # noqa: BLE001
            OUT[name] = {"error": repr(e), "tb": traceback.format_exc()[-1500:]}
        print(f"[{name}] {time.time()-t0:.1f}s -> {json.dumps(OUT[name])[:1500]}", flush=True)
        return fn
    return deco


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def rnd(m, n, dtype, order="c"):
    if order == "c":   # column-major: tensor of shape (m,n) with strides (1,m)
        t = torch.empty(n, m, dtype=dtype, device=dev).uniform_(-1, 1) if not dtype.is_complex else \
            torch.view_as_complex(torch.empty(n, m, 2, dtype=torch.float32 if dtype == torch.complex64 else torch.float64, device=dev).uniform_(-1, 1))
        return t.t()
    t = torch.empty(m, n, dtype=dtype, device=dev).uniform_(-1, 1) if not dtype.is_complex else \
        torch.view_as_complex(torch.empty(m, n, 2, dtype=torch.float32 if dtype == torch.complex64 else torch.float64, device=dev).uniform_(-1, 1))
    return t


GEMM = {torch.float32: A.bli_sgemm, torch.float64: A.bli_dgemm, torch.complex64: A.bli_cgemm, torch.complex128: A.bli_zgemm}
TRSM = {torch.float32: A.bli_strsm, torch.float64: A.bli_dtrsm, torch.complex64: A.bli_ctrsm, torch.complex128: A.bli_ztrsm}


def op(t, trans):
    if trans == BLIS_NO_TRANSPOSE: return t
    if trans == BLIS_TRANSPOSE: return t.t()
    if trans == BLIS_CONJ_NO_TRANSPOSE: return t.conj()
    return t.t().conj()


def gemm_check(dtype, m, n, k, ta, tb, oa, ob, oc, alpha, beta):
    am, ak = (m, k) if not (ta & BLIS_TRANSPOSE) else (k, m)
    bk, bn = (k, n) if not (tb & BLIS_TRANSPOSE) else (n, k)
    a, b, c = rnd(am, ak, dtype, oa), rnd(bk, bn, dtype, ob), rnd(m, n, dtype, oc)
    hi = torch.complex128 if dtype.is_complex else torch.float64
    ref = beta * c.to(hi) + alpha * (op(a, ta).to(hi) @ op(b, tb).to(hi))
    GEMM[dtype](ta, tb, m, n, k, alpha, a, a.stride(0), a.stride(1), b, b.stride(0), b.stride(1), beta, c, c.stride(0), c.stride(1))
    torch.cuda.synchronize()
    err = (c.to(hi) - ref).abs().max().item() / max(1.0, ref.abs().max().item())
    return err


@section("peaks")
def _():
    r = {}
    for kind in ("dfma", "dmma", "ffma"):
        r[kind + "_200ms"] = A.measure_peak(kind, 200)
        r[kind + "_2s"] = A.measure_peak(kind, 2000)
    return r


@section("gemm_correctness")
def _():
    res = {}
    worst = {}
    for dtype, tol in ((torch.float64, 1e-12), (torch.complex128, 1e-12), (torch.float32, 2e-4), (torch.complex64, 2e-4)):
        w = 0.0
        bad = []
        cases = [
            (256, 256, 256), (128, 128, 16), (1, 1, 1), (7, 5, 3), (129, 131, 67), (300, 77, 513), (64, 1000, 1), (1, 333, 40), (515, 1, 64), (255, 257, 33),
        ]
        for (m, n, k) in cases:
            for ta in (BLIS_NO_TRANSPOSE, BLIS_TRANSPOSE, BLIS_CONJ_NO_TRANSPOSE, BLIS_CONJ_TRANSPOSE):
                for tb in (BLIS_NO_TRANSPOSE, BLIS_TRANSPOSE, BLIS_CONJ_TRANSPOSE):
                    for (oa, ob, oc) in (("c", "c", "c"), ("r", "r", "r"), ("c", "r", "c"), ("r", "c", "r")):
                        for (al, be) in ((2.0, 1.2), (1.0, 0.0)):
                            if dtype.is_complex:
                                al, be = complex(al, 0.2), complex(be, 0.5 if be else 0.0)
                            e = gemm_check(dtype, m, n, k, ta, tb, oa, ob, oc, al, be)
                            w = max(w, e)
                            if not (e <= tol):
                                bad.append((m, n, k, ta, tb, oa, ob, oc, str(al), str(be), e))
        res[str(dtype)] = {"max_rel_err": w, "n_bad": len(bad), "bad": bad[:8]}
    return res


@section("gemm_beta0_nan")
def _():
    # beta == 0 must not read C (NaNs in C must not propagate)
    r = {}
    for dtype in (torch.float64, torch.float32, torch.complex128, torch.complex64):
        a, b = rnd(100, 50, dtype), rnd(50, 70, dtype)
        c = torch.full((70, 100), float("nan"), dtype=dtype, device=dev).t()
        GEMM[dtype](0, 0, 100, 70, 50, 1.0, a, 1, 100, b, 1, 50, 0.0, c, 1, 100)
        torch.cuda.synchronize()
        r[str(dtype)] = bool(torch.isfinite(torch.view_as_real(c) if dtype.is_complex else c).all().item())
    return r


@section("gemm_general_stride_and_host")
def _():
    r = {}
    # general stride views
    big = torch.empty(400, 600, dtype=torch.float64, device=dev).uniform_(-1, 1)
    a = big[::2, ::3][:150, :90]          # rs=1200, cs=3
    bb = torch.empty(300, 500, dtype=torch.float64, device=dev).uniform_(-1, 1)
    b = bb[::3, ::2][:90, :110]
    cc = torch.empty(700, 900, dtype=torch.float64, device=dev).uniform_(-1, 1)
    c = cc[::4, ::5][:150, :110]
    ref = 1.2 * c + 2.0 * (a @ b)
    A.bli_dgemm(0, 0, 150, 110, 90, 2.0, a, a.stride(0), a.stride(1), b, b.stride(0), b.stride(1), 1.2, c, c.stride(0), c.stride(1))
    torch.cuda.synchronize()
    r["general_stride_err"] = (c - ref).abs().max().item()
    # host operands (pageable and pinned), column-major with padding
    for pin in (False, True):
        ah = torch.empty(200, 300, dtype=torch.float64).uniform_(-1, 1)   # row-major host
        bh = torch.empty(70, 310, dtype=torch.float64).uniform_(-1, 1).t()[:300, :60]  # col-major, ld=310
        ch = torch.empty(60, 210, dtype=torch.float64).uniform_(-1, 1).t()[:200, :60]
        if pin:
            ah = ah.pin_memory()
        refh = 0.5 * ch + 1.5 * (ah @ bh)
        A.bli_dgemm(0, 0, 200, 60, 300, 1.5, ah, ah.stride(0), ah.stride(1), bh, bh.stride(0), bh.stride(1), 0.5, ch, ch.stride(0), ch.stride(1))
        r[f"host_pin{int(pin)}_err"] = (ch - refh).abs().max().item()
    return r


def trsm_check(dtype, side, uplo, trans, diag, m, n, order_a="c", order_b="c"):
    ma = m if side == BLIS_LEFT else n
    a = rnd(ma, ma, dtype, order_a)
    a = a + 0  # materialise with the same strides
    a.diagonal().add_(2.0 * ma ** 0.5)
    a.mul_(1.0 / (ma ** 0.5))
    b = rnd(m, n, dtype, order_b)
    hi = torch.complex128 if dtype.is_complex else torch.float64
    alpha = complex(2.0, 0.3) if dtype.is_complex else 2.0
    tri = torch.tril(a) if uplo == BLIS_LOWER else torch.triu(a)
    if diag == BLIS_UNIT_DIAG:
        tri = tri - torch.diag(torch.diagonal(tri)) + torch.eye(ma, dtype=dtype, device=dev)
    t = op(tri, trans).to(hi)
    b0 = b.to(hi).clone()
    TRSM[dtype](side, uplo, trans, diag, m, n, alpha, a, a.stride(0), a.stride(1), b, b.stride(0), b.stride(1))
    torch.cuda.synchronize()
    x = b.to(hi)
    resid = (t @ x - alpha * b0) if side == BLIS_LEFT else (x @ t - alpha * b0)
    return resid.abs().max().item() / max(1.0, (alpha * b0).abs().max().item())


@section("trsm_correctness")
def _():
    res = {}
    for dtype, tol in ((torch.float64, 1e-11), (torch.complex128, 1e-11), (torch.float32, 1e-3), (torch.complex64, 1e-3)):
        w, bad = 0.0, []
        for (m, n) in ((64, 64), (1, 1), (5, 9), (100, 37), (257, 130), (500, 64), (33, 700)):
            for side in (BLIS_LEFT, BLIS_RIGHT):
                for uplo in (BLIS_LOWER, BLIS_UPPER):
                    for trans in (BLIS_NO_TRANSPOSE, BLIS_TRANSPOSE, BLIS_CONJ_NO_TRANSPOSE, BLIS_CONJ_TRANSPOSE):
                        for diag in (BLIS_NONUNIT_DIAG, BLIS_UNIT_DIAG):
                            for (oa, ob) in (("c", "c"), ("r", "r"), ("c", "r")):
                                e = trsm_check(dtype, side, uplo, trans, diag, m, n, oa, ob)
                                w = max(w, e)
                                if not (e <= tol):
                                    bad.append((m, n, side, uplo, trans, diag, oa, ob, e))
        res[str(dtype)] = {"max_rel_resid": w, "n_bad": len(bad), "bad": bad[:8]}
    return res


@section("cublas_reference_points")
def _():
    r = {}
    for n in (8192, 16384):
        a, b = rnd(n, n, torch.float64), rnd(n, n, torch.float64)
        c = torch.empty(n, n, dtype=torch.float64, device=dev)
        t = timeit(lambda: torch.matmul(a, b, out=c), reps=3)
        r[f"cublas_dgemm_{n}_tflops"] = 2 * n ** 3 / t / 1e12
        del a, b, c
    n = 8192
    a, b = rnd(n, n, torch.float32), rnd(n, n, torch.float32)
    c = torch.empty(n, n, dtype=torch.float32, device=dev)
    t = timeit(lambda: torch.matmul(a, b, out=c), reps=3)
    r[f"cublas_sgemm_{n}_tflops"] = 2 * n ** 3 / t / 1e12
    return r


def gemm_perf(dtype, m, n, k, reps=3):
    a, b, c = rnd(m, k, dtype), rnd(k, n, dtype), rnd(m, n, dtype)
    f = lambda: GEMM[dtype](0, 0, m, n, k, 2.0, a, 1, m, b, 1, k, 1.2, c, 1, m)  # noqa: E731
    t = timeit(f, reps=reps)
    mul = 4 if dtype.is_complex else 1
    return mul * 2.0 * m * n * k / t / 1e12


@section("dgemm_cfg_sweep")
def _():
    r = {}
    for cfg in (0, 1, 2, 3):
        A.set_option("dgemm_cfg", cfg)
        for n in (4096, 8192):
            r[f"cfg{cfg}_n{n}"] = gemm_perf(torch.float64, n, n, n)
    best = max(range(4), key=lambda c: r[f"cfg{c}_n8192"])
    A.set_option("dgemm_cfg", best)
    r["best_cfg"] = best
    r[f"cfg{best}_n16384"] = gemm_perf(torch.float64, 16384, 16384, 16384)
    for gm in (2,):
        A.set_option("grid_mult", gm)
        r[f"cfg{best}_n8192_gridmult{gm}"] = gemm_perf(torch.float64, 8192, 8192, 8192)
    A.set_option("grid_mult", 1)
    return r


@section("other_gemm_perf")
def _():
    r = {}
    r["zgemm_4096"] = gemm_perf(torch.complex128, 4096, 4096, 4096)
    r["zgemm_8192"] = gemm_perf(torch.complex128, 8192, 8192, 8192)
    r["sgemm_8192"] = gemm_perf(torch.float32, 8192, 8192, 8192)
    r["sgemm_16384"] = gemm_perf(torch.float32, 16384, 16384, 16384)
    r["cgemm_8192"] = gemm_perf(torch.complex64, 8192, 8192, 8192)
    r["dgemm_skinny_16384_k64"] = gemm_perf(torch.float64, 16384, 16384, 64, reps=5)
    r["sgemm_skinny_16384_k64"] = gemm_perf(torch.float32, 16384, 16384, 64, reps=5)
    return r


@section("dtrsm_perf")
def _():
    r = {}
    for (m, n) in ((8192, 8192), (32768, 8192)):
        a = rnd(m, m, torch.float64)
        a.diagonal().add_(2.0 * m ** 0.5)
        b = rnd(m, n, torch.float64)
        f = lambda: A.bli_dtrsm(BLIS_LEFT, BLIS_LOWER, BLIS_NO_TRANSPOSE, BLIS_NONUNIT_DIAG, m, n, 1.0, a, 1, m, b, 1, m)  # noqa: E731
        t = timeit(f, reps=2)
        r[f"dtrsm_llnn_{m}x{n}_tflops"] = 1.0 * m * m * n / t / 1e12
        del a, b
    return r


os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe.json", "w") as f:
    json.dump(OUT, f, indent=1)
print("DONE")
