#!/usr/bin/env python3
"""Small problems through every production gemm/trsm kernel family, for compute-sanitizer (memcheck / racecheck / synccheck).

    compute-sanitizer --tool memcheck  python tools/sanitize_driver.py
    compute-sanitizer --tool racecheck python tools/sanitize_driver.py

Each case is checked against torch and the name of the kernel that ran is printed, so the sanitizer log shows which
kernels were covered: TMA (d/s/c/z), CST, warp-specialised cp.async, k-panel accumulation, triangular schedules, trsm.
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from blis_b200 import api  # noqa: E402

dev = "cuda"
torch.cuda.set_device(0)
g = torch.Generator(device=dev); g.manual_seed(1)


def rnd(m, n, dt):
    x = torch.rand(n, m, dtype=torch.float64, device=dev, generator=g) * 2 - 1
    if dt.is_complex:
        x = torch.complex(x, torch.rand(n, m, dtype=torch.float64, device=dev, generator=g) * 2 - 1)
    return x.to(dt).t()                                     # column-major m x n


def check(name, got, want, tol):
    err = float((got - want).abs().max() / max(1.0, float(want.abs().max())))
    print(f"{name:28s} {api.last_kernel():70s} err={err:.2e}", flush=True)
    assert err <= tol, (name, err)


GEMM = {torch.float64: api.bli_dgemm, torch.float32: api.bli_sgemm, torch.complex64: api.bli_cgemm, torch.complex128: api.bli_zgemm}
FORCE = {torch.float64: ("dgemm_cfg", 9, -1), torch.float32: ("sgemm_cfg", 3, -1), torch.complex64: ("cgemm_cfg", 3, -1),
         torch.complex128: ("zgemm_cfg", 2, 1)}

for dt, tol in ((torch.float64, 1e-12), (torch.float32, 5e-5), (torch.complex64, 5e-5), (torch.complex128, 1e-12)):
    m, n, k = 260, 388, 100
    a, b, c = rnd(m, k, dt), rnd(k, n, dt), rnd(m, n, dt)
    at = rnd(k, m, dt)
    key, forced, default = FORCE[dt]
    for label, opt in (("default", default), ("tma", forced)):
        api.set_option(key, opt)
        for beta in (1.2, 0.0):
            c1 = c.clone(memory_format=torch.preserve_format)
            GEMM[dt](0, 0, m, n, k, 2.0, a, 1, m, b, 1, k, beta, c1, 1, m); torch.cuda.synchronize()
            check(f"gemm NN {label} beta={beta}", c1, beta * c + 2.0 * (a @ b), tol)
        c1 = c.clone(memory_format=torch.preserve_format)
        GEMM[dt](8, 0, m, n, k, 2.0, at, 1, k, b, 1, k, 1.2, c1, 1, m); torch.cuda.synchronize()
        check(f"gemm TN {label}", c1, 1.2 * c + 2.0 * (at.t() @ b), tol)
        # unaligned view (offset by one element): cp.async kernels
        big = rnd(m + 1, k, dt)
        a_un = big[1:, :]
        c1 = c.clone(memory_format=torch.preserve_format)
        GEMM[dt](0, 0, m, n, k, 2.0, a_un, 1, m + 1, b, 1, k, 1.2, c1, 1, m); torch.cuda.synchronize()
        check(f"gemm unaligned {label}", c1, 1.2 * c + 2.0 * (a_un @ b), tol)
    api.set_option(key, default)

# k-panel accumulation (d): 3 panels in one launch
m, n, k = 260, 260, 64
ap = [rnd(m, k, torch.float64) for _ in range(3)]; bp = [rnd(k, n, torch.float64) for _ in range(3)]
c = rnd(m, n, torch.float64); c1 = c.clone(memory_format=torch.preserve_format)
api.bli_gemm_kpanels(torch.float64, 0, 0, m, n, k, 2.0, ap, 1, m, bp, 1, k, 1.2, c1, 1, m); torch.cuda.synchronize()
check("kpanels", c1, 1.2 * c + 2.0 * sum(x @ y for x, y in zip(ap, bp)), 1e-12)

# triangular schedules: syrk (TRI), trmm (ktri), and trsm (all variants of the solve kernels)
a = rnd(300, 90, torch.float64); c = rnd(300, 300, torch.float64); c1 = c.clone(memory_format=torch.preserve_format)
api.bli_dsyrk(0xC0, 0, 300, 90, 2.0, a, 1, 300, 1.2, c1, 1, 300); torch.cuda.synchronize()
check("syrk lower", torch.tril(c1), torch.tril(1.2 * c + 2.0 * (a @ a.t())), 1e-12)
for dt, tol in ((torch.float64, 1e-11), (torch.complex128, 1e-11), (torch.float32, 1e-3)):
    for mm, nn in ((300, 100), (600, 200)):
        t = rnd(mm, mm, dt) / 16; t.diagonal().add_(2.0)
        b = rnd(mm, nn, dt)
        for uplo, tri in ((0xC0, torch.tril), (0x60, torch.triu)):
            b1 = b.clone(memory_format=torch.preserve_format)
            {torch.float64: api.bli_dtrsm, torch.complex128: api.bli_ztrsm, torch.float32: api.bli_strsm}[dt](
                0, uplo, 0, 0, mm, nn, 2.0, t, 1, mm, b1, 1, mm)
            torch.cuda.synchronize()
            want = torch.linalg.solve_triangular(tri(t), 2.0 * b, upper=(uplo == 0x60))
            check(f"trsm {dt} {mm}x{nn} uplo={uplo:#x}", b1, want, tol)
b = rnd(300, 100, torch.float64); t = rnd(300, 300, torch.float64); b1 = b.clone(memory_format=torch.preserve_format)
api.bli_dtrmm(0, 0xC0, 0, 0, 300, 100, 2.0, t, 1, 300, b1, 1, 300); torch.cuda.synchronize()
check("trmm lower", b1, 2.0 * (torch.tril(t) @ b), 1e-12)

# ---- round 2 additions: split-k tail, fused trsm panel, grouped batch kernel, pipelined host trsm ---------------------
import os  # noqa: E402

# split-k tail of the TMA dgemm (156 tiles on 148 SMs -> 148 whole + 8 x 4 chunks); ragged m, n, k
m, n, k = 1412, 1540, 1028
a, b, c = rnd(m, k, torch.float64), rnd(k, n, torch.float64), rnd(m, n, torch.float64)
c1 = c.clone(memory_format=torch.preserve_format)
api.bli_dgemm(0, 0, m, n, k, 2.0, a, 1, m, b, 1, k, 1.2, c1, 1, m); torch.cuda.synchronize()
check("gemm split-k tail", c1, 1.2 * c + 2.0 * (a @ b), 1e-12)
assert "SK=1" in api.last_kernel(), api.last_kernel()

# fused diagonal-panel trsm kernel (m > 256 rows: panels of 256 + a ragged last one), lower and upper
for uplo, tri in ((0xC0, torch.tril), (0x60, torch.triu)):
    t = rnd(700, 700, torch.float64) / 16; t.diagonal().add_(2.0)
    b = rnd(700, 200, torch.float64); b1 = b.clone(memory_format=torch.preserve_format)
    api.bli_dtrsm(0, uplo, 0, 0, 700, 200, 2.0, t, 1, 700, b1, 1, 700); torch.cuda.synchronize()
    check(f"trsm fused panel uplo={uplo:#x}", b1, torch.linalg.solve_triangular(tri(t), 2.0 * b, upper=(uplo == 0x60)), 1e-11)
assert "trsm_panel_kernel" in " ".join(api.kernel_stats()), sorted(api.kernel_stats())

# grouped kernel: 60 small problems of three shapes in ONE launch (d and c)
for dt, tol in ((torch.float64, 1e-12), (torch.complex64, 5e-5)):
    groups, wants = [], []
    for (mm, nn, kk, ta) in ((33, 31, 17, 0), (64, 64, 64, 8), (5, 100, 40, 0)):
        aa = [rnd(*((kk, mm) if ta else (mm, kk)), dt) for _ in range(20)]; bb = [rnd(kk, nn, dt) for _ in range(20)]; cc = [rnd(mm, nn, dt) for _ in range(20)]
        wants += [1.2 * z + 2.0 * ((x.t() if ta else x) @ y) for x, y, z in zip(aa, bb, cc)]
        groups.append(dict(transa=ta, transb=0, m=mm, n=nn, k=kk, alpha=2.0, beta=1.2, a=aa, b=bb, c=cc))
    api.gemm_batch(dt, groups); torch.cuda.synchronize()
    got = [t for gr in groups for t in gr["c"]]
    err = max(float((x - y).abs().max()) for x, y in zip(got, wants))
    print(f"{'grouped batch ' + str(dt):28s} {api.last_kernel():70s} err={err:.2e}", flush=True)
    assert err <= tol * 10 and api.last_kernel().startswith("gemm_grouped_kernel")

# pinned host operands: B in row blocks, A in the order the solve reads it, X comes back block by block (trsm_host_rowpipe)
if not os.environ.get("SANITIZE_SKIP_HOST"):
    mm, nn = 4352, 2200
    t = rnd(mm, mm, torch.float64) / 64; t.diagonal().add_(2.0)
    b = rnd(mm, nn, torch.float64)
    th = torch.empty(mm, mm, dtype=torch.float64).pin_memory().t(); th.copy_(t)
    bh = torch.empty(nn, mm, dtype=torch.float64).pin_memory().t(); bh.copy_(b)
    api.bli_dtrsm(0, 0xC0, 0, 0, mm, nn, 2.0, th, 1, mm, bh, 1, mm)
    want = torch.linalg.solve_triangular(torch.tril(t), 2.0 * b, upper=False)
    check("trsm pinned host pipeline", bh.cuda(), want, 1e-10)
print("sanitize_driver: all cases ok; kernels:", sorted(api.kernel_stats()))
