"""dgemm and dtrsm (T1 shape, m = 2n x n/2) with PAGEABLE host operands (what a legacy BLAS caller passes) against pinned
ones; dtrsm also with the transfers in sequence (trsm_host_pipe = 0).  Dev tool.  usage: python -m tools.e2e_pageable [n]"""
import sys, time
import torch
from blis_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for pin in (() if len(sys.argv) > 2 and sys.argv[2] == "trsm" else (True, False)):
    a, b, c = (torch.empty(n, n, dtype=torch.float64) for _ in range(3))
    if pin:
        a, b, c = a.pin_memory(), b.pin_memory(), c.pin_memory()
    a, b, c = a.t(), b.t(), c.t()
    for t in (a, b, c):
        t.uniform_(-1, 1)
    for rep in range(3):
        t0 = time.perf_counter()
        api.bli_dgemm(0, 0, n, n, n, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n)
        dt = time.perf_counter() - t0
        print(f"{'pinned' if pin else 'pageable'} call {rep}: {1e3 * dt:.1f} ms = {2.0 * n ** 3 / dt / 1e12:.2f} TFLOP/s", flush=True)

m, nr = 2 * n, n // 2
for pin in (True, False):
    a, b, b0 = torch.empty(m, m, dtype=torch.float64), torch.empty(nr, m, dtype=torch.float64), torch.empty(nr, m, dtype=torch.float64)
    if pin:
        a, b = a.pin_memory(), b.pin_memory()
    a, b, b0 = a.t(), b.t(), b0.t()
    a.uniform_(-1, 1); a.mul_(2.0 / m ** 0.5); a.diagonal().add_(2.0); b0.uniform_(-1, 1)
    # (pipelined?, block rows): the engine's choice, larger block rows (longer lines for the packing threads), sequential
    for pipe, rb in ((1, 0), (0, 0)) if pin else ((1, 0), (1, 2048), (1, 4096), (1, 8192), (0, 0)):
        api.set_option("trsm_host_pipe", pipe); api.set_option("trsm_host_rb", rb)
        for rep in range(2):
            b.copy_(b0)
            t0 = time.perf_counter()
            api.bli_dtrsm(0, 0xC0, 0, 0, m, nr, 2.0, a, 1, m, b, 1, m)
            dt = time.perf_counter() - t0
            print(f"dtrsm {m}x{nr} {'pinned' if pin else 'pageable'} {'pipelined' if pipe else 'sequential'} rb={rb or 'auto'} call {rep}: {1e3 * dt:.1f} ms = "
                  f"{float(m) * m * nr / dt / 1e12:.2f} TFLOP/s", flush=True)
    api.set_option("trsm_host_pipe", 1); api.set_option("trsm_host_rb", 0)
