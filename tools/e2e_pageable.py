"""dgemm with PAGEABLE host operands (what a legacy BLAS caller passes) against pinned ones.  Dev tool.
usage: python -m tools.e2e_pageable [n]"""
import sys, time
import torch
from blis_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for pin in (True, False):
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
