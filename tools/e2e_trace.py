"""Timeline of one pipelined host-operand dgemm (host_trace = 1).  Dev tool.  usage: python -m tools.e2e_trace [n]"""
import sys, time
import torch
from blis_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
a, b, c = (torch.empty(n, n, dtype=torch.float64).pin_memory().t() for _ in range(3))
for t in (a, b, c):
    t.uniform_(-1, 1)
for rep in range(3):
    api.set_option("host_trace", 1 if rep == 2 else 0)
    t0 = time.perf_counter()
    api.bli_dgemm(0, 0, n, n, n, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n)
    print(f"call {rep}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
