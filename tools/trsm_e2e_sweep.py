"""dtrsm with pinned host operands end to end: the column-block pipeline (trsm_host_rb = -1) against the row-block
pipeline at several block sizes.  Dev tool.  usage: python -m tools.trsm_e2e_sweep [m n [m n ...]]
Prints one JSON line per shape: wall-clock ms of the whole call (best of 3 after one warm-up), TFLOP/s = m^2 n / t."""
import json
import sys
import time

import torch

from blis_b200 import api

shapes = [(32768, 8192), (16384, 8192), (8192, 8192), (16384, 2048)]
if len(sys.argv) > 2:
    v = [int(x) for x in sys.argv[1:]]
    shapes = list(zip(v[0::2], v[1::2]))
for m, n in shapes:
    a = torch.empty(m, m, dtype=torch.float64).pin_memory().t()          # column-major views of pinned memory
    b0 = torch.empty(n, m, dtype=torch.float64).pin_memory().t()
    b = torch.empty(n, m, dtype=torch.float64).pin_memory().t()
    a.uniform_(-1, 1); a.mul_(2.0 / m ** 0.5); a.diagonal().add_(2.0)
    b0.uniform_(-1, 1)
    out = {"m": m, "n": n, "ms": {}, "tflops": {}}
    ref = None
    variants = [("column_blocks", -1), ("auto", 0)] + [(f"rb{m // d}", m // d) for d in (64, 32, 16, 8) if m // d >= 256]
    for key, rb in variants:
        if rb > 0 and rb % 256:
            continue
        api.set_option("trsm_host_rb", rb)
        best = 1e30
        for rep in range(3):
            b.copy_(b0)
            t0 = time.perf_counter()
            api.bli_dtrsm(0, 0xC0, 0, 0, m, n, 2.0, a, 1, m, b, 1, m)
            dt = 1e3 * (time.perf_counter() - t0)
            if rep:
                best = min(best, dt)
        out["ms"][key] = round(best, 2); out["tflops"][key] = round(float(m) * m * n / best / 1e9, 2)
        if ref is None:
            ref = b.clone()
        else:
            out["max_abs_diff_vs_column_blocks"] = max(out.get("max_abs_diff_vs_column_blocks", 0.0), float((b - ref).abs().max()))
    api.set_option("trsm_host_rb", 0)
    print(json.dumps(out), flush=True)
    del a, b, b0, ref
