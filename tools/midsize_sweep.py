"""dgemm n^3 for mid-size n: the split-k tail schedule (dgemm_splitk 1) against whole tiles only (0).
Dev tool.  usage: python -m tools.midsize_sweep [n,n,...] [dgemm_cfg]      prints one JSON line"""
import json
import sys

import torch

from blis_b200 import api

ns = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1536, 1792, 2048, 2304, 2560, 3072, 3584, 4096]
if len(sys.argv) > 2:
    api.set_option("dgemm_cfg", int(sys.argv[2]))          # e.g. 9: force the 128x128 TMA kernel below its usual size range
dev = torch.device("cuda:0")


def rnd(m, nn):
    return torch.empty(nn, m, dtype=torch.float64, device=dev).uniform_(-1, 1).t()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


out = {}
for n in ns:
    a, b, c = rnd(n, n), rnd(n, n), rnd(n, n)
    row = {}
    for sk in (0, 1):
        api.set_option("dgemm_splitk", sk)
        t = timeit(lambda: api.bli_dgemm(0, 0, n, n, n, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n))
        row[f"splitk{sk}"] = {"tflops": round(2.0 * n ** 3 / t / 1e12, 2), "us": round(t * 1e6, 1), "kernel": api.last_kernel()}
    out[str(n)] = row
    print(n, {k: v["tflops"] for k, v in row.items()}, row["splitk1"]["kernel"][-12:], file=sys.stderr, flush=True)
api.set_option("dgemm_splitk", 1)
print(json.dumps(out))
