"""dgemm n x n x k for small k on the TMA kernel: lockstep + D staged through the ring (CST) against ping-pong (PP).
Dev tool.  usage: python -m tools.skinny_sweep [n] [k,k,...]      prints one JSON line"""
import json
import sys

import torch

from blis_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [64, 128, 256, 512, 1024]
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


out = {"n": n, "hbm_peak_gbs": 6457.1}
c = rnd(n, n)
for k in ks:
    a, b = rnd(n, k), rnd(k, n)
    row = {}
    for beta in (1.2, 0.0):
        for name, (pp, cst) in {"cst": (0, 1024), "plain": (0, 0), "pp": (1 << 20, 1024)}.items():
            api.set_option("dmma_pp", pp); api.set_option("dmma_cst", cst)
            t = timeit(lambda: api.bli_dgemm(0, 0, n, n, k, 2.0, a, 1, n, b, 1, k, beta, c, 1, n))
            byts = 8.0 * (n * k * 2 + n * n * (2 if beta else 1))
            row[f"{name}_beta{beta}"] = {"tflops": round(2.0 * n * n * k / t / 1e12, 2), "gbs": round(byts / t / 1e9), "us": round(t * 1e6, 1),
                                         "kernel": api.last_kernel()}
    out[f"k{k}"] = row
    print(k, {kk: (v["tflops"], v["gbs"]) for kk, v in row.items()}, file=sys.stderr, flush=True)
api.set_option("dmma_pp", 0); api.set_option("dmma_cst", 256)
print(json.dumps(out))
