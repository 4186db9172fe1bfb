"""Sweep of gemm kernel configurations: quick correctness + perf (dev tool).

usage: python -m tools.gpu_probe2 <s|d|c|z> cfg[,cfg...] shape[,shape...]   shape = n or nxk (m=n)
"""
import json
import os
import sys

import torch

from blis_b200 import api

dev = torch.device("cuda:0")
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
FN = {"s": api.bli_sgemm, "d": api.bli_dgemm, "c": api.bli_cgemm, "z": api.bli_zgemm}


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def rnd(m, n, dt):
    if dt.is_complex:
        r = torch.float64 if dt == torch.complex128 else torch.float32
        return torch.view_as_complex(torch.empty(n, m, 2, dtype=r, device=dev).uniform_(-1, 1)).t()
    return torch.empty(n, m, dtype=dt, device=dev).uniform_(-1, 1).t()


def check(dt, fn):
    worst = 0.0
    for (m, n, k) in ((256, 256, 256), (129, 131, 67), (300, 77, 513), (1, 333, 40), (515, 1, 64), (128, 128, 16), (7, 5, 3), (1000, 1000, 1000)):
        for ta in (0, 8):
            for tb in (0, 8):
                a = rnd(*((k, m) if ta else (m, k)), dt); b = rnd(*((n, k) if tb else (k, n)), dt); c = rnd(m, n, dt)
                ref = 1.2 * c + 2.0 * ((a.t() if ta else a) @ (b.t() if tb else b))
                fn(ta, tb, m, n, k, 2.0, a, a.stride(0), a.stride(1), b, b.stride(0), b.stride(1), 1.2, c, c.stride(0), c.stride(1))
                torch.cuda.synchronize()
                worst = max(worst, float((c - ref).abs().max()))
    return worst


def perf(dt, fn, n, k=None):
    k = k or n
    a, b, c = rnd(n, k, dt), rnd(k, n, dt), rnd(n, n, dt)
    t = timeit(lambda: fn(0, 0, n, n, k, 2.0, a, 1, n, b, 1, k, 1.2, c, 1, n))
    return (4 if dt.is_complex else 1) * 2.0 * n * n * k / t / 1e12


def main():
    ch = sys.argv[1]
    cfgs = [int(x) for x in sys.argv[2].split(",")]
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[3].split(",")]
    out = {}
    for cfg in cfgs:
        api.set_option(ch + "gemm_cfg", cfg)
        r = {"max_abs_err": check(DT[ch], FN[ch])}
        for sh in shapes:
            n, k = (sh[0], sh[0]) if len(sh) == 1 else sh
            r[f"{n}x{k}"] = round(perf(DT[ch], FN[ch], n, k), 2)
        out[f"{ch}gemm_cfg{cfg}"] = r
        print(ch, cfg, json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/probe2_{ch}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
