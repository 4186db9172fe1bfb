"""Perf of the level-3 operations next to gemm/trsm (dev tool): syrk/herk/syr2k/her2k/gemmt at square sizes.

usage: python -m tools.gpu_probe_l3 [n[,n...]] [chars]      flop counts as testsuite/src/test_libblis.c:3068-3117
"""
import json
import os
import sys

import torch

from blis_b200 import api
from tools.gpu_probe2 import DT, rnd, timeit

LOWER, UPPER = 0xC0, 0x60


def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,8192,16384").split(",")]
    chars = sys.argv[2] if len(sys.argv) > 2 else "dszc"
    out = {}
    for ch in chars:
        dt = DT[ch]
        cm = 4 if dt.is_complex else 1
        for n in sizes:
            k = n
            a, b, bt, c = rnd(n, k, dt), rnd(n, k, dt), rnd(k, n, dt), rnd(n, n, dt)
            r = {}
            g = getattr(api, f"bli_{ch}gemm")
            r["gemm"] = cm * 2.0 * n * n * k / timeit(lambda: g(0, 0, n, n, k, 2.0, a, 1, n, bt, 1, k, 1.2, c, 1, n)) / 1e12
            f = getattr(api, f"bli_{ch}gemmt")
            r["gemmt"] = cm * 1.0 * n * n * k / timeit(lambda: f(LOWER, 0, 0, n, k, 2.0, a, 1, n, bt, 1, k, 1.2, c, 1, n)) / 1e12
            f1 = getattr(api, f"bli_{ch}syrk")
            r["syrk"] = cm * 1.0 * n * n * k / timeit(lambda: f1(LOWER, 0, n, k, 2.0, a, 1, n, 1.2, c, 1, n)) / 1e12
            r["syrk_uT"] = cm * 1.0 * n * n * k / timeit(lambda: f1(UPPER, 8, n, k, 2.0, a, 1, n, 1.2, c, 1, n)) / 1e12
            f2 = getattr(api, f"bli_{ch}herk")
            r["herk"] = cm * 1.0 * n * n * k / timeit(lambda: f2(LOWER, 0, n, k, 2.0, a, 1, n, 1.2, c, 1, n)) / 1e12
            f3 = getattr(api, f"bli_{ch}syr2k")
            r["syr2k"] = cm * 2.0 * n * n * k / timeit(lambda: f3(LOWER, 0, 0, n, k, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n)) / 1e12
            f4 = getattr(api, f"bli_{ch}her2k")
            r["her2k"] = cm * 2.0 * n * n * k / timeit(lambda: f4(UPPER, 0, 0, n, k, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n)) / 1e12
            out[f"{ch}{n}"] = {k_: round(v, 2) for k_, v in r.items()}
            print(ch, n, json.dumps(out[f"{ch}{n}"]), flush=True)
            del a, b, bt, c
            torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_l3.json", "w"), indent=1)


if __name__ == "__main__":
    main()
