"""Perf of hemm / symm / trmm3 / trmm (dev tool).  usage: python -m tools.gpu_probe_l3b [m[,m...]] [chars]
Flop counts as the testsuite's (testsuite/src/test_libblis.c:3068-3117): hemm/symm 2 m^2 n, trmm/trmm3 m^2 n (left), x4 complex."""
import json
import os
import sys

import torch

from blis_b200 import api
from tools.gpu_probe2 import DT, rnd, timeit

LEFT, RIGHT, LOWER, UPPER = 0, 1, 0xC0, 0x60


def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8192,16384").split(",")]
    chars = sys.argv[2] if len(sys.argv) > 2 else "dszc"
    out = {}
    for ch in chars:
        dt = DT[ch]
        cm = 4 if dt.is_complex else 1
        for m in sizes:
            n = m
            a, b, c = rnd(m, m, dt), rnd(m, n, dt), rnd(m, n, dt)
            r = {}
            f = getattr(api, f"bli_{ch}symm")
            r["symm_L"] = cm * 2.0 * m * m * n / timeit(lambda: f(LEFT, LOWER, 0, 0, m, n, 2.0, a, 1, m, b, 1, m, 1.2, c, 1, m)) / 1e12
            f = getattr(api, f"bli_{ch}hemm")
            r["hemm_R"] = cm * 2.0 * m * m * n / timeit(lambda: f(RIGHT, UPPER, 0, 0, m, n, 2.0, a, 1, m, b, 1, m, 1.2, c, 1, m)) / 1e12
            f3 = getattr(api, f"bli_{ch}trmm3")
            r["trmm3_LL"] = cm * 1.0 * m * m * n / timeit(lambda: f3(LEFT, LOWER, 0, 0, 0, m, n, 2.0, a, 1, m, b, 1, m, 1.2, c, 1, m)) / 1e12
            r["trmm3_RU"] = cm * 1.0 * m * m * n / timeit(lambda: f3(RIGHT, UPPER, 0, 0, 0, m, n, 2.0, a, 1, m, b, 1, m, 1.2, c, 1, m)) / 1e12
            api.set_option("ktri_skip", 0)
            r["trmm3_LL_noskip"] = cm * 1.0 * m * m * n / timeit(lambda: f3(LEFT, LOWER, 0, 0, 0, m, n, 2.0, a, 1, m, b, 1, m, 1.2, c, 1, m)) / 1e12
            api.set_option("ktri_skip", 1)
            f4 = getattr(api, f"bli_{ch}trmm")
            r["trmm_LUT"] = cm * 1.0 * m * m * n / timeit(lambda: f4(LEFT, UPPER, 8, 0, m, n, 1.0, a, 1, m, b, 1, m)) / 1e12
            out[f"{ch}{m}"] = {k_: round(v, 2) for k_, v in r.items()}
            print(ch, m, json.dumps(out[f"{ch}{m}"]), flush=True)
            del a, b, c
            torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_l3b.json", "w"), indent=1)


if __name__ == "__main__":
    main()
