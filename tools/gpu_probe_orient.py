"""Perf per operand orientation (dev tool): column-major NN / NT / TN / TT for each datatype.
usage: python -m tools.gpu_probe_orient [n] [chars]"""
import json
import os
import sys

import torch

from blis_b200 import api
from tools.gpu_probe2 import DT, FN, rnd, timeit


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    chars = sys.argv[2] if len(sys.argv) > 2 else "sdcz"
    out = {}
    for ch in chars:
        dt = DT[ch]; fn = FN[ch]
        cm = 4 if dt.is_complex else 1
        a, b, c = rnd(n, n, dt), rnd(n, n, dt), rnd(n, n, dt)
        r = {}
        for name, ta, tb in (("NN", 0, 0), ("NT", 0, 8), ("TN", 8, 0), ("TT", 8, 8)):
            t = timeit(lambda: fn(ta, tb, n, n, n, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n))
            r[name] = round(cm * 2.0 * n ** 3 / t / 1e12, 2)
        if ch in "sc":
            api.set_option("transpose_y", 0)
            for name, ta, tb in (("TN_notr", 8, 0), ("TT_notr", 8, 8)):
                t = timeit(lambda: fn(ta, tb, n, n, n, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n))
                r[name] = round(cm * 2.0 * n ** 3 / t / 1e12, 2)
            api.set_option("transpose_y", 1)
        out[ch] = r
        print(ch, n, json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_orient.json", "w"), indent=1)


if __name__ == "__main__":
    main()
