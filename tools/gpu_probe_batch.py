"""Batched small gemm: a loop of engine calls versus b200_gemm_batch (dev tool).
usage: python -m tools.gpu_probe_batch [count] [sizes]"""
import json, os, sys
import torch
from blis_b200 import api
from tools.gpu_probe2 import timeit

count = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16,32,64,128,256").split(",")]
dev = "cuda"
out = {}
for ch, dt, fn in (("d", torch.float64, api.bli_dgemm), ("s", torch.float32, api.bli_sgemm)):
    for n in sizes:
        a = [torch.randn(n, n, dtype=dt, device=dev).t() for _ in range(count)]
        b = [torch.randn(n, n, dtype=dt, device=dev).t() for _ in range(count)]
        c = [torch.randn(n, n, dtype=dt, device=dev).t() for _ in range(count)]
        def loop():
            for x, y, z in zip(a, b, c):
                fn(0, 0, n, n, n, 2.0, x, 1, n, y, 1, n, 1.2, z, 1, n)
        g = [dict(transa=0, transb=0, m=n, n=n, k=n, alpha=2.0, beta=1.2, a=a, b=b, c=c)]
        t_loop = timeit(loop)
        api.set_option("batch_grouped", 0)
        t_pool = timeit(lambda: api.gemm_batch(dt, g))               # one launch per problem on the stream pool
        api.set_option("batch_grouped", 1)
        t_batch = timeit(lambda: api.gemm_batch(dt, g))              # small problems: ONE launch of the grouped kernel
        flop = 2.0 * n ** 3 * count
        out[f"{ch}{n}"] = {"loop_TF": round(flop / t_loop / 1e12, 3), "pool_TF": round(flop / t_pool / 1e12, 3), "batch_TF": round(flop / t_batch / 1e12, 3),
                           "speedup": round(t_loop / t_batch, 2), "kernel": api.last_kernel()}
        print(ch, n, count, json.dumps(out[f"{ch}{n}"]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_batch.json", "w"), indent=1)
