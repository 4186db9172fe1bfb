"""Sweep of dgemm/zgemm kernel configurations: quick correctness + perf (dev tool)."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blis_b200 import api
dev = torch.device("cuda:0")
OUT = {}

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

for cfg in (0, 3, 4, 5, 6):
    api.set_option("dgemm_cfg", cfg)
    r = {"max_abs_err": check(torch.float64, api.bli_dgemm)}
    for n in (2048, 4096, 8192, 16384):
        r[f"n{n}"] = perf(torch.float64, api.bli_dgemm, n)
    r["k64_16384"] = perf(torch.float64, api.bli_dgemm, 16384, 64)
    r["k512_8192"] = perf(torch.float64, api.bli_dgemm, 8192, 512)
    OUT[f"dgemm_cfg{cfg}"] = r
    print(cfg, json.dumps(r), flush=True)
for cfg in (0, 1):
    api.set_option("zgemm_cfg", cfg)
    r = {"max_abs_err": check(torch.complex128, api.bli_zgemm)}
    for n in (4096, 8192):
        r[f"n{n}"] = perf(torch.complex128, api.bli_zgemm, n)
    OUT[f"zgemm_cfg{cfg}"] = r
    print("z", cfg, json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(OUT, open("gpurun_out/probe2.json", "w"), indent=1)
