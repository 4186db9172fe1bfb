"""Small and mid-size square gemm: what a call costs end to end from Python (queue of back-to-back calls) against what
the kernels alone take (the same calls captured in a CUDA graph and replayed: no host work between launches).
Dev tool.  usage: python -m tools.small_sweep [d|s|z|c] [n,n,...]      prints one JSON line"""
import json
import sys

import torch

from blis_b200 import api

ch = sys.argv[1] if len(sys.argv) > 1 else "d"
ns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256, 384, 512, 768, 1024, 1280, 1536]
dev = torch.device("cuda:0")
DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[ch]
FN = {"s": api.bli_sgemm, "d": api.bli_dgemm, "c": api.bli_cgemm, "z": api.bli_zgemm}[ch]
FL = 8.0 if ch in "cz" else 2.0
REPS = 50


def rnd(m, nn):
    if DT.is_complex:
        r = torch.float64 if DT == torch.complex128 else torch.float32
        return torch.view_as_complex(torch.empty(nn, m, 2, dtype=r, device=dev).uniform_(-1, 1)).t()
    return torch.empty(nn, m, dtype=DT, device=dev).uniform_(-1, 1).t()


def events(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


out = {}
for n in ns:
    a, b, c = rnd(n, n), rnd(n, n), rnd(n, n)

    def calls():
        for _ in range(REPS):
            FN(0, 0, n, n, n, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n)
    calls()
    t_loop = min(events(calls) for _ in range(3)) / REPS
    row = {"loop_us": round(t_loop * 1e6, 1), "loop_tflops": round(FL * n ** 3 / t_loop / 1e12, 2), "kernel": api.last_kernel()}
    try:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            calls()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                calls()
        torch.cuda.synchronize()
        g.replay()
        t_graph = min(events(g.replay) for _ in range(3)) / REPS
        row.update({"graph_us": round(t_graph * 1e6, 1), "graph_tflops": round(FL * n ** 3 / t_graph / 1e12, 2)})
    except Exception as exc:                                     # capture is an experiment: report, do not fail
        row["graph_error"] = repr(exc)[:200]
        torch.cuda.synchronize()
    out[str(n)] = row
    print(n, row, file=sys.stderr, flush=True)
print(json.dumps(out))
