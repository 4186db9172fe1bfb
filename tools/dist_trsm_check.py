"""DistTrsm.step_host (the rank's block of B in pinned host memory, column sub-blocks pipelined) against
DistTrsm.step (device-resident) on one GPU; prints one JSON line.  python -m tools.dist_trsm_check [m n]"""
import json
import sys
import time

import torch

from blis_b200 import dist as bdist

m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 2048)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
job = bdist.DistTrsm(m, n, 1, 0, dev)
job.step(); torch.cuda.synchronize()
want = job.b.clone()
out = {"m": m, "n": n}
h = job.host_block()
for rep in range(3):
    h.copy_(job.b0); job.b.fill_(float("nan")); torch.cuda.synchronize()
    t0 = time.perf_counter(); job.step_host(h, nblk=4); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[f"rep{rep}_maxdiff"] = float((h - want.cpu()).abs().max())
    out[f"rep{rep}_ms"] = dt * 1e3
t0 = time.perf_counter(); job.step(); torch.cuda.synchronize(); out["device_resident_ms"] = (time.perf_counter() - t0) * 1e3
# the same transfer without the pipeline: upload, solve, download
h.copy_(job.b0); torch.cuda.synchronize()
t0 = time.perf_counter(); job.b.copy_(h, non_blocking=True); job._solve_cols(0, job.n_loc); h.copy_(job.b, non_blocking=True)
torch.cuda.synchronize(); out["sequential_ms"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(out))
