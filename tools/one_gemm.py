"""Run a few launches of one gemm configuration (for ncu captures). Dev tool.
usage: python -m tools.one_gemm <s|d|c|z> <n> <k> <cfg> [reps]"""
import sys, torch
from blis_b200 import api
ch, n, k, cfg = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}[ch]
fn = {"s": api.bli_sgemm, "d": api.bli_dgemm, "c": api.bli_cgemm, "z": api.bli_zgemm}[ch]
api.set_option(ch + "gemm_cfg", cfg)
dev = torch.device("cuda:0")
def rnd(m, nn):
    if dt.is_complex:
        r = torch.float64 if dt == torch.complex128 else torch.float32
        return torch.view_as_complex(torch.empty(nn, m, 2, dtype=r, device=dev).uniform_(-1, 1)).t()
    return torch.empty(nn, m, dtype=dt, device=dev).uniform_(-1, 1).t()
a, b, c = rnd(n, k), rnd(k, n), rnd(n, n)
for _ in range(reps):
    fn(0, 0, n, n, k, 2.0, a, 1, n, b, 1, k, 1.2, c, 1, n)
torch.cuda.synchronize()
print("done")
