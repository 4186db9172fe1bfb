"""Diagnostic: where do trmm3 results with / without k-range skipping differ? (dev tool)"""
import sys
import torch
from blis_b200 import api
from tools.gpu_probe2 import DT, rnd

LEFT, RIGHT, LOWER, UPPER = 0, 1, 0xC0, 0x60
ch = sys.argv[1] if len(sys.argv) > 1 else "s"
m, n = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8192, 2048)
dt = DT[ch]
torch.manual_seed(1)
for side, uplo in ((LEFT, LOWER), (RIGHT, UPPER), (LEFT, UPPER)):
    ma = m if side == LEFT else n
    a, b, c0 = rnd(ma, ma, dt) / 32, rnd(m, n, dt) / 32, rnd(m, n, dt) / 32
    a = a.t().contiguous().t(); b = b.t().contiguous().t(); c0 = c0.t().contiguous().t()
    f3 = getattr(api, f"bli_{ch}trmm3")
    res = []
    for skip in (1, 0, 1, 0):
        api.set_option("ktri_skip", skip)
        c = c0.clone()
        f3(side, uplo, 0, 0, 0, m, n, 2.0, a, 1, ma, b, 1, m, 1.2, c, 1, m)
        torch.cuda.synchronize()
        res.append(c)
    api.set_option("ktri_skip", 1)
    a_tri = (torch.tril(a) if uplo == LOWER else torch.triu(a)).t().contiguous().t()
    full = c0.clone()
    g = getattr(api, f"bli_{ch}gemm")
    if side == LEFT: g(0, 0, m, n, m, 2.0, a_tri, 1, ma, b, 1, m, 1.2, full, 1, m)
    else: g(0, 0, m, n, n, 2.0, b, 1, m, a_tri, 1, ma, 1.2, full, 1, m)
    torch.cuda.synchronize()
    def cmp(x, y, name):
        d = (x - y).abs()
        bad = d > 0
        nb = int(bad.sum())
        msg = f"{name}: differing {nb}, max {float(d.max()):.3e}, nan {int(torch.isnan(x.abs()).sum())}/{int(torch.isnan(y.abs()).sum())}"
        if nb:
            idx = bad.nonzero()
            rows, cols = idx[:, 0], idx[:, 1]
            msg += f" rows {int(rows.min())}-{int(rows.max())} cols {int(cols.min())}-{int(cols.max())}; rows%128 hist {torch.bincount(rows % 128, minlength=128)[:8].tolist()} cols%128 hist {torch.bincount(cols % 128, minlength=128)[:8].tolist()}"
            msg += f" first {idx[:5].tolist()}"
        print(side, hex(uplo), msg, flush=True)
    cmp(res[0], res[2], "skip vs skip")
    cmp(res[1], res[3], "noskip vs noskip")
    cmp(res[0], res[1], "skip vs noskip")
    cmp(res[1], full, "noskip vs gemm(zero-filled)")
    cmp(res[0], full, "skip vs gemm(zero-filled)")
