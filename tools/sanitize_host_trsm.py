#!/usr/bin/env python3
"""The host-operand trsm pipeline (csrc/host_trsm.cuh: trsm_host_rowpipe) alone, for compute-sanitizer memcheck: pinned and
pageable A/B, lower and upper, ragged block rows.  (tools/sanitize_driver.py covers the kernels; this covers the copies.)

    compute-sanitizer --tool memcheck python tools/sanitize_host_trsm.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from blis_b200 import api  # noqa: E402

torch.cuda.set_device(0)
g = torch.Generator(); g.manual_seed(3)
m, n = 4352 + 40, 1100
for pin in (True, False):
    for uplo, rb in ((0xC0, 0), (0x60, 768)):
        t = (torch.rand(m, m, dtype=torch.float64, generator=g) * 2 - 1) / 64
        t.diagonal().add_(2.0)
        b = torch.rand(m, n, dtype=torch.float64, generator=g) * 2 - 1
        th, bh = torch.empty(m, m, dtype=torch.float64), torch.empty(n, m, dtype=torch.float64)
        if pin:
            th, bh = th.pin_memory(), bh.pin_memory()
        th, bh = th.t(), bh.t()
        th.copy_(t); bh.copy_(b)
        api.set_option("trsm_host_rb", rb)
        api.bli_dtrsm(0, uplo, 0, 0, m, n, 2.0, th, 1, m, bh, 1, m)
        tri = torch.tril(t) if uplo == 0xC0 else torch.triu(t)
        want = torch.linalg.solve_triangular(tri.cuda(), 2.0 * b.cuda(), upper=(uplo == 0x60))
        err = float((bh.cuda() - want).abs().max())
        print(f"trsm host pipeline {'pinned' if pin else 'pageable'} {'lower' if uplo == 0xC0 else 'upper'} rb={rb or 'auto'}: err={err:.2e} "
              f"kernels={sorted(api.kernel_stats())}", flush=True)
        assert err < 1e-10
api.set_option("trsm_host_rb", 0)
print("sanitize_host_trsm: all cases ok")
