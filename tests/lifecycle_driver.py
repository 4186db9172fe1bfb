"""Run by tests/test_lifecycle_gpu.py in a fresh process: the engine's init/finalize cycle
(b200_init / b200_finalize stand in for bli_init / bli_finalize, frame/base/bli_init.c:87-99, which the
reference allows to be called repeatedly).  Prints one JSON line."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import gen                                    # noqa: E402
from refblis import Oracle                    # noqa: E402
from util import estr, rel_err, to_numpy, to_torch   # noqa: E402
from blis_b200 import _lib, api               # noqa: E402

orc, lib = Oracle(), _lib.load()
m, n, k = 200, 150, 70
a, b, c = gen.matrix("d", m, k, 1), gen.matrix("d", k, n, 2), gen.matrix("d", m, n, 3)
want = c.copy(order="K"); orc.gemm(0, 0, 2.0, a, b, 1.2, want)


def dgemm_device():
    ta, tb, tc = to_torch(a), to_torch(b), to_torch(c)
    api.bli_dgemm(0, 0, m, n, k, 2.0, ta, *estr(a), tb, *estr(b), 1.2, tc, *estr(c)); torch.cuda.synchronize()
    return to_numpy(tc)


def dgemm_host():                               # pageable host operands: pinned staging ring
    ch = torch.from_numpy(c.copy(order="K"))
    api.bli_dgemm(0, 0, m, n, k, 2.0, torch.from_numpy(a), *estr(a), torch.from_numpy(b), *estr(b), 1.2, ch, *estr(c))
    return ch.numpy()


def batch():
    g = dict(transa=0, transb=0, m=m, n=n, k=k, alpha=2.0, beta=1.2, a=[to_torch(a) for _ in range(3)],
             b=[to_torch(b) for _ in range(3)], c=[to_torch(c) for _ in range(3)])
    api.gemm_batch(torch.float64, [g]); torch.cuda.synchronize()
    return [to_numpy(x) for x in g["c"]]


out = {}
first = dgemm_device()
out["before"] = rel_err(first, want)
out["host_before"] = rel_err(dgemm_host(), want)
for cycle in range(3):
    lib.b200_finalize()
    lib.b200_finalize()                         # a second finalize is a no-op
    got = dgemm_device()                        # lazy re-initialisation
    out[f"same_bits_{cycle}"] = bool(np.array_equal(got, first))
    out[f"host_{cycle}"] = rel_err(dgemm_host(), want)
    out[f"batch_{cycle}"] = max(rel_err(x, want) for x in batch())
out["launches"] = api.launch_count()
print(json.dumps(out))
