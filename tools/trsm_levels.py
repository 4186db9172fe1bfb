"""Where does dtrsm's time go?  Times the base solves and the rank-k updates per recursion level (dev tool)."""
import sys, torch
from blis_b200 import api
m, n, NB = 32768, int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 64
dev = torch.device("cuda:0")
a = (torch.rand(m, m, dtype=torch.float64, device=dev) * 2 - 1) / m
a.diagonal().add_(2.0); a = a.t()
b = (torch.rand(n, m, dtype=torch.float64, device=dev) * 2 - 1).t()
def ev(): return torch.cuda.Event(enable_timing=True)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = ev(), ev(); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
# full trsm
t_full = timed(lambda: api.bli_dtrsm(0, 0xC0, 0, 0, m, n, 1.0, a, 1, m, b, 1, m))
print(f"full dtrsm {m}x{n}: {t_full:.1f} ms = {m*m*n/t_full/1e9:.1f} TF")
# base solves only
def bases():
    for i0 in range(0, m, NB):
        api.bli_dtrsm(0, 0xC0, 0, 0, NB, n, 1.0, a[i0:i0+NB, i0:i0+NB], 1, m, b[i0:i0+NB], 1, m)
print(f"512 base solves: {timed(bases):.2f} ms")
s = NB
tot = 0.0
while s < m:
    def level():
        for i0 in range(0, m, 2 * s):
            api.bli_dgemm(0, 0, s, n, s, -1.0, a[i0+s:i0+2*s, i0:i0+s], 1, m, b[i0:i0+s], 1, m, 1.0, b[i0+s:i0+2*s], 1, m)
    t = timed(level); tot += t
    fl = (m // (2 * s)) * 2.0 * s * s * n
    print(f"level s={s:6d}: {m//(2*s):4d} gemms  {t:8.2f} ms  {fl/t/1e9:6.1f} TF")
    s *= 2
print(f"sum of gemm levels {tot:.1f} ms")
