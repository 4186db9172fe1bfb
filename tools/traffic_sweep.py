"""DRAM traffic of the dgemm kernel versus raster / L2-promotion settings (dev tool, run under ncu):
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
      --clock-control none -k regex:gemm_ -c 12 --csv --log-file gpurun_out/traffic.csv python -m tools.traffic_sweep d 16384
Every launch uses the next configuration of CONFIGS (printed in order)."""
import sys
import torch
from blis_b200 import api
from tools.gpu_probe2 import DT, FN, rnd

CONFIGS = [(8, 2), (4, 2), (12, 2), (16, 2), (24, 2), (32, 2), (8, 3), (16, 3), (8, 0), (2, 2), (1, 2), (64, 2)]   # (raster group, TMA L2 promotion)
ch = sys.argv[1] if len(sys.argv) > 1 else "d"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
dt = DT[ch]
a, b, c = rnd(n, n, dt), rnd(n, n, dt), rnd(n, n, dt)
for raster, promo in CONFIGS:
    api.set_option("raster_group", raster)
    api.set_option("tma_l2_promotion", promo)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    FN[ch](0, 0, n, n, n, 2.0, a, 1, n, b, 1, n, 1.2, c, 1, n)
    e1.record(); torch.cuda.synchronize()
    print(f"raster {raster} hints {promo}: {e0.elapsed_time(e1):.2f} ms", flush=True)
