#!/bin/bash
# First GPU session of the next round: what this round left unmeasured.
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_next_round.sh 8'      (N = 2, 4 or 8; default 1)
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" -gt 1 ]; then
  # e2e with host-resident shards at N GPUs + the bit-for-bit check of step_host against step
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      tests/dist_host_driver.py > gpurun_out/dist_host_n$N.log 2>&1; grep -o "{.*}" gpurun_out/dist_host_n$N.log | cut -c1-200
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e'])"
else
  # dtrsm T1 with a host-resident A: triangle-only upload (stage_tri_to_device) against the full square
  timeout 300 python - <<'PY' > gpurun_out/trsm_host_a.log 2>&1
import time, torch
from blis_b200 import api
m, n = 32768, 8192
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = (torch.rand(m, m, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) / 128
a.diagonal().add_(2.0)
ah = torch.empty(m, m, dtype=torch.float64).pin_memory(); ah.copy_(a); ah = ah.t()       # column-major, lower triangle read
b0 = (torch.rand(n, m, dtype=torch.float64, device="cuda", generator=g) * 2 - 1).t()
for where, aa in (("device A", a.t()), ("pinned host A", ah)):
    for rep in range(3):
        b = b0.clone(memory_format=torch.preserve_format); torch.cuda.synchronize()
        t0 = time.perf_counter(); api.bli_dtrsm(0, 0xC0, 0, 0, m, n, 2.0, aa, 1, m, b, 1, m); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{where}: {dt * 1e3:.1f} ms  {m * m * n / dt / 1e12:.2f} TFLOP/s", flush=True)
PY
  tail -6 gpurun_out/trsm_host_a.log
fi
