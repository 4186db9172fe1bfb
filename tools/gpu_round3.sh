#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemm_gpu.py tests/test_trsm_gpu.py tests/test_gemmt_gpu.py -x -q > gpurun_out/pytest_l3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_l3.log
tail -6 gpurun_out/pytest_l3.log
timeout 300 python -m tools.gpu_probe2 d 9 16384x64,8192x128,16384x256,8192x512,4096x1024,4096x64,16384 > gpurun_out/probe_d.log 2>&1; tail -1 gpurun_out/probe_d.log
python - <<'PY'
from blis_b200 import api
api.set_option("dmma_cst", 0)
import subprocess, sys
PY
B200_NOCST=1 timeout 300 python -c "
from blis_b200 import api
api.set_option('dmma_cst', 0)
import sys; sys.argv=['x','d','9','16384x64,8192x128,16384x256,8192x512,4096x1024']
from tools import gpu_probe2; gpu_probe2.main()" > gpurun_out/probe_d_nocst.log 2>&1; tail -1 gpurun_out/probe_d_nocst.log
