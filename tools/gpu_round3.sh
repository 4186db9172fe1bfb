#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_strucmm_gpu.py tests/test_gemmt_gpu.py tests/test_gemm_gpu.py tests/test_blis_dropin_gpu.py -x -q > gpurun_out/pytest_l3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_l3.log
tail -15 gpurun_out/pytest_l3.log
timeout 600 python -m tools.gpu_probe_l3b 16384 dszc > gpurun_out/probe_l3b.log 2>&1; tail -6 gpurun_out/probe_l3b.log
timeout 300 python -m tools.gpu_probe2 d 9,7,8,6 16384x64,8192x128,4096x64 > gpurun_out/probe_skinny.log 2>&1; tail -5 gpurun_out/probe_skinny.log
