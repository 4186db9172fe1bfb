#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemm_gpu.py tests/test_gemmt_gpu.py -x -q > gpurun_out/pytest_l3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_l3.log
tail -6 gpurun_out/pytest_l3.log
timeout 300 python -m tools.gpu_probe2 d 9 16384x64,8192x128,4096x64,1024,2048,4096,16384 > gpurun_out/probe_d.log 2>&1; tail -2 gpurun_out/probe_d.log
timeout 300 python -m tools.gpu_probe2 s 3 16384x64,4096x64,1024,2048,4096,16384 > gpurun_out/probe_s.log 2>&1; tail -2 gpurun_out/probe_s.log
timeout 300 python -m tools.gpu_probe2 c 3 16384x64,4096x64,2048,8192 > gpurun_out/probe_c.log 2>&1; tail -2 gpurun_out/probe_c.log
