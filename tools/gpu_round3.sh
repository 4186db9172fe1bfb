#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemm_md_gpu.py tests/test_gemm_gpu.py -x -q > gpurun_out/pytest_l3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_l3.log
tail -12 gpurun_out/pytest_l3.log
timeout 300 python -m tools.gpu_probe2 d 9 16384x64,8192x128,16384 > gpurun_out/probe_d.log 2>&1; tail -2 gpurun_out/probe_d.log
