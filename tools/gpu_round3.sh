#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemmt_gpu.py tests/test_gemm_gpu.py -x -q > gpurun_out/pytest_gemmt.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemmt.log
tail -15 gpurun_out/pytest_gemmt.log
timeout 600 python -m tools.gpu_probe_orient 8192 sdcz > gpurun_out/probe_orient.log 2>&1; tail -6 gpurun_out/probe_orient.log
timeout 300 python -m tools.gpu_probe_l3 16384 sc > gpurun_out/probe_l3.log 2>&1; tail -4 gpurun_out/probe_l3.log
