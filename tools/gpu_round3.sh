#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemm_gpu.py tests/test_trsm_gpu.py -x -q > gpurun_out/pytest_l3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_l3.log
tail -6 gpurun_out/pytest_l3.log
timeout 300 python -m tools.gpu_probe2 d -1,7,10 256,512,768,1024,1536,2048 > gpurun_out/probe_d.log 2>&1; tail -3 gpurun_out/probe_d.log
timeout 300 python -m tools.gpu_probe2 s -1,3,4,5 256,512,768,1024,1536,2048 > gpurun_out/probe_s.log 2>&1; tail -4 gpurun_out/probe_s.log
