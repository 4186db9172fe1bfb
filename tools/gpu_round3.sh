#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_strucmm_gpu.py tests/test_blis_dropin_gpu.py -x -q > gpurun_out/pytest_l3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_l3.log
tail -12 gpurun_out/pytest_l3.log
