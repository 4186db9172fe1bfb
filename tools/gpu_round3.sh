#!/bin/bash
# gemmt-family round: new tests first, then drop-in, then perf probe and the headline bench (regression check)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemmt_gpu.py -x -q > gpurun_out/pytest_gemmt.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gemmt.log
tail -15 gpurun_out/pytest_gemmt.log
timeout 1500 python -m pytest tests/test_blis_dropin_gpu.py -x -q > gpurun_out/pytest_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dropin.log
tail -15 gpurun_out/pytest_dropin.log
timeout 600 python -m tools.gpu_probe_l3 8192,16384 dszc > gpurun_out/probe_l3.log 2>&1; tail -12 gpurun_out/probe_l3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-1200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
