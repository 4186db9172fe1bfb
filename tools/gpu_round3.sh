#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/pytest_l3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_l3.log
tail -6 gpurun_out/pytest_l3.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e'])"; tail -3 gpurun_out/bench.err
timeout 300 python -m tools.gpu_probe2 d -1 512,768,1024,1280,1536,2048 > gpurun_out/probe_d.log 2>&1; tail -1 gpurun_out/probe_d.log
timeout 300 python -m tools.gpu_probe2 s -1 512,768,1024,1280,1536,2048 > gpurun_out/probe_s.log 2>&1; tail -1 gpurun_out/probe_s.log
