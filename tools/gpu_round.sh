#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), launch list + full ncu capture of the dgemm kernel.
set -x
mkdir -p gpurun_out
lscpu | head -25 > gpurun_out/lscpu.txt
nvidia-smi > gpurun_out/nvidia-smi.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --op dtrsm --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_dtrsm.json 2> gpurun_out/bench_dtrsm.err; cat gpurun_out/bench_dtrsm.json; tail -3 gpurun_out/bench_dtrsm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-peak > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 3 -c 1 -o gpurun_out/dgemm_full python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-peak > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
