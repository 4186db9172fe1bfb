#!/bin/bash
# Final round-1 session: full parity suite, smoke, both bench arms, dtrsm, the config #3 sweep, ncu summaries.
mkdir -p gpurun_out
lscpu | head -25 > gpurun_out/lscpu.txt
nvidia-smi > gpurun_out/nvidia-smi.txt
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 500 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --op dtrsm --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_dtrsm.json 2> gpurun_out/bench_dtrsm.err; cut -c1-200 gpurun_out/bench_dtrsm.json
for ch in d s c z; do timeout 400 python -m tools.gpu_probe2 $ch -1 512,1024,2048,4096,8192,16384,512x64,4096x64,16384x64 > gpurun_out/sweep_$ch.log 2>&1; tail -1 gpurun_out/sweep_$ch.log; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-peak > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none -k regex:gemm_dmma -s 3 -c 1 -o /tmp/dgemm_full -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-peak > gpurun_out/ncu_full.log 2>&1; python tools/ncu_key.py /tmp/dgemm_full.ncu-rep > gpurun_out/ncu_dgemm_16384.txt 2>&1
ncu --set full --clock-control none -k regex:gemm_ -s 1 -c 1 -o /tmp/sk -f python -m tools.one_gemm d 16384 64 -1 2 > /dev/null 2>&1; python tools/ncu_key.py /tmp/sk.ncu-rep > gpurun_out/ncu_dgemm_k64_cst.txt 2>&1
ncu --set full --clock-control none -k regex:gemm_ -s 1 -c 1 -o /tmp/sg -f python -m tools.one_gemm s 16384 16384 -1 2 > /dev/null 2>&1; python tools/ncu_key.py /tmp/sg.ncu-rep > gpurun_out/ncu_sgemm_16384.txt 2>&1
ncu --set full --clock-control none -k regex:gemm_ -s 1 -c 1 -o /tmp/cg -f python -m tools.one_gemm c 8192 8192 -1 2 > /dev/null 2>&1; python tools/ncu_key.py /tmp/cg.ncu-rep > gpurun_out/ncu_cgemm_8192.txt 2>&1
ncu --set full --clock-control none -k regex:gemm_ -s 1 -c 1 -o /tmp/zg -f python -m tools.one_gemm z 8192 8192 1 2 > /dev/null 2>&1; python tools/ncu_key.py /tmp/zg.ncu-rep > gpurun_out/ncu_zgemm_8192.txt 2>&1
head -12 gpurun_out/ncu_dgemm_16384.txt
ls -la gpurun_out | head -40
