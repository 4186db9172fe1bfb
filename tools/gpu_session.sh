#!/bin/bash
# One GPU session, parameterised by stage names (replaces the round-1 one-off scripts):
#   tools/gpu_session.sh [stage ...]       stages run in the order given; outputs land in gpurun_out/
# Stages:
#   newtests   the parity tests added this round (TMA paths, skinny shapes, tile counters, fused trsm)
#   dropin     tests/test_blis_dropin_gpu.py (reference testsuite + netlib ?blat3 on the plugin, the gemmsup slot, config/b200)
#   pytest     the whole -m gpu suite
#   smoke      __graft_entry__.smoke()
#   sanitize   compute-sanitizer memcheck + racecheck over tools/sanitize_driver.py
#   bench      both bench arms (reference first)
#   dtrsm      bench.py --op dtrsm
#   launches   ncu launch list of the default bench command
#   ncu_dgemm / ncu_trsm / ncu_skinny   one `ncu --set full` capture of that kernel + tools/ncu_key.py summary
#   sweep      configs[2] sweep (s/d/c/z squares + k=64)
#   midsize / ncu_splitk   dgemm split-k tail: mid-size sweep with and without it, one ncu capture at 2048^3
#   batchprobe   batched small gemm: loop vs stream pool vs grouped kernel
#   probes     tools/lds_probe.cu and tools/ffma2_probe.cu microbenchmarks
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
lscpu | head -25 > gpurun_out/lscpu.txt
for stage in "$@"; do
echo "=== stage $stage ($(date +%T))"
case "$stage" in
newtests)
	( time timeout 1500 python -m pytest tests -x -q -m gpu -k "fused or tma_paths or large_ragged or skinny_k64 or tile_counters" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
	tail -12 gpurun_out/pytest_new.log ;;
dropin)
	( time timeout 2400 python -m pytest tests/test_blis_dropin_gpu.py -q -m gpu ) > gpurun_out/pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dropin.log
	tail -25 gpurun_out/pytest_dropin.log ;;
distN)
	# needs gpurun --gpus N: the native multi-GPU driver under torchrun, then the bench at N (reference arm skipped)
	N=$(nvidia-smi -L | wc -l)
	timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/dist_native_driver.py > gpurun_out/dist_native_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/dist_native_n$N.log
	grep -c '_ok": true' gpurun_out/dist_native_n$N.log; grep -o '"[a-z_0-9]*_ok": false' gpurun_out/dist_native_n$N.log | sort | uniq -c; tail -3 gpurun_out/dist_native_n$N.log | cut -c1-600
	timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
	cat gpurun_out/bench_n$N.json | cut -c1-6000; tail -5 gpurun_out/bench_n$N.err ;;
pytest)
	( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
	tail -8 gpurun_out/pytest_gpu.log ;;
smoke)
	timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log ;;
sanitize)
	timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_driver.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck.log
	tail -4 gpurun_out/sanitize_memcheck.log
	timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_driver.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck.log
	tail -4 gpurun_out/sanitize_racecheck.log ;;
bench)
	timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
	timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
dtrsm)
	timeout 600 python bench.py --op dtrsm --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_dtrsm.json 2> gpurun_out/bench_dtrsm.err; cat gpurun_out/bench_dtrsm.json; tail -3 gpurun_out/bench_dtrsm.err ;;
launches)
	timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-peak > gpurun_out/bench_under_ncu.log 2>&1 ;;
ncu_dgemm)
	ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 3 -c 1 -o gpurun_out/dgemm_full -f python bench.py --op dgemm --steps 2 --warmup 3 --no-cpu --no-e2e --no-peak > gpurun_out/ncu_full.log 2>&1
	python tools/ncu_key.py gpurun_out/dgemm_full.ncu-rep > gpurun_out/ncu_dgemm_16384.txt 2>&1; head -12 gpurun_out/ncu_dgemm_16384.txt; rm -f gpurun_out/dgemm_full.ncu-rep ;;
ncu_trsm)
	ncu --set full --clock-control none --import-source on -k regex:trsm_ -s 40 -c 1 -o gpurun_out/trsm_full -f python bench.py --op dtrsm --steps 1 --warmup 3 --no-cpu --no-e2e --no-peak > gpurun_out/ncu_trsm.log 2>&1
	python tools/ncu_key.py gpurun_out/trsm_full.ncu-rep > gpurun_out/ncu_trsm_panel.txt 2>&1; head -12 gpurun_out/ncu_trsm_panel.txt; rm -f gpurun_out/trsm_full.ncu-rep ;;
ncu_skinny)
	ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 1 -c 1 -o gpurun_out/skinny_full -f python -m tools.one_gemm d 16384 64 -1 2 > /dev/null 2>&1
	python tools/ncu_key.py gpurun_out/skinny_full.ncu-rep > gpurun_out/ncu_dgemm_k64.txt 2>&1; head -12 gpurun_out/ncu_dgemm_k64.txt
	python tools/ncu_src.py gpurun_out/skinny_full.ncu-rep 25 > gpurun_out/ncu_dgemm_k64_src.txt 2>&1; rm -f gpurun_out/skinny_full.ncu-rep ;;
ncu_sgemm)
	ncu --set full --clock-control none --import-source on -k regex:gemm_ffma_tma -s 1 -c 1 -o gpurun_out/sgemm_full -f python -m tools.one_gemm s 8192 8192 -1 2 > /dev/null 2>&1
	python tools/ncu_key.py gpurun_out/sgemm_full.ncu-rep > gpurun_out/ncu_sgemm_8192.txt 2>&1; head -30 gpurun_out/ncu_sgemm_8192.txt
	python tools/ncu_src.py gpurun_out/sgemm_full.ncu-rep 60 > gpurun_out/ncu_sgemm_8192_src.txt 2>&1; rm -f gpurun_out/sgemm_full.ncu-rep ;;
ncu_splitk)
	ncu --set full --clock-control none --import-source on -k regex:gemm_dmma_tma -s 1 -c 1 -o gpurun_out/splitk_full -f python -m tools.one_gemm d 2048 2048 -1 2 > /dev/null 2>&1
	python tools/ncu_key.py gpurun_out/splitk_full.ncu-rep > gpurun_out/ncu_dgemm_2048_splitk.txt 2>&1; head -30 gpurun_out/ncu_dgemm_2048_splitk.txt
	python tools/ncu_src.py gpurun_out/splitk_full.ncu-rep 70 > gpurun_out/ncu_dgemm_2048_splitk_src.txt 2>&1; rm -f gpurun_out/splitk_full.ncu-rep ;;
midsize)
	timeout 300 python -m tools.midsize_sweep 1536,1792,2048,2304,2560,2816,3072,3328,3584,4096 > gpurun_out/midsize_sweep.json 2> gpurun_out/midsize_sweep.err; cat gpurun_out/midsize_sweep.err ;;
batchprobe)
	timeout 400 python -m tools.gpu_probe_batch 512 16,32,64,128,256 2>&1 | tail -12 ;;
probes)
	nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/lds_probe.cu -o /tmp/lds_probe && /tmp/lds_probe > gpurun_out/lds_probe.txt 2>&1
	nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ffma2_probe.cu -o /tmp/ffma2_probe && /tmp/ffma2_probe > gpurun_out/ffma2_probe.txt 2>&1; cat gpurun_out/ffma2_probe.txt ;;
sweep)
	for ch in d s c z; do timeout 400 python -m tools.gpu_probe2 $ch -1 512,1024,2048,4096,8192,16384,512x64,4096x64,16384x64 > gpurun_out/sweep_$ch.log 2>&1; tail -1 gpurun_out/sweep_$ch.log; done ;;
*)
	echo "unknown stage $stage" ;;
esac
done
ls -la gpurun_out | head -40
