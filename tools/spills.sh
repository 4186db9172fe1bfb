#!/bin/bash
# list kernels with register spills (name, stack, spill stores, spill loads) over all gemm translation units
cd /root/repo/blis_b200/csrc
for f in gemm_d.cu gemm_z.cu gemm_s.cu gemm_c.cu; do
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -c $f -o /dev/null 2>&1 | \
  awk '/Function properties for/ {name=$NF} /bytes spill/ { if ($0 !~ / 0 bytes spill stores, 0 bytes spill loads/) print name, $1, $5, $9 }' | c++filt | sed 's/(b200::GemmArgs<[a-z0-9]*>[^)]*)//' &
done | sort
wait
