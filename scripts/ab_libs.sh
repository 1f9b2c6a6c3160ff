#!/bin/bash
# A/B of two builds of the library on one box: per-kernel timings, tower throughput, score kernels.  usage: ab_libs.sh <variant.so>
V=$(realpath "$1"); mkdir -p gpurun_out; O=gpurun_out/ab_libs.txt; : > $O
for lib in "" "$V"; do
  echo "##### LOVR_B200_LIB=${lib:-<product>}" | tee -a $O
  export LOVR_B200_LIB=$lib
  timeout 300 python scripts/bench_kernels.py 64 2>&1 | tail -14 | tee -a $O
  timeout 300 python scripts/bench_towers.py 2>&1 | grep tower | tee -a $O
  timeout 300 python scripts/bench_scores.py 2>&1 | tail -6 | tee -a $O
done
unset LOVR_B200_LIB
