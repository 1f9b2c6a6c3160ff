#!/bin/bash
# last visit of the round: default bench + the score-kernel ncu tables for the final code (the reports stay on the box)
mkdir -p gpurun_out
rm -f gpurun_out/diag_summary.txt
bash scripts/gpu_diag.sh tests/test_gpu_c_score.py tests/test_gpu_extract.py > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_tcgen05|ascore_tcgen05|kpsim_tcgen05|rownorm_inv|kp_sim_rows|flow_rows|gemm_tcgen05' -c 30 -f -o gpurun_out/r02_scores \
  python scripts/score_kernels_once.py > gpurun_out/ncu_scores.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02_scores.ncu-rep > gpurun_out/r02_ncu_scores.md 2> gpurun_out/ncu_summary.err
rm -f gpurun_out/*.ncu-rep
grep -n "passed\|failed\|exit" gpurun_out/diag_summary.txt
