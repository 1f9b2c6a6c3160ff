#!/bin/bash
# A score: tests for both kernel forms, then C3 timing with the 2-CTA kernel and with the 1-CTA kernel
mkdir -p gpurun_out; O=gpurun_out/ab_ascore.txt; : > $O
timeout 600 python -m pytest tests/test_gpu_a_score.py -x -q -m gpu 2>&1 | tail -15 | tee -a $O
for v in 1 0; do
  echo "##### LOVR_A_SCORE_2CTA=$v" | tee -a $O
  LOVR_A_SCORE_2CTA=$v timeout 300 python scripts/bench_scores.py 2>&1 | tail -4 | tee -a $O
done
