#!/bin/bash
# One GPU-box visit for the round's evidence: parity tests (one process per file), smoke, bench (+ reference arm), per-kernel
# timings, ncu launch list of one bench step and full ncu captures of the hot kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
bash scripts/gpu_validate.sh
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | tee -a gpurun_out/diag_summary.txt
echo "=== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline --no-a-score > gpurun_out/ncu_launches.log 2>&1
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm2?_tcgen05|attention4_tcgen05|layernorm' -s 40 -c 10 -f -o gpurun_out/prof_gemm_attn \
  python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline --no-a-score > gpurun_out/ncu_full.log 2>&1
timeout 280 ncu --set full --import-source on --clock-control none -k regex:resample -s 4 -c 2 -f -o gpurun_out/pre python scripts/pre_one.py > gpurun_out/ncu_pre.log 2>&1
ls -la gpurun_out | tail -20
