#!/bin/bash
# One GPU-box visit for round 2's evidence: every -m gpu test file (one process each), smoke, default bench + reference arm,
# per-kernel timings, ncu launch list of one bench step, full ncu captures of the step's kernels and of the score kernels.
mkdir -p gpurun_out
bash scripts/gpu_validate.sh nobench
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/bench_ref.err
echo "=== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 400 --csv --log-file gpurun_out/r02_launches_bench_b128.csv \
  python bench.py --steps 1 --warmup 1 --batch 128 --chunk 128 --no-cpu-baseline --no-scores > gpurun_out/ncu_launches.log 2>&1
echo "=== ncu full (step kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm2?_tcgen05|attention4_tcgen05|layernorm' -s 40 -c 12 -f -o gpurun_out/r02_step \
  python bench.py --steps 1 --warmup 1 --batch 128 --chunk 128 --no-cpu-baseline --no-scores > gpurun_out/ncu_full.log 2>&1
echo "=== ncu full (score kernels)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_tcgen05|ascore_tcgen05|kpsim_tcgen05|rownorm_inv|kp_sim_rows|flow_rows|gemm_tcgen05' -c 30 -f -o gpurun_out/r02_scores \
  python scripts/score_kernels_once.py > gpurun_out/ncu_scores.log 2>&1
# the reports are too large to travel back (64 MiB limit): summarise them here, keep the tables and the traffic figure only
python scripts/ncu_summary.py gpurun_out/r02_step.ncu-rep gpurun_out/r02_scores.ncu-rep > gpurun_out/r02_ncu_summary.md 2> gpurun_out/ncu_summary.err
cp profiles/r02_traffic.json gpurun_out/r02_traffic.json 2>/dev/null
rm -f gpurun_out/r02_step.ncu-rep gpurun_out/r02_scores.ncu-rep
ls -la gpurun_out | tail -12
