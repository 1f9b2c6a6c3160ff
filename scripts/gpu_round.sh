#!/bin/bash
# One GPU-box visit: parity tests (one process per file), smoke, bench, per-kernel timings, ncu launch list and
# one full ncu capture of the two tensor-core kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/diag_summary.txt
bash scripts/gpu_diag.sh "$@"
echo "=== smoke" | tee -a gpurun_out/diag_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/diag_summary.txt
echo "=== bench" | tee -a gpurun_out/diag_summary.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json | tee -a gpurun_out/diag_summary.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | tee -a gpurun_out/diag_summary.txt
echo "=== kernels" | tee -a gpurun_out/diag_summary.txt
timeout 300 python scripts/bench_kernels.py 64 2>&1 | tee gpurun_out/bench_kernels.log | tee -a gpurun_out/diag_summary.txt
echo "=== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline --no-a-score > gpurun_out/ncu_launches.log 2>&1
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm2?_tcgen05|attention2_tcgen05|layernorm' -s 40 -c 10 -f -o gpurun_out/prof_gemm_attn \
  python bench.py --steps 1 --warmup 1 --batch 128 --no-cpu-baseline --no-a-score > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -20
