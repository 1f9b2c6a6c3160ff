#!/bin/bash
# parity tests (one process per file) + smoke + default bench; no profilers
mkdir -p gpurun_out
rm -f gpurun_out/diag_summary.txt
bash scripts/gpu_diag.sh tests/test_gpu_kernels.py tests/test_gpu_tower.py tests/test_gpu_a_score.py tests/test_gpu_c_score.py tests/test_gpu_cka.py
echo "=== smoke" | tee -a gpurun_out/diag_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/diag_summary.txt
echo "=== bench" | tee -a gpurun_out/diag_summary.txt
timeout 600 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json | tee -a gpurun_out/diag_summary.txt
