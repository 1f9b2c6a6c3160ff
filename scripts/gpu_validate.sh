#!/bin/bash
# One GPU-box visit without profilers: every -m gpu test file (one process each), smoke, per-kernel timings, default bench.
mkdir -p gpurun_out
rm -f gpurun_out/diag_summary.txt
bash scripts/gpu_diag.sh tests/test_gpu_preprocess.py tests/test_gpu_c_score.py tests/test_gpu_kernels.py tests/test_gpu_tower.py \
  tests/test_gpu_a_score.py tests/test_gpu_cka.py tests/test_gpu_projector.py tests/test_gpu_edges.py tests/test_gpu_fullsize.py \
  tests/test_gpu_extract.py
echo "=== smoke" | tee -a gpurun_out/diag_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/diag_summary.txt
echo "=== kernels" | tee -a gpurun_out/diag_summary.txt
timeout 300 python scripts/bench_kernels.py 64 2>&1 | tee gpurun_out/bench_kernels.log | tail -20 | tee -a gpurun_out/diag_summary.txt
if [ "$1" != "nobench" ]; then
echo "=== bench" | tee -a gpurun_out/diag_summary.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json | tee -a gpurun_out/diag_summary.txt
fi
