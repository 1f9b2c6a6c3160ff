#!/bin/bash
# Run every -m gpu test file in its own process (a device trap poisons the CUDA context) with a timeout
# (a hung pipeline must not hold the box), logging into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/diag_smi.txt 2>&1
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" >> gpurun_out/diag_smi.txt 2>&1
for f in "$@"; do
  name=$(echo "$f" | tr '/:[]' '____')
  echo "=== $f" | tee -a gpurun_out/diag_summary.txt
  timeout 600 python -m pytest "$f" -q -m gpu --timeout=300 -s -p no:cacheprovider > "gpurun_out/diag_${name}.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/diag_summary.txt
  tail -n 25 "gpurun_out/diag_${name}.log" | tee -a gpurun_out/diag_summary.txt
done
