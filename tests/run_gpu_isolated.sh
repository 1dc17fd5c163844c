#!/bin/bash
# Run every GPU test function in its own process (a trapped kernel poisons the CUDA context of
# its process only) and collect a summary under gpurun_out/.
mkdir -p gpurun_out
: > gpurun_out/isolated_summary.txt
FILE=${1:-tests/test_kernels_gpu.py}
for fn in $(grep -oE "^def (test_[a-zA-Z0-9_]+)" "$FILE" | awk '{print $2}'); do
  timeout 300 python -m pytest "$FILE" -m gpu -q -k "$fn" -p no:cacheprovider > "gpurun_out/iso_$fn.log" 2>&1
  rc=$?
  echo "$fn rc=$rc $(tail -1 gpurun_out/iso_$fn.log)" | tee -a gpurun_out/isolated_summary.txt
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/isolated_summary.txt 2>&1
