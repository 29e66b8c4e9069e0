#!/bin/bash
# First GPU session: every test file in its own process (a trapped kernel poisons the CUDA context),
# bounded by timeouts; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_gpu_geometry_metrics test_gpu_ops; do
  timeout 600 python -m pytest tests/$t.py -q -m gpu --timeout 120 --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
  tail -n 25 gpurun_out/$t.log
done
timeout 300 python scripts/gpu_debug_model.py vits 70 98 1 > gpurun_out/debug_model.log 2>&1; echo "debug_model exit $?" >> gpurun_out/summary.txt
tail -n 40 gpurun_out/debug_model.log
cat gpurun_out/summary.txt
