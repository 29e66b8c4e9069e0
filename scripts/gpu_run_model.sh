#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_host_metrics.py tests/test_gpu_pose.py -q -m gpu --timeout 300 --no-header -p no:cacheprovider > gpurun_out/model.log 2>&1; echo "model exit $?" >> gpurun_out/summary.txt; tail -n 25 gpurun_out/model.log
cat gpurun_out/summary.txt
