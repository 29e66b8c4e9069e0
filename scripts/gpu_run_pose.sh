#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_pose.py -q -m gpu --timeout 300 --no-header -p no:cacheprovider > gpurun_out/pose.log 2>&1; echo "pose exit $?" >> gpurun_out/summary.txt; tail -n 30 gpurun_out/pose.log
cat gpurun_out/summary.txt
