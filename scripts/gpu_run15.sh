#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for bo in 0 1; do
  DAV2_HALO_BO=$bo timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k conv3x3 --timeout 120 --no-header -p no:cacheprovider > gpurun_out/conv_bo$bo.log 2>&1; echo "conv bo=$bo exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/conv_bo$bo.log
done
cat gpurun_out/summary.txt
