#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu_model.py -x -q -m gpu -s -k "fp32 or matches_oracle" > gpurun_out/pytest_model.log 2>&1; echo "pytest model exit $?" >> gpurun_out/summary.txt
