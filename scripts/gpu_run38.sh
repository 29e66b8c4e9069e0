#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "bilinear" > gpurun_out/pytest_a.log 2>&1; echo "pytest bilinear exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v7.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "matches_oracle" > gpurun_out/pytest_model.log 2>&1; echo "pytest model exit $?" >> gpurun_out/summary.txt
