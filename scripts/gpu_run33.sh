#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/pytest_a.log 2>&1; echo "pytest ops exit $?" >> gpurun_out/summary.txt
DAV2_TIME=1 timeout 300 python scripts/prof_ops.py gemm 1 > gpurun_out/gemm_time.log 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
