#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 --no-header -p no:cacheprovider -x > gpurun_out/ops.log 2>&1; echo "ops exit $?" >> gpurun_out/summary.txt; tail -n 15 gpurun_out/ops.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 --no-header -p no:cacheprovider > gpurun_out/model.log 2>&1; echo "model exit $?" >> gpurun_out/summary.txt; tail -n 8 gpurun_out/model.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g2.log 2>&1; echo "bench g2 exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_g2.log | cut -c1-400
DAV2_GEMM2=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g1.log 2>&1; echo "bench g1 exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_g1.log | cut -c1-400
cat gpurun_out/summary.txt
