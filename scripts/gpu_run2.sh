#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_geometry_metrics.py -q -m gpu --timeout 120 --no-header -p no:cacheprovider > gpurun_out/geom.log 2>&1; echo "geom exit $?" >> gpurun_out/summary.txt; tail -n 5 gpurun_out/geom.log
timeout 300 python scripts/gpu_debug_model.py vits 70 98 1 > gpurun_out/debug_model.log 2>&1; echo "debug_model exit $?" >> gpurun_out/summary.txt
tail -n 40 gpurun_out/debug_model.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu --timeout 300 --no-header -p no:cacheprovider > gpurun_out/model.log 2>&1; echo "model exit $?" >> gpurun_out/summary.txt; tail -n 30 gpurun_out/model.log
timeout 300 python bench.py --encoder vits --batch 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vits.log 2>&1; echo "bench_vits exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/bench_vits.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vitl.log 2>&1; echo "bench_vitl exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/bench_vitl.log
cat gpurun_out/summary.txt
