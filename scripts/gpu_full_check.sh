#!/bin/bash
# What the driver runs at round end (pytest -m gpu in ONE process, smoke, bench).  Usage: gpurun --timeout 3000 -- 'bash scripts/gpu_full_check.sh'
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 2400 python -m pytest tests/ -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?" >> gpurun_out/summary.txt; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_default.log | cut -c1-400
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "bench reference exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_reference.log | cut -c1-300
cat gpurun_out/summary.txt
