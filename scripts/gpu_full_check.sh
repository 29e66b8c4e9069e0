#!/bin/bash
# What the driver runs at round end (pytest -m gpu in ONE process, smoke, bench) + the ncu launch list with DRAM traffic.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?" >> gpurun_out/summary.txt; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_default.log | cut -c1-300
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 648 -c 216 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
