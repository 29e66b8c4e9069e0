#!/bin/bash
O=gpurun_out/prof2; mkdir -p $O
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/bench_plain.log 2>&1 && \
BENCH_CUDA_PROFILER=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file $O/launches_lnfold.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/ncu_launches.log 2>&1
echo "launch list exit $?"
