#!/bin/bash
# The ncu evidence under profiles/ (one GPU; every command first exits 0 without ncu).
# Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh'; then copy gpurun_out/{launches_step.csv,ncu_ops_metrics.csv} and the
# `ncu -i <rep> --page details` text of the .ncu-rep files into profiles/ (named per round).
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
# 1. per-launch time + DRAM bytes of exactly one timed bench step
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
BENCH_CUDA_PROFILER=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
# 2. pipe utilisation per operator-shaped launch
timeout 300 python scripts/prof_ops.py all 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_ops_metrics.csv python scripts/prof_ops.py all 1 > gpurun_out/ncu_ops.log 2>&1
echo "ops metrics exit $?" >> gpurun_out/summary.txt
# 3. full captures of the kernels under work
DAV2_QKV_SCALE=0.35 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention' -c 1 -o gpurun_out/prof_attn python scripts/prof_ops.py attn 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'backproject|depth_metrics' -c 3 -o gpurun_out/prof_geom python scripts/prof_ops.py geom 1 > gpurun_out/ncu_geom.log 2>&1
echo "ncu geom exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
