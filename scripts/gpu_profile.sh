#!/bin/bash
# The ncu evidence under profiles/ (one GPU; every command first exits 0 without ncu).
# Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh r02 [launches|ops|attn|attn5477|geom|conv|all]' -- ONE section
# (= one ncu session) per gpurun call is the rule on this pool, "all" is for a local box; then
# scripts/summarise_profiles.py copies the extracts of gpurun_out/prof/* into profiles/ (named per round).
R=${1:-r02}
S=${2:-all}
O=gpurun_out/prof
mkdir -p $O; rm -f $O/summary.txt
want() { [ "$S" = all ] || [ "$S" = "$1" ]; }
BENCH1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline"

if want launches; then  # 1. per-launch time + DRAM bytes of exactly one timed bench step
  timeout 600 $BENCH1 > $O/bench_plain.log 2>&1 && \
  BENCH_CUDA_PROFILER=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches_${R}_step.csv $BENCH1 > $O/ncu_launches.log 2>&1
  echo "launch list exit $?" >> $O/summary.txt
fi
if want ops; then       # 2. pipe utilisation per operator-shaped launch
  timeout 300 python scripts/prof_ops.py all 1 > $O/prof_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none --csv --log-file $O/ncu_ops_metrics_${R}.csv python scripts/prof_ops.py all 1 > $O/ncu_ops.log 2>&1
  echo "ops metrics exit $?" >> $O/summary.txt
fi
# 3. full captures of the kernels under work
if want attn; then
  DAV2_QKV_SCALE=0.35 timeout 300 python scripts/prof_ops.py attn 1 > $O/attn_plain.log 2>&1 && \
  DAV2_QKV_SCALE=0.35 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention' -c 1 -o $O/prof_attn_${R} python scripts/prof_ops.py attn 1 > $O/ncu_attn.log 2>&1
  echo "ncu attn exit $?" >> $O/summary.txt
fi
if want attn5477; then
  DAV2_QKV_SCALE=0.35 timeout 300 python scripts/prof_ops.py attn5477 1 > $O/attn5477_plain.log 2>&1 && \
  DAV2_QKV_SCALE=0.35 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention' -c 1 -o $O/prof_attn5477_${R} python scripts/prof_ops.py attn5477 1 > $O/ncu_attn5477.log 2>&1
  echo "ncu attn5477 exit $?" >> $O/summary.txt
fi
if want geom; then
  timeout 300 python scripts/prof_ops.py geom 1 > $O/geom_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'backproject|depth_metrics' -c 4 -o $O/prof_geom_${R} python scripts/prof_ops.py geom 1 > $O/ncu_geom.log 2>&1
  echo "ncu geom exit $?" >> $O/summary.txt
fi
if want conv; then      # the N = 128 / N = 32 halo convolutions inside one bench step (crossbar bytes, tensor pipe)
  timeout 600 $BENCH1 > $O/bench_plain.log 2>&1 && \
  BENCH_CUDA_PROFILER=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'conv_halo_tcgen05_kernel<\(int\)(128|32),' -c 2 -o $O/prof_conv_${R} $BENCH1 > $O/ncu_conv.log 2>&1
  echo "ncu conv exit $?" >> $O/summary.txt
fi
cat $O/summary.txt; grep -h "attention 16 x\|us/launch" $O/attn5477_plain.log $O/geom_plain.log 2>/dev/null; true
