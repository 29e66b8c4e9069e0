#!/bin/bash
# final round-1 evidence: launch list of one bench step (+DRAM bytes), full captures of attention / geometry kernels
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 648 -c 216 --csv --log-file gpurun_out/launches_r01_step_v3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
timeout 300 python scripts/prof_ops.py all 1 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_ops_metrics_r01_v5.csv python scripts/prof_ops.py all 1 > gpurun_out/ncu_ops.log 2>&1
echo "ops metrics exit $?" >> gpurun_out/summary.txt
DAV2_QKV_SCALE=0.35 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention' -c 1 -o gpurun_out/prof_attn_r01_v5 python scripts/prof_ops.py attn 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'backproject|depth_metrics' -c 3 -o gpurun_out/prof_geom_r01_v5 python scripts/prof_ops.py geom 1 > gpurun_out/ncu_geom.log 2>&1
echo "ncu geom exit $?" >> gpurun_out/summary.txt
