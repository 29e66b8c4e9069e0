#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python scripts/prof_ops.py all 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/ops_metrics.csv python scripts/prof_ops.py all 1 > gpurun_out/ncu_ops.log 2>&1
echo "ncu ops exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
