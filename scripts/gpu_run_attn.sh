#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python scripts/prof_ops.py attn 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 1 -o gpurun_out/prof_attn_r01_v2 python scripts/prof_ops.py attn 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
