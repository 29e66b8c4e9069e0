#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python scripts/prof_ops.py conv 1 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_halo -c 3 -o gpurun_out/prof_halo_r01 python scripts/prof_ops.py conv 1 > gpurun_out/ncu_halo.log 2>&1
echo "ncu halo exit $?" >> gpurun_out/summary.txt
