#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_geometry_metrics.py -x -q -m gpu > gpurun_out/pytest_a.log 2>&1; echo "pytest geom exit $?" >> gpurun_out/summary.txt
timeout 300 python scripts/prof_ops.py geom 1 > gpurun_out/prof_geom_plain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'backproject|depth_metrics' -c 3 -o gpurun_out/prof_geom_r01_v4 python scripts/prof_ops.py geom 1 > gpurun_out/ncu_geom.log 2>&1
echo "ncu geom exit $?" >> gpurun_out/summary.txt
