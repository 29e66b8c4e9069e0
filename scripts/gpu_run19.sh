#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_geometry_metrics.py tests/test_gpu_ops.py -x -q -m gpu > gpurun_out/pytest_a.log 2>&1; echo "pytest geom+ops exit $?" >> gpurun_out/summary.txt
for e in 0 2 3 4; do DAV2_TIME=1 DAV2_ATTN_EMU=$e timeout 300 python scripts/prof_ops.py attn 1 >> gpurun_out/attn_time.log 2>&1; done
timeout 300 python scripts/prof_ops.py geom 1 > gpurun_out/prof_geom_plain.log 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu > gpurun_out/pytest_model.log 2>&1; echo "pytest model exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention_kernel' -c 1 -o gpurun_out/prof_attn_r01_v3 python scripts/prof_ops.py attn 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'backproject|depth_metrics' -c 3 -o gpurun_out/prof_geom_r01_v2 python scripts/prof_ops.py geom 1 > gpurun_out/ncu_geom.log 2>&1
echo "ncu geom exit $?" >> gpurun_out/summary.txt
