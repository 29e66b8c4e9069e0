#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/attn_time.log
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention or attn" > gpurun_out/pytest_a.log 2>&1; echo "pytest attn exit $?" >> gpurun_out/summary.txt
for sp in 1 0; do for e in 0 2 3; do DAV2_TIME=1 DAV2_ATTN_SPLIT=$sp DAV2_ATTN_EMU=$e timeout 300 python scripts/prof_ops.py attn 1 2>&1 | sed "s/^/split=$sp /" >> gpurun_out/attn_time.log; done; done
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu > gpurun_out/pytest_model.log 2>&1; echo "pytest model exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v4.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attention' -c 1 -o gpurun_out/prof_attn_r01_v4 python scripts/prof_ops.py attn 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
