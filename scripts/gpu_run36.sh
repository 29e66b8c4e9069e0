#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/attn_time.log
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention or attn" > gpurun_out/pytest_a.log 2>&1; echo "pytest attn kv64 exit $?" >> gpurun_out/summary.txt
DAV2_ATTN_KV=128 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention or attn" > gpurun_out/pytest_a128.log 2>&1; echo "pytest attn kv128 exit $?" >> gpurun_out/summary.txt
for kv in 64 128; do for e in 0 2 3; do DAV2_QKV_SCALE=0.35 DAV2_TIME=1 DAV2_ATTN_KV=$kv DAV2_ATTN_EMU=$e timeout 300 python scripts/prof_ops.py attn 1 2>&1 | sed "s/^/scale0.35 kv=$kv /" >> gpurun_out/attn_time.log; done; done
for kv in 64 128; do DAV2_TIME=1 DAV2_ATTN_KV=$kv timeout 300 python scripts/prof_ops.py attn 1 2>&1 | sed "s/^/scale1 kv=$kv /" >> gpurun_out/attn_time.log; done
