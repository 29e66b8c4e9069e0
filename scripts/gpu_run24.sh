#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/attn_time.log
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention or attn" > gpurun_out/pytest_a.log 2>&1; echo "pytest attn exit $?" >> gpurun_out/summary.txt
for sp in 1 0; do for e in 0 2 3; do DAV2_TIME=1 DAV2_ATTN_SPLIT=$sp DAV2_ATTN_EMU=$e timeout 300 python scripts/prof_ops.py attn 1 2>&1 | sed "s/^/split=$sp /" >> gpurun_out/attn_time.log; done; done
DAV2_ATTN_SPLIT=0 timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention or attn" > gpurun_out/pytest_a0.log 2>&1; echo "pytest attn split0 exit $?" >> gpurun_out/summary.txt
