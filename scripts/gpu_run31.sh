#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/attn_time.log
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention or attn" > gpurun_out/pytest_a.log 2>&1; echo "pytest attn exit $?" >> gpurun_out/summary.txt
for e in 0 2 3; do DAV2_QKV_SCALE=0.35 DAV2_TIME=1 DAV2_ATTN_EMU=$e timeout 300 python scripts/prof_ops.py attn 1 2>&1 | sed "s/^/scale0.35 /" >> gpurun_out/attn_time.log; done
DAV2_TIME=1 timeout 300 python scripts/prof_ops.py attn 1 2>&1 | sed "s/^/scale1 /" >> gpurun_out/attn_time.log
DAV2_QKV_SCALE=0.35 DAV2_LIB_PATH=$PWD/gpurun_variants/libdav2_trace.so timeout 300 python scripts/prof_ops.py attntrace 1 > gpurun_out/attn_trace.log 2>&1; echo "trace exit $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "matches_oracle or bf16" > gpurun_out/pytest_model.log 2>&1; echo "pytest model exit $?" >> gpurun_out/summary.txt
