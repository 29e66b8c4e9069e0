#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/attn_time.log
DAV2_QKV_SCALE=0.35 DAV2_LIB_PATH=$PWD/gpurun_variants/libdav2_trace.so DAV2_ATTN_SPLIT=0 DAV2_ATTN_EMU=2 timeout 300 python scripts/prof_ops.py attntrace 1 > gpurun_out/attn_trace.log 2>&1; echo "trace exit $?" >> gpurun_out/summary.txt
for sp in 1 0; do DAV2_QKV_SCALE=0.35 DAV2_TIME=1 DAV2_ATTN_SPLIT=$sp DAV2_ATTN_EMU=2 timeout 300 python scripts/prof_ops.py attn 1 2>&1 | sed "s/^/scale0.35 split=$sp /" >> gpurun_out/attn_time.log; done
