#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/attn_time.log
for d in 1 2 3 4; do DAV2_LIB_PATH=$PWD/gpurun_variants/libdav2_dbg$d.so DAV2_TIME=1 DAV2_ATTN_SPLIT=0 DAV2_ATTN_EMU=0 timeout 300 python scripts/prof_ops.py attn 1 2>&1 | sed "s/^/dbg=$d /" >> gpurun_out/attn_time.log; done
echo done >> gpurun_out/summary.txt
