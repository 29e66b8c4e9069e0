#!/bin/bash
# round-2 run 3: attention A/B variants on ONE box (same clocks), then one ncu capture of the current kernel
mkdir -p gpurun_out/r2
for v in r1 v1 nopre nofence mmaunroll; do
  for sc in 0.35; do for e in 2 3; do
    PROF_LIB=$PWD/gpurun_variants/libdav2_attn_$v.so DAV2_QKV_SCALE=$sc DAV2_TIME=1 DAV2_ATTN_EMU=$e timeout 120 python scripts/prof_ops.py attn 1 2>&1 | grep "attention EMU" | sed "s/^/$v scale $sc /"
  done; done
done
PROF_LIB=$PWD/gpurun_variants/libdav2_attn_v1.so DAV2_QKV_SCALE=0.35 timeout 200 python scripts/prof_ops.py attn 1 > gpurun_out/r2/attn_plain.log 2>&1 && \
PROF_LIB=$PWD/gpurun_variants/libdav2_attn_v1.so DAV2_QKV_SCALE=0.35 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -c 1 -o gpurun_out/r2/prof_attn_v1 python scripts/prof_ops.py attn 1 > gpurun_out/r2/ncu_attn_v1.log 2>&1
echo "ncu exit $?"
