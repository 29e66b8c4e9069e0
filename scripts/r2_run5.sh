#!/bin/bash
mkdir -p gpurun_out/r2
for rep in 1 2; do
for v in r1 v3 cl clnopre; do
  for sc in 0.35; do for e in 2; do
    PROF_LIB=$PWD/gpurun_variants/libdav2_attn_$v.so DAV2_QKV_SCALE=$sc DAV2_TIME=1 DAV2_ATTN_EMU=$e timeout 120 python scripts/prof_ops.py attn 1 2>&1 | grep "attention EMU" | sed "s/^/$v scale $sc /"
  done; done
done; done
for v in cl clnopre; do PROF_LIB=$PWD/gpurun_variants/libdav2_attn_$v.so timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" -p no:cacheprovider 2>&1 | tail -n 1; done
