#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for kv in 64 128; do
DAV2_ATTN_KV=$kv timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kv$kv.log 2>&1; echo "bench kv$kv exit $?" >> gpurun_out/summary.txt
DAV2_ATTN_KV=$kv timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --size 1036 --batch 16 > gpurun_out/bench_cfg5_kv$kv.log 2>&1; echo "bench cfg5 kv$kv exit $?" >> gpurun_out/summary.txt
done
