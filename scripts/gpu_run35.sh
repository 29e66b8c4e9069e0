#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python scripts/gpu_eager_baseline.py vitl 16 518 > gpurun_out/eager_vitl.log 2>&1; echo "eager exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --size 1036 --batch 16 > gpurun_out/bench_cfg5.log 2>&1; echo "bench cfg5 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --encoder vitb --batch 32 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --encoder vits --batch 64 > gpurun_out/bench_vits.log 2>&1; echo "bench vits exit $?" >> gpurun_out/summary.txt
