#!/bin/bash
# round-2 run 1: precision study + batch-size sweep (L2-resident sub-batching question)
mkdir -p gpurun_out/r2
for e in vitl vits; do timeout 600 python tests/gpu_precision_study.py $e 518 1 > gpurun_out/r2/prec_$e.log 2>&1; tail -n 4 gpurun_out/r2/prec_$e.log | cut -c1-400; done
for b in 64 32 16 8; do timeout 600 python bench.py --batch $b --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_b$b.log 2>&1; tail -n 1 gpurun_out/r2/bench_b$b.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('B',d['config']['batch_per_gpu'],'fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'clk',d['clocks']['sm_mhz'],{k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"; done
timeout 1200 python -m pytest tests/test_gpu_parity_configs.py tests/test_gpu_geometry_metrics.py tests/test_host_metrics.py -q -m gpu -s -p no:cacheprovider > gpurun_out/r2/pytest_parity.log 2>&1; echo "pytest parity exit $?"; grep -E "fp16|bf16|tap|config 1|passed|failed|Error|assert" gpurun_out/r2/pytest_parity.log | cut -c1-330 | tail -40
