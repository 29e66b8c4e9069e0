#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_gpu_model.py -q -x -p no:cacheprovider -k "forward_matches_oracle or taps or batch_consistency or bf16" 2>&1 | tail -n 15
timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -x -s -p no:cacheprovider -k "depth_logits_per_pixel or all_taps" 2>&1 | grep -E "fp16:|tap|passed|failed|Error|assert" | cut -c1-330 | tail -n 20
timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/bench_lnfold.log 2>&1; tail -n 1 gpurun_out/r2/bench_lnfold.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'],{k:(round(v['ms_per_step'],2), v['tflops'] and round(v['tflops'])) for k,v in d['kernels'].items()})"
