#!/bin/bash
# round-2 run 2: attention v3 (speculative reference max + cross-tile prefetch): parity, timing per EMU, bench
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" -p no:cacheprovider > gpurun_out/r2/pytest_attn.log 2>&1; echo "pytest attention exit $?"; tail -n 3 gpurun_out/r2/pytest_attn.log | cut -c1-300
export PROF_LIB=$PWD/gpurun_variants/libdav2_b200_knobs.so
for sc in 1.0 0.35; do for e in 0 2 3 4; do DAV2_QKV_SCALE=$sc DAV2_TIME=1 DAV2_ATTN_EMU=$e timeout 120 python scripts/prof_ops.py attn 1 2>&1 | tail -n 1 | sed "s/^/scale $sc /"; done; done
unset PROF_LIB
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_configs.py -q -p no:cacheprovider -s > gpurun_out/r2/pytest_model.log 2>&1; echo "pytest model exit $?"; grep -E "fp16:|passed|failed|Error" gpurun_out/r2/pytest_model.log | cut -c1-400 | tail -12
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_attn3.log 2>&1; tail -n 1 gpurun_out/r2/bench_attn3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'clk',d['clocks']['sm_mhz'],{k:(round(v['ms_per_step'],2), v['tflops'] and round(v['tflops'])) for k,v in d['kernels'].items()})"
