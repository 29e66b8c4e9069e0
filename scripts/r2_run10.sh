#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_geometry_metrics.py tests/test_gpu_preprocess.py -q -x -s -p no:cacheprovider 2>&1 | grep -E "err |passed|failed|Error" | tail -n 12
for q in 2 4; do echo "QPT $q"; PROF_LIB=$PWD/gpurun_variants/libdav2_b200_knobs.so DAV2_BP_QPT=$q timeout 300 python scripts/prof_ops.py geom 1 2>&1 | grep -E "us/launch|Error|error"; done
timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/bench_bpfused3.log 2>&1; tail -n 1 gpurun_out/r2/bench_bpfused3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'],'bp',d.get('roofline_backproject'), d['kernels'].get('backproject'))"
