#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_geometry_metrics.py -q -x -p no:cacheprovider 2>&1 | tail -n 3
for lib in knobs knobs_minb3; do for q in 2 4; do echo "== $lib QPT $q"; PROF_LIB=$PWD/gpurun_variants/libdav2_b200_$lib.so DAV2_BP_QPT=$q timeout 300 python scripts/prof_ops.py geom 1 2>&1 | grep -E "us/launch|Error|error" | grep -v metrics0; done; done
timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/bench_bp4.log 2>&1; tail -n 1 gpurun_out/r2/bench_bp4.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'],'bp',d.get('roofline_backproject',{}).get('frac'), d['kernels'].get('backproject'), 'alone', d.get('roofline_backproject_alone',{}).get('frac'))"
