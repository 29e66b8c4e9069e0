#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_geometry_metrics.py -q -x -p no:cacheprovider 2>&1 | tail -n 4
timeout 300 python scripts/prof_ops.py geom 1 2>&1 | grep -E "us/launch|voxel|Error|error" 
timeout 200 python scripts/prof_ops.py geom 1 > gpurun_out/r2/geom_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'backproject|depth_metrics' -c 4 -o gpurun_out/r2/prof_geom python scripts/prof_ops.py geom 1 > gpurun_out/r2/ncu_geom.log 2>&1; echo "ncu exit $?"
timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/bench_bpfused.log 2>&1; tail -n 1 gpurun_out/r2/bench_bpfused.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'],'bp',d.get('roofline_backproject'), d['kernels'].get('backproject'), d['metrics'])"
