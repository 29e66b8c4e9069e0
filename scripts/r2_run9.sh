#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_geometry_metrics.py tests/test_gpu_preprocess.py tests/test_host_metrics.py -q -x -p no:cacheprovider 2>&1 | tail -n 6
timeout 300 python scripts/prof_ops.py geom 1 2>&1 | grep -E "us/launch|voxel|Error|error"
timeout 200 python scripts/prof_ops.py geom 1 > gpurun_out/r2/geom_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'backproject|depth_metrics' -c 8 --csv --log-file gpurun_out/r2/ncu_geom_metrics.csv python scripts/prof_ops.py geom 1 > gpurun_out/r2/ncu_geom.log 2>&1; echo "ncu exit $?"
timeout 600 python bench.py --steps 6 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/bench_bpfused2.log 2>&1; tail -n 1 gpurun_out/r2/bench_bpfused2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'],'bp',d.get('roofline_backproject'), d['kernels'].get('backproject'))"
