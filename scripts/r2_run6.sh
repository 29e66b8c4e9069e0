#!/bin/bash
# round-2 run 6: the new bench.py on one GPU, every config (short), reference arm
mkdir -p gpurun_out/r2
show() { tail -n 1 $1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
except Exception as e:
    print('NO JSON', e); sys.exit()
k={a:(round(v['ms_per_step'],2)) for a,v in d.get('kernels',{}).items()}
print(d['config']['baseline_config'], 'fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'e2e_small',round(d.get('e2e_metrics_only',{}).get('value',0),1),'clk',d['clocks']['sm_mhz'],'eager',d.get('gpu_eager_baseline'),'cpu',d.get('cpu_baseline',{}).get('value'),'bp',d.get('roofline_backproject',{}).get('frac'),'met',d.get('roofline_depth_metrics',{}).get('frac'), k)"; }
timeout 900 python bench.py > gpurun_out/r2/bench_c3.log 2>&1; echo "c3 exit $?"; show gpurun_out/r2/bench_c3.log
timeout 600 python bench.py --config 2 --steps 4 --no-cpu-baseline > gpurun_out/r2/bench_c2.log 2>&1; echo "c2 exit $?"; show gpurun_out/r2/bench_c2.log
timeout 600 python bench.py --config 5 --steps 3 --no-cpu-baseline > gpurun_out/r2/bench_c5.log 2>&1; echo "c5 exit $?"; show gpurun_out/r2/bench_c5.log
timeout 900 python bench.py --config 4 --steps 2 --no-cpu-baseline > gpurun_out/r2/bench_c4.log 2>&1; echo "c4 exit $?"; show gpurun_out/r2/bench_c4.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2/bench_ref.log 2>&1; echo "ref exit $?"; tail -n 1 gpurun_out/r2/bench_ref.log | cut -c1-200
timeout 600 python bench.py --precision bf16 --steps 4 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/bench_c3_bf16.log 2>&1; echo "bf16 exit $?"; show gpurun_out/r2/bench_c3_bf16.log
