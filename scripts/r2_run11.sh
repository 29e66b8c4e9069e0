#!/bin/bash
mkdir -p gpurun_out/r2
N=${1:-2}
show() { tail -n 1 $1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
except Exception as e:
    print('NO JSON', e); sys.exit()
ks=sum(v['ms_per_step'] for v in d['kernels'].values())
print(d['config']['baseline_config'], 'n',d['n_gpus'],'fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'sum_kernels',round(ks,2),'e2e',round(d['e2e']['value'],1),'verified',d.get('gather_verified'),'clk',d['clocks']['sm_mhz'], 'bp ms', d['kernels'].get('backproject',{}).get('ms_per_step'), d['config'].get('gather_overlap'))"; }
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2/bench_n${N}_ov.log 2>&1; echo "overlap n$N exit $?"; show gpurun_out/r2/bench_n${N}_ov.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 --no-overlap > gpurun_out/r2/bench_n${N}_noov.log 2>&1; echo "no-overlap n$N exit $?"; show gpurun_out/r2/bench_n${N}_noov.log
timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/bench_n1_same_box.log 2>&1; echo "n1 exit $?"; tail -n 1 gpurun_out/r2/bench_n1_same_box.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n1 fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'],'bp',d.get('roofline_backproject',{}).get('frac'))"
timeout 600 python -m pytest tests/test_gpu_preprocess.py -q -s -p no:cacheprovider 2>&1 | grep -E "err |passed|failed" | tail -n 8
