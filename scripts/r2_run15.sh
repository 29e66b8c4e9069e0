#!/bin/bash
# round-2 run 15: 8 GPUs -- configs 3 / 5 / 4 under torchrun, 1-GPU config 3 on the same box for the efficiency
mkdir -p gpurun_out/r2
N=${1:-8}
show() { tail -n 1 $1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
except Exception as e:
    print('NO JSON', e); sys.exit()
ks=sum(v['ms_per_step'] for v in d['kernels'].values())
print('config',d['config']['baseline_config'], 'n',d['n_gpus'],'fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'sum_kernels',round(ks,2),'e2e',round(d['e2e']['value'],1),'verified',d.get('gather_verified'),'clk',d['clocks']['sm_mhz'], 'bp ms', d['kernels'].get('backproject',{}).get('ms_per_step'))"; }
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N ${@:3} > gpurun_out/r2/$2 2>&1; echo "$2 exit $?"; show gpurun_out/r2/$2; }
run 29511 bench_n${N}_c3.log --steps 8
timeout 600 python bench.py --steps 8 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/bench_n1_on_n${N}box.log 2>&1; echo "n1 exit $?"; show gpurun_out/r2/bench_n1_on_n${N}box.log
run 29512 bench_n${N}_c5.log --config 5 --steps 4
run 29513 bench_n${N}_c4.log --config 4 --steps 3
