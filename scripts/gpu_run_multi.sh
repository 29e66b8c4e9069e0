#!/bin/bash
# N GPUs of one box: config 3 (fused gather on a side stream, self-verified), optionally configs 5 / 4, the gather unit check and
# the reference arm.  Usage: gpurun --gpus N --timeout 2400 -- 'bash scripts/gpu_run_multi.sh N [3 5 4]'
mkdir -p gpurun_out
N=${1:-2}; shift
CFGS=${@:-3}
show() { tail -n 1 $1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
except Exception as e:
    print('NO JSON', e); sys.exit()
ks=sum(v['ms_per_step'] for v in d['kernels'].values())
print('config',d['config']['baseline_config'],'n',d['n_gpus'],'fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'sum_kernels',round(ks,2),'e2e',round(d['e2e']['value'],1),'verified',d.get('gather_verified'),'clk',d['clocks']['sm_mhz'])"; }
port=29511
for c in $CFGS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config $c > gpurun_out/bench_n${N}_c$c.log 2>&1; echo "config $c n$N exit $?"; show gpurun_out/bench_n${N}_c$c.log
  port=$((port+1))
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tests/mgpu_gather_check.py 2>&1 | tail -n 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((port+1)) bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "bench ref n$N exit $?"; tail -n 1 gpurun_out/bench_ref_n$N.log | cut -c1-200
