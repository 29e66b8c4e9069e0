#!/bin/bash
# round-2 run 7: 2 GPUs -- config 3 (fused gather + verification), config 4 (sharded reconstruction), gather unit check
mkdir -p gpurun_out/r2
N=${1:-2}
show() { tail -n 1 $1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
except Exception as e:
    print('NO JSON', e); sys.exit()
print(d['config']['baseline_config'], 'n',d['n_gpus'],'fps',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'verified',d.get('gather_verified'),'clk',d['clocks']['sm_mhz'], 'bp ms', d['kernels'].get('backproject',{}).get('ms_per_step'), d['config'].get('gather_note'))"; }
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2/bench_n${N}_c3.log 2>&1; echo "c3 n$N exit $?"; show gpurun_out/r2/bench_n${N}_c3.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 4 --steps 2 > gpurun_out/r2/bench_n${N}_c4.log 2>&1; echo "c4 n$N exit $?"; show gpurun_out/r2/bench_n${N}_c4.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_gather_check.py 2>&1 | tail -n 2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2/bench_n${N}_ref.log 2>&1; echo "ref n$N exit $?"; tail -n 1 gpurun_out/r2/bench_n${N}_ref.log | cut -c1-160
