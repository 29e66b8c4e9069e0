#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_fused.log 2>&1; echo "bench n$N fused exit $?" >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tests/mgpu_gather_check.py > gpurun_out/mgpu_n${N}.log 2>&1; echo "mgpu check n$N exit $?" >> gpurun_out/summary.txt
