#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench n$N exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/bench_n$N.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "bench ref n$N exit $?" >> gpurun_out/summary.txt; tail -n 1 gpurun_out/bench_ref_n$N.log | cut -c1-300
cat gpurun_out/summary.txt
