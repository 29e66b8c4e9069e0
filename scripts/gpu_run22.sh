#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_geometry_metrics.py -x -q -m gpu > gpurun_out/pytest_a.log 2>&1; echo "pytest geom exit $?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2_fused.log 2>&1; echo "bench n2 fused exit $?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --gather nccl > gpurun_out/bench_n2_nccl.log 2>&1; echo "bench n2 nccl exit $?" >> gpurun_out/summary.txt
