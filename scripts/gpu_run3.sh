#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 --no-header -p no:cacheprovider > gpurun_out/ops.log 2>&1; echo "ops exit $?" >> gpurun_out/summary.txt; tail -n 6 gpurun_out/ops.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_geometry_metrics.py -q -m gpu --timeout 300 --no-header -p no:cacheprovider > gpurun_out/model.log 2>&1; echo "model+geom exit $?" >> gpurun_out/summary.txt; tail -n 30 gpurun_out/model.log
timeout 120 python scripts/gpu_debug_model.py vits 70 98 1 fp16 > gpurun_out/debug_fp16.log 2>&1; tail -n 4 gpurun_out/debug_fp16.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_vitl_fp16.log 2>&1; echo "bench_vitl exit $?" >> gpurun_out/summary.txt; tail -n 2 gpurun_out/bench_vitl_fp16.log
timeout 300 python scripts/prof_ops.py all 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 6 -o gpurun_out/prof_gemm_r01 python scripts/prof_ops.py all 2 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/summary.txt
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 1 -o gpurun_out/prof_attn_r01 python scripts/prof_ops.py attn 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
