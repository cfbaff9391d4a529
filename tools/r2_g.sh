#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py tests/test_fullshape_gpu.py -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2g_tests.log
timeout 200 python tools/mlp_bench.py 2>&1 | tee gpurun_out/r2g_mlp_bench.log
timeout 100 python tools/mlp_trace.py 2>&1 | tail -12
for f in 1 0; do
  PT_FUSED_MLP=$f timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-legs 2>&1 | tail -1 > gpurun_out/r2g_bench_fused$f.json
  python -c "
import json
d=json.load(open('gpurun_out/r2g_bench_fused$f.json'))
print('fused=$f', d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernel_classes'].items()})"
done
