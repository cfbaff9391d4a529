#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "fused or geglu" > gpurun_out/r2d_mlp_tests.log 2>&1; echo "mlp tests rc=$?"
tail -15 gpurun_out/r2d_mlp_tests.log
timeout 200 python tools/mlp_bench.py 2>&1 | tee gpurun_out/r2d_mlp_bench.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_fullshape_gpu.py -x -q > gpurun_out/r2d_model_tests.log 2>&1; echo "model tests rc=$?"
tail -5 gpurun_out/r2d_model_tests.log
for f in 1 0; do
  PT_FUSED_MLP=$f timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-legs 2>&1 | tail -1 > gpurun_out/r2d_bench_fused$f.json
  python -c "
import json
d=json.load(open('gpurun_out/r2d_bench_fused$f.json'))
print('fused=$f', d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernel_classes'].items()})"
done
