#!/bin/bash
tag=${1:-epi}
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -2
for d in 1 2; do
  echo "== PT_EPI_DEPTH=$d"
  PT_EPI_DEPTH=$d timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/${tag}_d$d.json
  python -c "
import json
d=json.load(open('gpurun_out/${tag}_d$d.json'))
print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernel_classes'].items()})"
  PT_EPI_DEPTH=$d timeout 100 python tools/gemm_bench.py 2>&1 | tail -12
done
