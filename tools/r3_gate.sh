#!/bin/bash
# GEGLU gate as hv + hv tanh(u): tests, mlp micro-benchmark, bench lines, whole GPU suite
tag=${1:-r3h}
out=gpurun_out/${tag}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py -x -q -k "geglu or mlp" 2>&1 | tail -3 | tee ${out}_ktests.log
timeout 200 python tools/mlp_bench.py 2>&1 | tee ${out}_mlp.log
for v in 1 2; do
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-legs 2>&1 | tail -1 | tee -a ${out}_bench.log | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l)
    print('ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 3), d['clocks']['sm_mhz'], {k: v['ms'] for k, v in d['kernel_classes'].items()})
except Exception as e:
    print('unparsed:', l[-400:])
"
done
timeout 900 python -m pytest tests -m gpu -x -q > ${out}_tests.log 2>&1; echo "tests rc=$?" | tee -a ${out}_tests.log
tail -3 ${out}_tests.log
cp gpurun_out/parity_fullshape.jsonl ${out}_parity.jsonl
