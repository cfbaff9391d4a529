#!/bin/bash
# GroupNorm SiLU as h + h tanh(h) (one MUFU) A/B: kernel tests, micro-benchmark, bench lines, whole GPU suite with it on
tag=${1:-r3g}
out=gpurun_out/${tag}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -k groupnorm 2>&1 | tail -2 | tee ${out}_ktests.log
for v in 0 1; do
echo "== PT_GN_SILU_TANH=$v" | tee -a ${out}_norm.log
PT_GN_SILU_TANH=$v timeout 200 python tools/norm_bench.py 2>&1 | grep "groupnorm.*silu=True" | tee -a ${out}_norm.log
done
for v in 0 1 0 1; do
echo "== PT_GN_SILU_TANH=$v" | tee -a ${out}_bench.log
PT_GN_SILU_TANH=$v timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-legs 2>&1 | tail -1 | tee -a ${out}_bench.log | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l)
    print('ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 3), d['clocks']['sm_mhz'], {k: v['ms'] for k, v in d['kernel_classes'].items()})
except Exception as e:
    print('unparsed:', l[-400:])
"
done
cp gpurun_out/parity_fullshape.jsonl ${out}_parity_before.jsonl 2>/dev/null
rm -f gpurun_out/parity_fullshape.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > ${out}_tests.log 2>&1; echo "tests rc=$?" | tee -a ${out}_tests.log
tail -3 ${out}_tests.log
cp gpurun_out/parity_fullshape.jsonl ${out}_parity_tanh.jsonl
