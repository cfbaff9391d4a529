#!/bin/bash
tag=${1:-r3i}
out=gpurun_out/${tag}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm" 2>&1 | tail -2 | tee ${out}_ktests.log
for v in "0 1" "1 0" "1 1"; do
set -- $v
echo "== PT_LN_PACKED=$1 PT_LN_FULL=$2" | tee -a ${out}_glue.log
PT_LN_PACKED=$1 PT_LN_FULL=$2 timeout 200 python tools/glue_bench.py 2>&1 | grep layernorm | tee -a ${out}_glue.log
done
for v in 0 1 0 1; do
echo "== PT_LN_FULL=$v" | tee -a ${out}_bench.log
PT_LN_FULL=$v timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-legs 2>&1 | tail -1 | tee -a ${out}_bench.log | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l)
    print('ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 3), d['clocks']['sm_mhz'], {k: v['ms'] for k, v in d['kernel_classes'].items()})
except Exception as e:
    print('unparsed:', l[-400:])
"
done
