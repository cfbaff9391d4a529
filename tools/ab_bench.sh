#!/bin/bash
# A/B of the PDL launches and the smem-resident GroupNorm on one B200 (gpurun): tests first, then bench lines.
tag=${1:-ab}
out=gpurun_out/${tag}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > ${out}_tests.log 2>&1; echo "tests rc=$?" | tee -a ${out}_tests.log
tail -3 ${out}_tests.log
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  echo "== PT_PDL=$1 PT_GN_RESIDENT=$2" | tee -a ${out}_bench.log
  PT_PDL=$1 PT_GN_RESIDENT=$2 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee -a ${out}_bench.log | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l)
    print('ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 3), {k: v['ms'] for k, v in d['kernel_classes'].items()})
except Exception as e:
    print('unparsed:', l[-400:])
"
done
for r in 0 1; do
  echo "== norm_bench PT_GN_RESIDENT=$r" | tee -a ${out}_norm.log
  PT_GN_RESIDENT=$r timeout 200 python tools/norm_bench.py 2>&1 | tee -a ${out}_norm.log
done
