#!/bin/bash
# small_linear row-striding warps: whole GPU suite, smoke, bench line
out=gpurun_out/r3k
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > ${out}_tests.log 2>&1; echo "tests rc=$?" | tee -a ${out}_tests.log
tail -3 ${out}_tests.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-legs 2>&1 | tail -1 | tee -a ${out}_bench.log | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l)
    print('ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 3), d['clocks']['sm_mhz'], {k: v['ms'] for k, v in d['kernel_classes'].items()})
except Exception as e:
    print('unparsed:', l[-400:])
"
