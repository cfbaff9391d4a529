#!/bin/bash
# end-of-session check: whole GPU suite, smoke(), bench line without the extra legs
tag=${1:-r3j}
out=gpurun_out/${tag}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > ${out}_tests.log 2>&1; echo "tests rc=$?" | tee -a ${out}_tests.log
tail -3 ${out}_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee ${out}_smoke.log
timeout 200 python tools/glue_bench.py 2>&1 | tee ${out}_glue.log | tail -3
timeout 200 python tools/norm_bench.py 2>&1 | tee ${out}_norm.log | tail -3
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
