#!/bin/bash
# ControlNet || UNet-encoder on two streams inside the captured step: tests, then A/B bench lines (one B200)
out=gpurun_out/r2s
timeout 900 python -m pytest tests -m gpu -x -q > ${out}_tests.log 2>&1; echo "tests rc=$?" | tee -a ${out}_tests.log
tail -3 ${out}_tests.log
for c in 0 1 0 1; do
  echo "== PT_TWO_STREAM=$c" | tee -a ${out}_bench.log
  PT_TWO_STREAM=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-legs 2>&1 | tail -1 | tee -a ${out}_bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip())
print('ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 3), 'videos/min', round(d['e2e']['videos_per_min'], 2), 'clk', d['clocks']['sm_mhz'])
"
done
