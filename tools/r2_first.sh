#!/bin/bash
# round-2 first GPU call: whole GPU suite (incl. the full-shape parity tests), then the 16-warp epilogue A/B
mkdir -p gpurun_out
rm -f gpurun_out/parity_fullshape.jsonl gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -30 gpurun_out/r2a_tests.log
for w in 0 10 20; do
  echo "== PT_EPI16=$w"
  PT_EPI16=$w timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2a_epi16_$w.json
  python -c "
import json
d=json.load(open('gpurun_out/r2a_epi16_$w.json'))
print(d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernel_classes'].items()})"
done
