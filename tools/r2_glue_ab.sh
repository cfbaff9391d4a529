#!/bin/bash
# A/B on one B200 (gpurun) of: LayerNorm variants, the one-MUFU SiLU of the GroupNorm apply phase, the one-MUFU GEGLU gate
# of the fused feed-forward.  Kernel tests under the candidate configuration, micro-benchmarks per variant, bench lines.
tag=r3c
out=gpurun_out/${tag}
mkdir -p gpurun_out
PT_LN_ONEPASS=1 PT_MLP_GATE_FMA=1 timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_gpu.py -x -q -k "layernorm or groupnorm or mlp or geglu" 2>&1 | tail -2 | tee -a ${out}_glue.log
for cfg in "0 0 0" "1 0 0" "1 0 1" "1 1 0" "1 1 1"; do
  set -- $cfg
  echo "== PT_LN_PACKED=$1 PT_LN_ONEPASS=$2 PT_LN_FULL=$3" | tee -a ${out}_glue.log
  PT_LN_PACKED=$1 PT_LN_ONEPASS=$2 PT_LN_FULL=$3 timeout 200 python tools/glue_bench.py 2>&1 | grep layernorm | tee -a ${out}_glue.log
done
for s in 0 1; do
  echo "== PT_GN_SILU_FMA=$s" | tee -a ${out}_glue.log
  PT_GN_SILU_FMA=$s timeout 200 python tools/norm_bench.py 2>&1 | grep groupnorm | tee -a ${out}_glue.log
  echo "== PT_MLP_GATE_FMA=$s" | tee -a ${out}_glue.log
  PT_MLP_GATE_FMA=$s timeout 200 python tools/mlp_bench.py 2>&1 | tee -a ${out}_glue.log
done
# columns: LN_ONEPASS LN_FULL GN_SILU_FMA MLP_GATE_FMA
for cfg in "0 0 0 0" "1 1 1 0" "1 1 1 1" "0 0 0 0" "1 1 1 1" "0 1 1 0"; do
  set -- $cfg
  echo "== PT_LN_ONEPASS=$1 PT_LN_FULL=$2 PT_GN_SILU_FMA=$3 PT_MLP_GATE_FMA=$4" | tee -a ${out}_bench.log
  PT_LN_ONEPASS=$1 PT_LN_FULL=$2 PT_GN_SILU_FMA=$3 PT_MLP_GATE_FMA=$4 timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-legs 2>&1 | tail -1 | tee -a ${out}_bench.log | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l)
    print('ms/step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 3), d['clocks']['sm_mhz'], {k: v['ms'] for k, v in d['kernel_classes'].items()})
except Exception as e:
    print('unparsed:', l[-400:])
"
done
