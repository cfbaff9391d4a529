#!/bin/bash
# ncu --set full of the level-0 GroupNorm(+SiLU / plain) launch and the level-0 packed LayerNorm launch (micro-benchmarks)
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:gn_fused -s 2 -c 1 -f -o gpurun_out/r3_gn_silu python tools/norm_bench.py > gpurun_out/r3_ncu1.log 2>&1; tail -2 gpurun_out/r3_ncu1.log
ncu --set full --import-source on --clock-control none -k regex:gn_fused -s 13 -c 1 -f -o gpurun_out/r3_gn_nosilu python tools/norm_bench.py > gpurun_out/r3_ncu2.log 2>&1; tail -2 gpurun_out/r3_ncu2.log
ncu --set full --import-source on --clock-control none -k regex:layernorm_packed -s 2 -c 1 -f -o gpurun_out/r3_ln python tools/glue_bench.py > gpurun_out/r3_ncu3.log 2>&1; tail -2 gpurun_out/r3_ncu3.log
ls -la gpurun_out/*.ncu-rep
