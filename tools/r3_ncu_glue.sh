#!/bin/bash
# ncu --set full of the level-0 GroupNorm(+SiLU) launch (micro-benchmark)
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:gn_fused -s 2 -c 1 -f -o gpurun_out/r3e_gn_silu python tools/norm_bench.py > gpurun_out/r3_ncu1.log 2>&1; tail -2 gpurun_out/r3_ncu1.log
