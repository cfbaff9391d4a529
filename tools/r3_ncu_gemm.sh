#!/bin/bash
# ncu --set full of the first 14 tensor-core launches of the end-of-round step (level 0 of the ControlNet)
mkdir -p gpurun_out
PT_OPLIST=gpurun_out/r3y_oplist.json ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"gemm_tcgen05|mlp_geglu" -s 1 -c 14 \
    -f -o gpurun_out/r3y_gemm_l0 python tools/profile_step.py > gpurun_out/r3y_full.log 2>&1
tail -2 gpurun_out/r3y_full.log
ncu -i gpurun_out/r3y_gemm_l0.ncu-rep --page raw --csv > gpurun_out/r3y_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/r3y_raw.csv gpurun_out/r3y_oplist.json 1 | tee gpurun_out/r3y_gemm_full_table.md
