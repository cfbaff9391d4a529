#!/bin/bash
# round-2 ncu evidence: launch list with DRAM bytes of one eager step, then --set full of the first level-0 launches
mkdir -p gpurun_out
PT_OPLIST=gpurun_out/r2f_oplist.json ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/r2f_ncu.csv python tools/profile_step.py > gpurun_out/r2f_prof.log 2>&1
tail -2 gpurun_out/r2f_prof.log
python tools/aggregate_traffic.py gpurun_out/r2f_ncu.csv gpurun_out/r2f_oplist.json gpurun_out/r2f_dram_traffic.json gpurun_out/r2f_launches.csv
python tools/join_launches.py gpurun_out/r2f_launches.csv gpurun_out/r2f_oplist.json gpurun_out/r2f_step_launches.md > /dev/null
head -14 gpurun_out/r2f_step_launches.md
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"gemm_tcgen05|mlp_geglu" -s 1 -c 14 \
    -o gpurun_out/r2f_gemm_l0 python tools/profile_step.py > gpurun_out/r2f_full.log 2>&1
tail -2 gpurun_out/r2f_full.log
