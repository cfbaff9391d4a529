#!/bin/bash
# end-of-session evidence: ncu launch list with DRAM bytes of one eager denoise step (r3z), then the default bench line
mkdir -p gpurun_out
PT_OPLIST=gpurun_out/r3z_oplist.json ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/r3z_ncu.csv python tools/profile_step.py > gpurun_out/r3z_prof.log 2>&1
tail -2 gpurun_out/r3z_prof.log
python tools/aggregate_traffic.py gpurun_out/r3z_ncu.csv gpurun_out/r3z_oplist.json gpurun_out/r3z_dram_traffic.json gpurun_out/r3z_launches.csv
python tools/join_launches.py gpurun_out/r3z_launches.csv gpurun_out/r3z_oplist.json gpurun_out/r3z_step_launches.md > /dev/null
head -14 gpurun_out/r3z_step_launches.md
cp gpurun_out/r3z_dram_traffic.json profiles/r3z_dram_traffic.json
( time timeout 900 python bench.py ) > gpurun_out/r3_bench_1gpu.log 2>&1
tail -4 gpurun_out/r3_bench_1gpu.log | cut -c1-600
grep '^{' gpurun_out/r3_bench_1gpu.log | tail -1 > gpurun_out/r3_bench_1gpu.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r3_bench_ref.log 2>&1
tail -4 gpurun_out/r3_bench_ref.log | cut -c1-600
