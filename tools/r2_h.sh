#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_fullshape.jsonl gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2h_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2h_bench.log 2>&1
tail -c 3000 gpurun_out/r2h_bench.log
