#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_fullshape.jsonl gpurun_out/parity_errors.jsonl
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_fullshape_gpu.py -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2b_bench.log 2>&1
tail -c 6000 gpurun_out/r2b_bench.log
