#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_fullshape_gpu.py -x -q -k two_step > gpurun_out/r2c_tests.log 2>&1; echo "fullshape rc=$?"
tail -3 gpurun_out/r2c_tests.log
timeout 600 python -m pytest tests/test_sharding_gpu.py -x -q > gpurun_out/r2c_shard_tests.log 2>&1; echo "sharding rc=$?"
tail -5 gpurun_out/r2c_shard_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2c_bench2.log 2>&1
tail -c 5000 gpurun_out/r2c_bench2.log
