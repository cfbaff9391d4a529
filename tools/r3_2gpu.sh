#!/bin/bash
# 2-GPU check of the multi-GPU paths with this session's kernels: NCCL / P2P tests, then the bench under torchrun
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharding_gpu.py tests/test_train_step_gpu.py -x -q -m gpu -k "sharding or world2 or cfg or frame" > gpurun_out/r3_2gpu_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r3_2gpu_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/r3_bench_2gpu.log 2>&1
grep '^{' gpurun_out/r3_bench_2gpu.log | tail -1 > gpurun_out/r3_bench_2gpu.json
tail -5 gpurun_out/r3_bench_2gpu.log | cut -c1-400
