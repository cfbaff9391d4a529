#!/bin/bash
# frame sharding on N GPUs: parity tests, then ms/step with the NCCL exchange (PT_P2P=0) and the fused P2P exchange
N=${1:-2}; tag=${2:-p2p}; F=${3:-25}; H=${4:-72}; W=${5:-128}
out=gpurun_out/${tag}
if [ "$N" = "2" ]; then
  timeout 500 python -m pytest tests/test_sharding_gpu.py -x -q > ${out}_tests.log 2>&1; echo "tests rc=$?"; tail -15 ${out}_tests.log
fi
for p in 0 1; do
  echo "== PT_P2P=$p" | tee -a ${out}_run.log
  PT_P2P=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + p)) \
      tools/run_sharded.py $F $H $W 6 2>&1 | grep -v "^W\|^\[W\|NCCL version\|^$" | tail -8 | tee -a ${out}_run.log
done
