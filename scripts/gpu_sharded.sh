#!/bin/bash
# Multi-GPU visit: sharded parity tests + the weak-scaling bench at the box's GPU count.
# Usage (under gpurun --gpus N):  bash scripts/gpu_sharded.sh <tag> <N> [steps]
set -u
TAG=${1:-r1}; N=${2:-2}; STEPS=${3:-50}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest peer/sharded"; timeout 900 python -m pytest tests/test_crd_gpu.py tests/test_sharded_gpu.py -m gpu -x -q -k "peer or sharded" 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_sharded.txt
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus $N --steps $STEPS --warmup 5 2>&1 | tail -4 | tee $OUT/${TAG}_bench_g$N.json
