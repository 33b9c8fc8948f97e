#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): sharded parity tests + the driver's bench command at N ranks (both arms).
# Usage:  bash scripts/gpu_sharded.sh <tag> <N> [bench]     ("bench": only the GPU arm of bench.py, for the expensive N)
set -u
TAG=${1:-r2}
N=${2:-2}
MODE=${3:-all}
OUT=gpurun_out
mkdir -p $OUT
[ "$MODE" = "bench" ] || { echo "== pytest sharded" ; timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_sharded.txt; }
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 2> $OUT/${TAG}_bench_g$N.err | tee $OUT/${TAG}_bench_g$N.json
tail -6 $OUT/${TAG}_bench_g$N.err
[ "$MODE" = "bench" ] && exit 0
echo "== reference arm N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 2> /dev/null | tee $OUT/${TAG}_bench_ref_g$N.json
