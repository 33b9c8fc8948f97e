#!/bin/bash
# Usage (gpurun --gpus N): bash scripts/gpu_profile_sharded.sh <tag> <N>
TAG=${1:-r2}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    scripts/profile_sharded_graph.py 2> gpurun_out/${TAG}_timeline_g$N.err | tee gpurun_out/${TAG}_timeline_g$N.txt | tail -5
