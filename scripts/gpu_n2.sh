#!/bin/bash
# 2-GPU visit: sharded + CRD tests, the step timeline, the bench line.   Usage (gpurun --gpus 2): bash scripts/gpu_n2.sh <tag>
set -u
TAG=${1:-r2}
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_crd_gpu.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_2gpu.txt
bash scripts/gpu_profile_sharded.sh $TAG 2
bash scripts/gpu_sharded.sh $TAG 2 bench
