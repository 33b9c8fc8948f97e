#!/bin/bash
# KNN positives: tests, timing, ncu --set full of the tcgen05 pass.   Usage (under gpurun): bash scripts/gpu_knn.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest knn" ; timeout 600 python -m pytest tests/test_crd_knn_gpu.py -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_knn.txt
echo "== bench knn" ; timeout 600 python scripts/bench_knn.py 2>&1 | grep workload | tee $OUT/${TAG}_bench_knn.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:knn_' -s 10 -c 5 -f \
    -o $OUT/${TAG}_prof_knn python scripts/bench_knn.py --cpu-anchors 0 --iters 2 > $OUT/${TAG}_ncu_knn.log 2>&1
ls -la $OUT | tail -4
