#!/bin/bash
# K13 class k-means: tests, timing, sanitizers, ncu --set full of the assign pass.   Usage (under gpurun): bash scripts/gpu_kmeans.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest kmeans + knn module goldens"
timeout 600 python -m pytest tests/test_crd_kmeans_gpu.py tests/test_crd_knn_gpu.py -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_kmeans.txt
echo "== bench kmeans" ; timeout 600 python scripts/bench_kmeans.py --cpu 2>&1 | grep class_kmeans | tee $OUT/${TAG}_bench_kmeans.jsonl
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py kmeans > $OUT/${TAG}_sanitize_${tool}_kmeans.txt 2>&1
  echo "== $tool"; tail -3 $OUT/${TAG}_sanitize_${tool}_kmeans.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:kmeans_' -s 6 -c 4 -f \
    -o $OUT/${TAG}_prof_kmeans python scripts/bench_kmeans.py > $OUT/${TAG}_ncu_kmeans.log 2>&1
ls -la $OUT | tail -4
