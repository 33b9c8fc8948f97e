#!/bin/bash
# One GPU-box visit for the Kronecker kernels: parity tests, the config-3/4 sweep points, ncu --set full of K1/K2/K3.
# Usage (under gpurun, from the repo root):  bash scripts/gpu_kron.sh <tag> [quick]
set -u
TAG=${1:-r2}
MODE=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest fusion" ; timeout 1200 python -m pytest tests/test_fusion_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_fusion.txt
echo "== sweep points (fwd, wgrad, dgrad timed separately)"
: > $OUT/${TAG}_kron.jsonl
for shp in 16384,128,128,256 65536,128,128,256 65536,64,64,128 65536,128,128,64 65536,32,32,64 8192,32,32,32,96; do
  timeout 300 python scripts/bench_kron.py --bwd --iters 20 --only $shp 2>&1 | tail -1 | tee -a $OUT/${TAG}_kron.jsonl
done
for shp in 16384,128,128,256 8192,32,32,32,96; do
  timeout 300 python scripts/bench_kron.py --bwd --iters 20 --dropout 0.25 --only $shp 2>&1 | tail -1 | tee -a $OUT/${TAG}_kron.jsonl
done
[ "$MODE" = "quick" ] && exit 0
echo "== ncu full K1/K2/K3"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:kron_(fwd|wgrad|dgrad)_tc_kernel' -s 3 -c 3 -f \
    -o $OUT/${TAG}_prof_kron python scripts/ncu_kron.py 16384,128,128,256 > $OUT/${TAG}_ncu_kron.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:kron_(fwd|wgrad|dgrad)_tc_kernel' -s 3 -c 3 -f \
    -o $OUT/${TAG}_prof_kron_c4 python scripts/ncu_kron.py 8192,32,32,32,96 > $OUT/${TAG}_ncu_kron_c4.log 2>&1
for shp in 65536,64,64,128 65536,128,128,64; do
  tagn=$(echo $shp | tr ',' '_')
  timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:kron_(fwd|wgrad|dgrad)_tc_kernel' -s 3 -c 3 -f \
      -o $OUT/${TAG}_prof_kron_$tagn python scripts/ncu_kron.py $shp > $OUT/${TAG}_ncu_kron_$tagn.log 2>&1
done
ls -la $OUT | tail -8
