#!/bin/bash
# Dropout path of the Kronecker kernels: parity tests, timed points, ncu --set full with source counters.
# Usage (under gpurun):  bash scripts/gpu_kron_drop.sh <tag> [noprof]
set -u
TAG=${1:-r2}
MODE=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest fusion + select" ; timeout 1200 python -m pytest tests/test_fusion_gpu.py tests/test_crd_select_gpu.py -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_fusion.txt
: > $OUT/${TAG}_kron.jsonl
for shp in 16384,128,128,256 8192,32,32,32,96 65536,64,64,128; do
  timeout 300 python scripts/bench_kron.py --bwd --iters 20 --dropout 0.25 --only $shp 2>&1 | tail -1 | tee -a $OUT/${TAG}_kron.jsonl
done
timeout 300 python scripts/bench_kron.py --bwd --iters 20 --dropout 0.1 --only 8192,32,32,32,96 2>&1 | tail -1 | tee -a $OUT/${TAG}_kron.jsonl
echo "== selection variant" ; timeout 300 python scripts/bench_select.py 2>&1 | grep config | tee $OUT/${TAG}_bench_select.jsonl
[ "$MODE" = "noprof" ] && exit 0
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:kron_(fwd|wgrad|dgrad)_tc_kernel' -s 3 -c 3 -f \
    -o $OUT/${TAG}_prof_kron_drop python scripts/ncu_kron.py 16384,128,128,256 0.25 > $OUT/${TAG}_ncu_kron_drop.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:kron_(fwd|wgrad|dgrad)_tc_kernel' -s 3 -c 3 -f \
    -o $OUT/${TAG}_prof_kron_drop_c4 python scripts/ncu_kron.py 8192,32,32,32,96 0.25 > $OUT/${TAG}_ncu_kron_drop_c4.log 2>&1
ls -la $OUT | tail -5
