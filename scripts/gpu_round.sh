#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full captures of the gather and relation kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag> [prof]     ("prof": only the ncu part)
set -u
TAG=${1:-r1}
MODE=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ "$MODE" != "prof" ]; then
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $OUT/${TAG}_smoke.txt
echo "== bench" ; timeout 600 python bench.py --steps 100 --warmup 10 2>&1 | tail -3 | tee $OUT/${TAG}_bench.json
fi
echo "== selection variant" ; timeout 300 python scripts/bench_select.py 2>&1 | grep config | tee $OUT/${TAG}_bench_select.jsonl
echo "== ncu launches (timed region only: cudaProfilerStart/Stop around the K steps)"
MML_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-kron > $OUT/${TAG}_ncu_launch.log 2>&1
echo "== ncu full: gather"
MML_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:crd_gather_kernel -c 2 -f -o $OUT/${TAG}_prof_crd python bench.py --steps 2 --warmup 3 --no-cpu --no-kron > $OUT/${TAG}_ncu_full.log 2>&1
echo "== ncu full: relation"
MML_SELECT_ONLY=big timeout 600 ncu --set full --clock-control none --import-source on -k regex:crd_relation_kernel -s 4 -c 1 -f \
    -o $OUT/${TAG}_prof_rel python scripts/bench_select.py > $OUT/${TAG}_ncu_rel.log 2>&1
ls -la $OUT | tail -14
