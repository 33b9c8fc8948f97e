#!/bin/bash
# One GPU-box visit: the whole GPU test suite, smoke, the driver's bench command (both arms), error report.
# Usage (under gpurun, from the repo root):  bash scripts/gpu_full.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $OUT/${TAG}_smoke.txt
echo "== fusion error report" ; timeout 300 python scripts/report_fusion_errors.py 2>&1 | tail -40 | tee $OUT/${TAG}_fusion_errors.txt
echo "== reference arm" ; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> $OUT/${TAG}_bench_ref.err | tee $OUT/${TAG}_bench_ref.json
echo "== bench" ; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json
tail -5 $OUT/${TAG}_bench.err
ls -la $OUT | tail -8
