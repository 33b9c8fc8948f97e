#!/bin/bash
# compute-sanitizer runs of every kernel family at small shapes: logs -> gpurun_out/<tag>_sanitize_*.txt
set -u
TAG=${1:-r2}
FAMS=${2:-kron crd select}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck synccheck racecheck; do
  for fam in $FAMS; do
    echo "== $tool $fam"
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py $fam > $OUT/${TAG}_sanitize_${tool}_${fam}.txt 2>&1
    tail -4 $OUT/${TAG}_sanitize_${tool}_${fam}.txt
  done
done
