#!/bin/bash
# A/B of library variants (starst3r_b200/libst3r_var_*.so + the default) on the headline training step.
set -u
OUT=gpurun_out
TAG=${1:-libvar}
MODES=${2:-init}
mkdir -p $OUT
export PYTHONPATH=$PWD
: > $OUT/${TAG}.jsonl
for m in $MODES; do
  timeout 300 python scripts/step_kernels.py $m 2>&1 | tail -1 | tee -a $OUT/${TAG}.jsonl
  for lib in starst3r_b200/libst3r_var_*.so; do
    ST3R_B200_LIB=$PWD/$lib timeout 300 python scripts/step_kernels.py $m 2>&1 | tail -1 | tee -a $OUT/${TAG}.jsonl
  done
done
