#!/bin/bash
# A/B of matcher builds on the GPU: every starst3r_b200/libst3r_nn_*.so + the default library.
set -u
OUT=gpurun_out
TAG=${1:-nnvar}
mkdir -p $OUT
export PYTHONPATH=$PWD
: > $OUT/${TAG}.jsonl
timeout 300 python scripts/nn_variants.py 2>&1 | tail -1 | tee -a $OUT/${TAG}.jsonl
for lib in starst3r_b200/libst3r_nn_*.so; do
  ST3R_B200_LIB=$PWD/$lib timeout 300 python scripts/nn_variants.py 2>&1 | tail -1 | tee -a $OUT/${TAG}.jsonl
done
