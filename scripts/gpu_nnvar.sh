#!/bin/bash
# A/B of matcher builds on the GPU: every starst3r_b200/libst3r_var_*.so + the default library (dbg* builds: cycle counters).
set -u
OUT=gpurun_out
TAG=${1:-nnvar}
mkdir -p $OUT
export PYTHONPATH=$PWD
: > $OUT/${TAG}.jsonl
timeout 300 python scripts/nn_variants.py 2>&1 | tail -1 | tee -a $OUT/${TAG}.jsonl
for lib in starst3r_b200/libst3r_var_*.so; do
  case $lib in
    *dbg*) echo "== $lib"; ST3R_B200_LIB=$PWD/$lib timeout 300 python scripts/dbg_tc_cycles.py 2>&1 | tail -2 ;;
    *) ST3R_B200_LIB=$PWD/$lib timeout 300 python scripts/nn_variants.py 2>&1 | tail -1 | tee -a $OUT/${TAG}.jsonl ;;
  esac
done
