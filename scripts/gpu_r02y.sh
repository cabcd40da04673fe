#!/bin/bash
# GPU pass r02y: ncu capture (with source) of the persistent ALIGN kernel.
set -u
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
ST3R_PROF_PASSES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_persist_kernel -c 1 -f \
    -o $OUT/${TAG}_prof_align_persist python scripts/prof_align.py > $OUT/ncu_align_persist.log 2>&1
tail -2 $OUT/ncu_align_persist.log
python scripts/ncu_summary.py $OUT/${TAG}_prof_align_persist.ncu-rep | head -30
