#!/bin/bash
# Opt-in kernel variants on one GPU box: parity first, then A/B timings against the default kernels.
# Usage (repo root, on the GPU box): bash scripts/gpu_variants.sh [tag]       (~6 GPU-minutes)
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== experimental parity tests"
ST3R_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_experimental_gpu.py -m gpu -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_experimental.txt
echo "== variant A/B"
timeout 900 python scripts/bench_variants.py > $OUT/${TAG}_variants.json 2> $OUT/${TAG}_variants.err
cat $OUT/${TAG}_variants.json; tail -5 $OUT/${TAG}_variants.err
