#!/bin/bash
# A/B of library variants on the headline step + the parity tests that cover what changed.
set -u
OUT=gpurun_out
TAG=${1:-r02af}
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gs_gpu.py -m gpu -q --tb=short 2>&1 | tail -8 | tee $OUT/${TAG}_pytest_gs.txt
bash scripts/gpu_libvar.sh ${TAG}_libvar "${2:-init}"
