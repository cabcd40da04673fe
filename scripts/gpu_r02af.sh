#!/bin/bash
# Pool-kernel configuration per direction: parity of the new default, A/B of library variants on the headline step.
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gs_gpu.py -m gpu -q --tb=short 2>&1 | tail -8 | tee $OUT/r02af_pytest_gs.txt
bash scripts/gpu_libvar.sh r02af_libvar "init rand"
