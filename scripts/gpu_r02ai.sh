#!/bin/bash
# Loop-based register sort network: parity (binning both variants, merge), timing of the merge and of the binning.
set -u
OUT=gpurun_out
TAG=${1:-r02ai}
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gs_gpu.py tests/test_match_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
timeout 300 python scripts/merge_variants.py 2>&1 | tail -1 | tee $OUT/${TAG}_merge.json
bash scripts/gpu_libvar.sh ${TAG}_libvar "init rand"
