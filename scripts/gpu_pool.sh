#!/bin/bash
# Blend kernels on one GPU box: parity (both kernel pairs), A/B timing in both scale regimes, per-step evolution,
# ncu captures of the 20th training step.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gs_gpu.py -m gpu -q --tb=short 2>&1 | tail -30 | tee $OUT/${TAG}_pytest_gs.txt
ST3R_VARIANTS="${VARIANTS:-1,0}" timeout 900 python scripts/bench_variants.py raster > $OUT/${TAG}_variants.json 2> $OUT/${TAG}_variants.err
cat $OUT/${TAG}_variants.json; tail -5 $OUT/${TAG}_variants.err
timeout 600 python scripts/diag_steps.py > $OUT/${TAG}_diag.json 2> $OUT/${TAG}_diag.err; tail -3 $OUT/${TAG}_diag.err
if [ "${NCU:-1}" = "1" ]; then
for k in raster_fwd_pool_kernel raster_bwd_pool_kernel; do
  ST3R_PROF_STEPS=21 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 19 -c 1 -f -o $OUT/${TAG}_prof_step20_$k python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
fi
