#!/bin/bash
# Fragment-pool blend kernels on one GPU box: parity, A/B timing against the default kernels, ncu captures.
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
ST3R_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k "raster" --tb=short 2>&1 | tail -40 | tee $OUT/${TAG}_pytest_pool.txt
ST3R_VARIANTS="${VARIANTS:-0,3}" timeout 900 python scripts/bench_variants.py raster > $OUT/${TAG}_variants.json 2> $OUT/${TAG}_variants.err
cat $OUT/${TAG}_variants.json; tail -5 $OUT/${TAG}_variants.err
if [ "${NCU:-1}" = "1" ]; then
for k in raster_fwd_pool_kernel raster_bwd_pool_kernel; do
  ST3R_RASTER_VARIANT=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o $OUT/${TAG}_prof_$k python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
fi
