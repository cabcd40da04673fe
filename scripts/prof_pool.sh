set -u
export PYTHONPATH=$PWD
OUT=gpurun_out
TAG=${1:-r02g}
# the 20th training step: --launch-skip counts launches of the matching kernel
for k in raster_fwd_pool_kernel raster_bwd_pool_kernel; do
  ST3R_PROF_STEPS=21 ST3R_RASTER_VARIANT=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 19 -c 1 -f -o $OUT/${TAG}_prof_step20_$k python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
for k in raster_fwd_kernel raster_bwd_kernel; do
  ST3R_PROF_STEPS=21 ST3R_RASTER_VARIANT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 19 -c 1 -f -o $OUT/${TAG}_prof_step20_$k python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
ls -la $OUT | tail -5
