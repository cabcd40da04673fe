set -u
export PYTHONPATH=$PWD
OUT=gpurun_out
for k in raster_fwd_pool_kernel raster_bwd_pool_kernel; do
  ST3R_RASTER_VARIANT=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o $OUT/r02b_prof_$k python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
ls -la $OUT | tail -5
