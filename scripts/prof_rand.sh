set -u
export PYTHONPATH=$PWD
OUT=gpurun_out
for k in raster_fwd_pool_kernel raster_bwd_pool_kernel; do
  ST3R_SCALE_MODE=rand ST3R_PROF_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 -c 1 -f -o $OUT/r02k_prof_rand_$k python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
for k in raster_fwd_kernel raster_bwd_kernel; do
  ST3R_RASTER_VARIANT=1 ST3R_SCALE_MODE=rand ST3R_PROF_STEPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 -c 1 -f -o $OUT/r02k_prof_rand_$k python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
