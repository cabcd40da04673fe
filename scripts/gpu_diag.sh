#!/bin/bash
# Diagnostic GPU pass: ncu --set full captures (with source) of the ALIGN kernels inside a warm reconstruct_scene and of
# the secondary kernels of the training step (projection backward, tile emit / hist, Adam).
# Usage (repo root, GPU box): bash scripts/gpu_diag.sh [tag]
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
for k in "align_loss_seg_kernel<0>:700" "align_loss_seg_kernel<1>:300" "align_cam_bwd_kernel:1000" "align_cam_fwd_kernel:1000"; do
  name=${k%%:*}; skip=${k##*:}
  short=$(echo $name | tr -d '<>')
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$(echo $name | sed 's/[<>]/./g')" --launch-skip $skip -c 1 -f \
      -o $OUT/${TAG}_prof_$short python scripts/prof_reconstruct.py > $OUT/ncu_$short.log 2>&1
  tail -2 $OUT/ncu_$short.log
done
for k in gs_project_bwd_kernel tile_emit_kernel tile_hist_kernel adam_kernel gs_project_kernel; do
  ST3R_PROF_STEPS=21 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 19 -c 1 -f -o $OUT/${TAG}_prof_$k \
      python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
python scripts/ncu_summary.py $OUT/${TAG}_prof_*.ncu-rep > $OUT/${TAG}_ncu_summary.txt 2>&1
cat $OUT/${TAG}_ncu_summary.txt | grep -E "^##|kernel:|duration|dram__bytes|issue_active|warps_active|registers|grid_size"
