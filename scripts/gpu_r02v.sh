#!/bin/bash
# GPU pass r02v: parity suite + bench after the loss / tile-sort / projection-backward rewrites, ALIGN time breakdown,
# ncu captures of the ALIGN loss kernels and of the rewritten kernels.
set -u
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== align breakdown"
timeout 300 python scripts/prof_align.py 2>&1 | tail -1 | tee $OUT/${TAG}_align_breakdown.json
echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02v_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["step_breakdown_ms"])
print(d["reconstruct"]["seconds"], d["reconstruct"]["stages_s"], d["match"]["value"])
PY
tail -3 $OUT/${TAG}_bench.err
echo "== ncu captures"
for k in ssim_l1_fwd_kernel ssim_l1_bwd_kernel tile_sort_kernel gs_project_bwd_kernel; do
  ST3R_PROF_STEPS=21 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 19 -c 1 -f -o $OUT/${TAG}_prof_$k \
      python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
ST3R_PROF_PASSES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_loss_seg_kernel --launch-skip 499 -c 2 -f \
    -o $OUT/${TAG}_prof_align_loss_seg python scripts/prof_align.py > $OUT/ncu_align_loss_seg.log 2>&1
python scripts/ncu_summary.py $OUT/${TAG}_prof_*.ncu-rep > $OUT/${TAG}_ncu_summary.txt 2>&1
grep -E "^##|duration|issue_active|warps_active|registers" $OUT/${TAG}_ncu_summary.txt
