#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch lists and full captures of the top kernels.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench"
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 6000 $OUT/${TAG}_bench.json
tail -5 $OUT/${TAG}_bench.err
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_train_step.csv \
    python scripts/prof_step.py step > $OUT/ncu_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches_match_pair.csv \
    python scripts/prof_step.py match > $OUT/ncu_match.log 2>&1
echo "== ncu full captures"
for k in raster_bwd_kernel raster_fwd_kernel tile_sort_kernel ssim_l1_fwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o $OUT/${TAG}_prof_$k \
      python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_tc_kernel --launch-skip 1 -c 1 -f -o $OUT/${TAG}_prof_nn_tc_M4096 \
    python scripts/prof_nn.py tcgen05 4096 > $OUT/ncu_nn_tc.log 2>&1
ls -la $OUT
