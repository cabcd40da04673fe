#!/bin/bash
# GPU pass r02w: parity suite + bench after the cp.async loss kernels, the cheaper projection-backward output pass (vs the
# serial kernel), the padding-skipping register sort and the pipelined pair loop.
set -u
TAG=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== align breakdown"
timeout 300 python scripts/prof_align.py 2>&1 | tail -1 | tee $OUT/${TAG}_align_breakdown.json
echo "== step breakdown: camera-parallel vs serial projection backward"
for s in 0 1; do
ST3R_PROJECT_BWD_SERIAL=$s timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
from starst3r_b200 import gs
dev = torch.device("cuda:0")
params, states, truth, cams = bench.make_workload(dev, 0)
plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
for i in range(25):
    gs.train_step(params, states, truth, cams, bench.W, bench.H, i + 1, plan=plan)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    gs.train_step(params, states, truth, cams, bench.W, bench.H, 26 + i, plan=plan)
e1.record(); torch.cuda.synchronize()
gs.PROF = {}
for i in range(10):
    gs.train_step(params, states, truth, cams, bench.W, bench.H, 46 + i, plan=plan)
ps = gs.prof_summary()
print("serial" if os.environ.get("ST3R_PROJECT_BWD_SERIAL") == "1" else "parallel", "ms/step (no flush)", e0.elapsed_time(e1) / 20,
      {k: round(v[1] / v[0], 4) for k, v in ps.items()})
PY
done
echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["step_breakdown_ms"])
print(d["reconstruct"]["seconds"], d["reconstruct"]["stages_s"], d["match"]["value"])
PY
tail -3 $OUT/${TAG}_bench.err
echo "== ncu captures"
for k in ssim_l1_fwd_kernel ssim_l1_bwd_kernel tile_sort_kernel gs_project_bwd_kernel; do
  ST3R_PROF_STEPS=21 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 19 -c 1 -f -o $OUT/${TAG}_prof_$k \
      python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
python scripts/ncu_summary.py $OUT/${TAG}_prof_*.ncu-rep > $OUT/${TAG}_ncu_summary.txt 2>&1
grep -E "^##|duration|issue_active|warps_active|registers" $OUT/${TAG}_ncu_summary.txt
