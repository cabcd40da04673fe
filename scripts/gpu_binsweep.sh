#!/bin/bash
# Sweep of the binning kernels' entries per CTA (grid size of tile_hist / tile_emit) on the headline step.
set -u
export PYTHONPATH=$PWD
for e in 8192 4096 2048 1024 16384; do
  echo "entries $e"
  ST3R_BIN_ENTRIES=$e ST3R_TRAIN_GRAPH=0 timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from starst3r_b200 import gs
dev = torch.device("cuda:0")
params, states, truth, cams = bench.make_workload(dev, 0)
plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
for i in range(20):
    gs.train_step(params, states, truth, cams, bench.W, bench.H, i + 1, plan=plan)
torch.cuda.synchronize()
gs.PROF = {}
for i in range(10):
    gs.train_step(params, states, truth, cams, bench.W, bench.H, 21 + i, plan=plan)
p = gs.prof_summary()
print({k: round(v[1] / v[0], 4) for k, v in p.items() if "bin" in k or "raster_fwd" in k})
PY
done
