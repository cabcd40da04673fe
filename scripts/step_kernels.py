"""Per-entry-point times (CUDA events, gs.PROF) of the headline training step after 20 warm-up iterations, for A/B runs
of library variants: `ST3R_B200_LIB=<lib> python scripts/step_kernels.py [init|rand]` prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from starst3r_b200 import gs  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "init"
gs.TRAIN_GRAPH = False
dev = torch.device("cuda:0")
params, states, truth, cams = bench.make_workload(dev, 0, scale_mode=mode)
plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
for i in range(20):
    gs.train_step(params, states, truth, cams, bench.W, bench.H, i + 1, plan=plan)
torch.cuda.synchronize()
gs.PROF = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(10):
    loss, fr = gs.train_step(params, states, truth, cams, bench.W, bench.H, 21 + i, plan=plan)
p = gs.prof_summary()
out = {"lib": os.path.basename(os.environ.get("ST3R_B200_LIB", "default")), "scales": mode, "loss": round(float(loss), 6)}
out.update({k.replace("st3r_gs_", ""): round(v[1] / v[0], 4) for k, v in p.items()})
out["sum"] = round(sum(v[1] / v[0] for v in p.values()), 4)
print(json.dumps(out))
