"""Development aid: save the splat geometry of the headline workload after N training steps (for offline statistics)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from starst3r_b200 import gs
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
params, states, truth, cams = bench.make_workload(dev, 0)
plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
for i in range(1, steps + 1):
    loss, fr = gs.train_step(params, states, truth, cams, bench.W, bench.H, i, plan=plan)
torch.cuda.synchronize()
C, N = fr.C, fr.N
# view 0 only: 2-D geometry of every Gaussian + the sorted per-tile lists
n = int(fr.n_isect)
keys = fr.keys[:n].cpu()
vals = fr.vals[:n].cpu()
sel = vals < N          # entries of camera 0
out = {"geomA": fr.geomA[:N].cpu().half().float().half(), "geomA32": fr.geomA[:N].cpu(), "geomB32": fr.geomB[:N].cpu(),
       "radii": fr.radii[:N].cpu(), "vals0": vals[sel].int(), "offsets0": fr.offsets[0].cpu(),
       "scales": params["scales"].cpu(), "opacities": params["opacities"].cpu(), "n_isect": n}
del out["geomA"]
torch.save(out, f"gpurun_out/state_step{steps}.pt")
print("saved", {k: (tuple(v.shape) if hasattr(v, "shape") else v) for k, v in out.items()})
