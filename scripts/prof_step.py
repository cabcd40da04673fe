"""Few train steps + one extract_correspondences call (for ncu launch lists / kernel captures)."""
import os
import sys
import torch
sys.path.insert(0, "/root/repo")
import bench
from starst3r_b200 import gs, match, synth
dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "both"
if what in ("both", "step"):
    params, states, truth, cams = bench.make_workload(dev, 0, scale_mode=os.environ.get("ST3R_SCALE_MODE", "init"))
    plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
    for i in range(int(os.environ.get("ST3R_PROF_STEPS", "3"))):
        loss, fr = gs.train_step(params, states, truth, cams, bench.W, bench.H, i + 1, plan=plan)
    torch.cuda.synchronize()
if what in ("both", "match"):
    A, B = synth.descriptor_pair(512, 512, seed=0, device=dev)
    A2, B2 = synth.descriptor_pair(512, 512, seed=100, device=dev)
    q = [1 + 9 * torch.rand(512, 512, device=dev) for _ in range(4)]
    for _ in range(2):
        out = match.extract_correspondences_device([A, B, B2, A2], q, 8)
    torch.cuda.synchronize()
