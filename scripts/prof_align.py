"""Where the ALIGN stage of reconstruct_scene spends its time (8 views 512 x 512, the bench's MATCH + ALIGN leg):
host-side problem flattening vs the two optimiser phases on the device (CUDA events), second (warm) pass.
ST3R_PROF_PASSES=1 runs a single pass (for ncu captures)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, "/root/repo")
import bench
from starst3r_b200 import reconstruct as rc, synth

dev = torch.device("cuda:0")
n = bench.N_VIEWS
net = synth.SyntheticMast3r(n, bench.W, bench.H, seed=0, device="cpu", arc_deg=120.0)
imgs = net.images()
model = bench._CachedNet(net, imgs, dev)
names = [f"{i}.png" for i in range(n)]
stats = {}
_flat, _phase = rc.flatten_problem, rc._optimize_phase


def flat(*a, **k):
    torch.cuda.synchronize()
    t0 = time.time()
    r = _flat(*a, **k)
    torch.cuda.synchronize()
    stats.setdefault("flatten_problem_s", []).append(time.time() - t0)
    t = r[0]
    stats["entries"] = {"n3": int(t["e3_a1"].numel()), "n2": int(t["e2_img1"].numel()), "nd": int(t["ed_a1"].numel()),
                        "anchors": int(t["anc_img"].numel()), "core": int(t["core"].numel())}
    return r


def phase(t, meta, params, mode, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    r = _phase(t, meta, params, mode, *a, **k)
    e1.record()
    torch.cuda.synchronize()
    stats.setdefault(f"phase{mode}_wall_s", []).append(time.time() - t0)
    stats.setdefault(f"phase{mode}_gpu_ms", []).append(e0.elapsed_time(e1))
    return r


rc.flatten_problem, rc._optimize_phase = flat, phase
for p in range(int(os.environ.get("ST3R_PROF_PASSES", "2"))):
    rc.clear_cache()
    torch.cuda.synchronize()
    t0 = time.time()
    scene, _ = rc.reconstruct_scene(model, imgs, names, dev)
    torch.cuda.synchronize()
    stats.setdefault("reconstruct_scene_s", []).append(time.time() - t0)
print(json.dumps(stats))
