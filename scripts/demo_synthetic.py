"""The reference's main.py sequence (main.py:48-88) on a synthetic scene, through the unchanged `starster` API:
reconstruct (MATCH + ALIGN) -> init_3dgs -> run_3dgs_optim -> render.  The MASt3R network is out of scope, so a
scene-consistent synthetic stand-in produces its outputs.  Usage: python scripts/demo_synthetic.py [n_views] [size]"""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import starster                                   # noqa: E402  (alias of starst3r_b200)
from starst3r_b200 import synth                   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
model = synth.SyntheticMast3r(n, size, size, seed=0, device=dev, arc_deg=100.0)
scene = starster.Scene(device=dev)
t0 = time.time()
scene.add_images(model, model.images())
torch.cuda.synchronize()
print(f"reconstruct: {time.time() - t0:.2f} s, {sum(p.shape[0] for p in scene.dense_pts)} dense points")
scene.init_3dgs()
t0 = time.time()
losses = scene.run_3dgs_optim(200)
torch.cuda.synchronize()
print(f"3DGS: 200 iterations in {time.time() - t0:.2f} s, loss {losses[0]:.4f} -> {losses[-1]:.4f}")
img, alpha, info = scene.render_3dgs_original(size, size)
print("render", tuple(img.shape), "intersections", info["isect_ids"].numel())
