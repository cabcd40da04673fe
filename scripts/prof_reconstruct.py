"""cProfile of reconstruct_scene on the bench's MATCH + ALIGN leg (host-side hot spots)."""
import cProfile, pstats, sys, time
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
rc.reconstruct_scene(model, imgs, names, dev)
rc._MEMO.clear()
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.time()
pr.enable()
scene, _ = rc.reconstruct_scene(model, imgs, names, dev)
scene.get_dense_pts3d(clean_depth=True)
torch.cuda.synchronize()
pr.disable()
print("seconds", time.time() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
