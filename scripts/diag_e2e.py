"""Diagnostics: (1) where the e2e train-step loop loses time against the device-resident loop (uploads on / off, sizes);
(2) run-to-run determinism of the correspondence lists (5 passes of forward_mast3r, plain / split matcher)."""
import sys
import time

import torch

sys.path.insert(0, "/root/repo")
import bench
from starst3r_b200 import gs, match, reconstruct as rc, synth

dev = torch.device("cuda:0")
params, states, truth, cams = bench.make_workload(dev, 0)
plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
for i in range(25):
    gs.train_step(params, states, truth, cams, bench.W, bench.H, i + 1, plan=plan)
torch.cuda.synchronize()
truth_host = truth.cpu().pin_memory()
truth_dev = [torch.empty_like(truth), torch.empty_like(truth)]
copy_stream = torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream()
step = [26]


def loop(n, upload_frac, read_loss):
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = torch.zeros(n, dtype=torch.float32).pin_memory()
    nel = int(truth_host.numel() * upload_frac)

    def upload(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            if nel:
                truth_dev[b].view(-1)[:nel].copy_(truth_host.view(-1)[:nel], non_blocking=True)
            ready[b].record(copy_stream)
    for b in (0, 1):
        freed[b].record(main)
    torch.cuda.synchronize()
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    copy_stream.wait_event(e0)
    upload(0)
    for i in range(n):
        b = i & 1
        if i + 1 < n:
            upload(1 - b)
        main.wait_event(ready[b])
        loss, _ = gs.train_step(params, states, truth_dev[b] if upload_frac == 1.0 else truth, cams, bench.W, bench.H, step[0], plan=plan)
        step[0] += 1
        freed[b].record(main)
        if read_loss:
            loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
    e1.record()
    host_ms = (time.time() - t0) * 1e3 / n
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4), round(host_ms, 4)


for n in (10, 40):
    for frac, rd in ((0.0, False), (0.0, True), (0.1, True), (1.0, True)):
        print(f"steps {n} upload_frac {frac} read_loss {rd}: (gpu ms/step, host enqueue ms/step) =", loop(n, frac, rd))

# ---- determinism of the correspondences
n = bench.N_VIEWS
net = synth.SyntheticMast3r(n, bench.W, bench.H, seed=0, device="cpu", arc_deg=120.0)
imgs = net.images()
model = bench._CachedNet(net, imgs, dev)
names = [f"{i}.png" for i in range(n)]
pairs_in = rc.convert_dust3r_pairs_naming(names, rc.make_pairs(rc.prepare_images_for_mast3r(imgs), "complete", None, True))
ref = None
for tag, split, pipe in (("auto", "auto", True), ("auto", "auto", True), ("auto-seq", "auto", False), ("plain", False, True), ("split", True, True),
                         ("split-seq", True, False), ("plain-seq", False, False)):
    match.NN_SPLIT = split
    rc.PIPELINE_PAIRS = pipe
    rc.clear_cache()
    rc.forward_mast3r(pairs_in, model, cache_path="det", subsample=8, desc_conf="desc_conf", device=dev)
    memo = rc._memo("det")
    cur = {k: (v[0][2], v[1][0].clone(), v[1][1].clone(), v[1][2].clone()) for k, v in memo["corres"].items()}
    tot = sum(v[0] for v in cur.values())
    if ref is None:
        ref = cur
    diff = [(k, cur[k][0], ref[k][0]) for k in ref if cur[k][0] != ref[k][0] or not all(torch.equal(a, b) for a, b in zip(cur[k][1:], ref[k][1:]))]
    print(tag, "total correspondences", tot, "pairs differing from the first pass:", diff[:4], len(diff))
