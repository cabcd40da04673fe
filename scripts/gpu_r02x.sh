#!/bin/bash
# GPU pass r02x: the ALIGN loop as one cooperative launch (variant 7) vs three launches per iteration (variant 3), the
# pipelined pair loop on / off, host-side profile of reconstruct_scene, parity suite, bench.
set -u
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== align breakdown, variant 7 / 3"
for v in 7 3; do
  ST3R_ALIGN_VARIANT=$v timeout 300 python scripts/prof_align.py 2>&1 | tail -1 | tee $OUT/${TAG}_align_breakdown_v$v.json
done
echo "== pair loop pipelined / sequential (forward_mast3r, 28 pairs, warm)"
timeout 300 python - <<'PY'
import sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
from starst3r_b200 import reconstruct as rc, synth
dev = torch.device("cuda:0")
n = bench.N_VIEWS
net = synth.SyntheticMast3r(n, bench.W, bench.H, seed=0, device="cpu", arc_deg=120.0)
imgs = net.images()
model = bench._CachedNet(net, imgs, dev)
names = [f"{i}.png" for i in range(n)]
pairs_in = rc.convert_dust3r_pairs_naming(names, rc.make_pairs(rc.prepare_images_for_mast3r(imgs), "complete", None, True))
for rep in range(3):
    for pipe in (True, False):
        rc.PIPELINE_PAIRS = pipe
        rc.clear_cache(); torch.cuda.synchronize(); t0 = time.time()
        rc.forward_mast3r(pairs_in, model, cache_path="x", subsample=8, desc_conf="desc_conf", device=dev)
        torch.cuda.synchronize()
        print("rep", rep, "pipelined" if pipe else "sequential", round(time.time() - t0, 4), "s")
PY
echo "== cProfile of reconstruct_scene (warm pass)"
timeout 300 python scripts/prof_reconstruct.py 2>&1 | tail -45 | tee $OUT/${TAG}_cprofile_reconstruct.txt | head -60
echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["step_breakdown_ms"])
print(d["reconstruct"]["seconds"], d["reconstruct"]["stages_s"], d["match"]["value"], d["reconstruct"].get("align"))
PY
tail -3 $OUT/${TAG}_bench.err
