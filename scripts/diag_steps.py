"""Development aid: how the headline workload evolves over training steps (intersections, blends, scales) and what the
blend kernels cost at each stage, per raster variant."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from starst3r_b200 import gs

dev = torch.device("cuda:0")
out = {}
variants = [int(v) for v in os.environ.get("ST3R_VARIANTS", "1,0").split(",")]
for variant in variants:
    gs.RASTER_VARIANT = variant
    params, states, truth, cams = bench.make_workload(dev, 0)
    plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    for i in range(1, 41):
        prof = i in (1, 3, 6, 10, 15, 20, 25, 30, 40)
        if prof:
            gs.PROF = {}
            flush.add_(1)
        loss, fr = gs.train_step(params, states, truth, cams, bench.W, bench.H, i, plan=plan, count_blends=prof)
        if prof:
            p = gs.prof_summary()
            gs.PROF = None
            sc = params["scales"].abs()
            out[f"v{variant} step {i}"] = {
                "n_isect": int(fr.n_isect), "n_blend": int(fr.n_blend.item()), "loss": round(float(loss.item()), 4),
                "scale_abs_mean": round(float(sc.mean()), 5), "scale_abs_max": round(float(sc.max()), 5),
                "fwd_ms": round(p["st3r_gs_raster_fwd"][1], 4), "bwd_ms": round(p["st3r_gs_raster_bwd"][1], 4),
                "bin_ms": round(p["st3r_gs_bin_tiles"][1], 4)}
print(json.dumps(out, indent=1))
