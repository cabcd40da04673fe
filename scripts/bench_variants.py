"""A/B timing of the opt-in kernel variants of DESIGN.md §10 against the default kernels (CUDA events, one GPU).
Development aid for the first GPU call of a round; bench.py stays the judged entry point.

  python scripts/bench_variants.py [raster] [align] [match]      (default: all three)

Prints one JSON object; every leg is isolated in try/except so that one failing variant does not hide the others.
Run the parity tests first: ST3R_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -m gpu -q"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from starst3r_b200 import gs, match, synth  # noqa: E402
from starst3r_b200 import reconstruct as rc  # noqa: E402


def ev_time(fn, warm, iters, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)           # evicts the 126 MB L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def raster_leg(dev, out):
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    for scale_mode in ("init", "rand"):      # 3e-3 initial scale / exp(N(-4, 0.5)) sweep of SURVEY 8d
        params, states, truth, cams = bench.make_workload(dev, 0, scale_mode=scale_mode)
        for variant in [int(v) for v in os.environ.get("ST3R_VARIANTS", "0,1").split(",")]:
            key = f"train_step_ms[{scale_mode}][raster_variant={variant}]"
            try:
                gs.RASTER_VARIANT = variant
                p = {k: v.clone() for k, v in params.items()}
                s = {k: (a.clone(), b.clone()) for k, (a, b) in states.items()}
                plan = gs.TrainPlan(bench.N_GAUSS, bench.N_VIEWS, bench.W, bench.H, dev)
                it = [0]

                def step():
                    it[0] += 1
                    gs.train_step(p, s, truth, cams, bench.W, bench.H, it[0], plan=plan)
                out[key] = ev_time(step, 12, 10, flush)
                gs.PROF = {}
                for _ in range(5):
                    flush.add_(1)
                    step()
                prof = gs.prof_summary()
                gs.PROF = None
                out[key + "[kernels_ms]"] = {k.replace("st3r_gs_", ""): round(v[1] / v[0], 4) for k, v in prof.items()}
            except Exception as e:
                out[key] = "ERROR " + repr(e)[:300]
            finally:
                gs.RASTER_VARIANT = 0


def align_leg(dev, out):
    n = bench.N_VIEWS
    net = synth.SyntheticMast3r(n, bench.W, bench.H, seed=0, device="cpu", arc_deg=120.0)
    imgs = net.images()
    model = bench._CachedNet(net, imgs, dev)
    for variant in (0, 1, 2, 3):
        key = f"reconstruct_s[align_variant={variant}]"
        try:
            rc.ALIGN_VARIANT = variant
            for _ in range(2):
                rc._MEMO.clear()
                torch.cuda.synchronize()
                t0 = time.time()
                scene, _ = rc.reconstruct_scene(model, imgs, [f"{i}.png" for i in range(n)], dev)
                scene.get_dense_pts3d(clean_depth=True)
                torch.cuda.synchronize()
                out[key] = time.time() - t0
        except Exception as e:
            out[key] = "ERROR " + repr(e)[:300]
        finally:
            rc.ALIGN_VARIANT = 0
    return model


def match_leg(dev, out, model=None):
    A, B = synth.descriptor_pair(512, 512, seed=0, device=dev)
    A2, B2 = synth.descriptor_pair(512, 512, seed=100, device=dev)
    q = [1 + 9 * torch.rand(512, 512, device=dev) for _ in range(4)]
    cases = {"random": ([A, B, B2, A2], q)}
    if model is None:
        net = synth.SyntheticMast3r(2, 512, 512, seed=0, device="cpu", arc_deg=30.0)
        model = bench._CachedNet(net, net.images(), dev)
    res = model.cache[1, 0]
    cases["smooth"] = ([r["desc"][0].float().contiguous() for r in res], [r["desc_conf"][0].float().contiguous() for r in res])
    for name, (feats, qonfs) in cases.items():
        for coop in (False, True):
            for split in (False, True):
                key = f"match_ms_per_pair[{name}][coop={int(coop)}][split={int(split)}]"
                try:
                    match.NN_COOPERATIVE, match.NN_SPLIT = coop, split
                    out[key] = ev_time(lambda: match.extract_correspondences_device(feats, qonfs, 8), 3, 8)
                except Exception as e:
                    out[key] = "ERROR " + repr(e)[:300]
                finally:
                    match.NN_COOPERATIVE, match.NN_SPLIT = "auto", "auto"


def main():
    legs = sys.argv[1:] or ["raster", "align", "match"]
    dev = torch.device("cuda:0")
    out, model = {}, None
    if "raster" in legs:
        raster_leg(dev, out)
    if "align" in legs:
        model = align_leg(dev, out)
    if "match" in legs:
        match_leg(dev, out, model)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
