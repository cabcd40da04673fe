#!/usr/bin/env python
"""Benchmark of the two Starst3r hot paths on B200 (contract: see the task statement / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline line (`metric` = Gaussians/s rasterised): one *step* = one 3DGS training iteration of
starster/gs.py:143-161 on BASELINE.json configs[1] (8 views of 512x512, 200 k Gaussians): render all views,
L1+SSIM loss, backward, Adam.  `match` carries the second half of BASELINE.json's metric (512x512 image pairs
matched per second, extract_correspondences of sparse_ga.py:595-630).  Inputs are synthetic.
At N > 1 every rank renders its own 8 views of the replicated splat (weak scaling), the per-Gaussian gradients
are all-reduced over NCCL and Adam is replicated; image pairs are independent and shard without a collective.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_GAUSS, N_VIEWS, W, H = 200_000, 8, 512, 512
WORKLOAD = ("8-view 512x512 synthetic scene, 200k Gaussians, 3DGS train step (render fwd + L1/SSIM loss + bwd + Adam), "
            "BASELINE.json configs[1]")
MATCH_HW = 512


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None      # wall-clock window of the GPU-busy region

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "250"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in ln.split(",")]))

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        rows = [r for ts, r in self.rows if (self.t0 is None or ts >= self.t0) and (self.t1 is None or ts <= self.t1 + 0.05)]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = max([int(r[1]) for _, r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm), "window": "warm-up + timed steps + e2e + match legs (GPU busy throughout)"}


def make_workload(dev, seed, scale_mode="init"):
    from starst3r_b200 import gs, synth
    viewmats, Ks = synth.look_at_cameras(N_VIEWS, W, H, device=dev)
    target = synth.random_splats(N_GAUSS, seed=seed, scale_mode=scale_mode, device=dev)
    with torch.no_grad():
        truth, _, _ = gs.rasterization(target["means"], target["quats"], target["scales"], target["opacities"],
                                       target["shN"], viewmats, Ks, W, H)
    g = torch.Generator().manual_seed(seed + 1)
    params = {k: v.clone().contiguous() for k, v in target.items()}
    params["means"] += 0.01 * torch.randn(N_GAUSS, 3, generator=g).to(dev)        # start off the optimum
    params["shN"] += 0.1 * torch.randn(N_GAUSS, 24, 3, generator=g).to(dev)
    states = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}
    cams = gs.make_cams(viewmats, Ks)
    return params, states, truth.clamp(0, 1).contiguous(), cams


def allreduce_grads(fr, world):
    from starst3r_b200 import dist as sd
    sd.allreduce_gradients(fr.grads)


def run_ours(args):
    from starst3r_b200 import _lib, gs, match, synth
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    hbm_peak, bf16_peak, peak_kind = peaks()

    params, states, truth, cams = make_workload(dev, seed=rank)
    if world > 1:   # replicated splat: every rank starts from rank 0's parameters
        for v in params.values():
            dist.broadcast(v, 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    plan = gs.TrainPlan(N_GAUSS, N_VIEWS, W, H, dev)     # persistent buffers: no allocation / host sync per step
    hook, exchange = None, "none"
    if world > 1:
        # gradient exchange: peer loads inside the fused Adam kernel (symmetric memory over NVLink); NCCL all-reduce
        # only if symmetric memory cannot be set up on this box
        try:
            from starst3r_b200 import dist as sd
            plan.peer = sd.PeerGradExchange(N_GAUSS, dev)
            exchange = ("P2P reduce-scatter + all-gather of the gradient sum over NVLink peer memory (st3r_grad_reduce_scatter) "
                        "+ Adam" if plan.peer.scatter else
                        "fused P2P gradient sum + Adam (st3r_adam_step_peers, NVLink peer loads)")
        except Exception as e:      # noqa: BLE001
            print(f"bench: symmetric memory unavailable ({e!r}); using NCCL all-reduce", file=sys.stderr)
            hook = lambda fr: allreduce_grads(fr, world)
            exchange = "NCCL all-reduce of the gradients + Adam"

    def step_fn(i, prof=False, images=None):
        return gs.train_step(params, states, truth if images is None else images, cams, W, H, i + 1,
                             count_blends=prof, grad_hook=hook, plan=plan)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clk = ClockSampler(local)
    clk.__enter__()
    time.sleep(0.3)                 # let nvidia-smi start sampling
    clk.t0 = time.time()
    for i in range(args.warmup):
        step_fn(i)
    barrier()
    # ---- device-resident timing: K steps, L2 flushed between steps (flush not timed) -----------------
    launches0 = lib.st3r_launch_count()
    evs = []
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss, fr = step_fn(args.warmup + i)
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = lib.st3r_launch_count() - launches0
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    print("per-step ms:", [round(a.elapsed_time(b), 3) for a, b in evs], file=sys.stderr)
    # per-entry-point CUDA events (and the blend counter) in a separate, untimed pass: they cost host time
    gs.PROF = {}
    for i in range(3):
        flush.fill_(i)
        loss, fr = step_fn(args.warmup + args.steps + i, prof=True)
    prof = gs.prof_summary()
    gs.PROF = None
    n_done = args.warmup + args.steps + 3
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = N_GAUSS * N_VIEWS * world / (ms_step * 1e-3)
    n_isect, n_vis = fr.n_isect, int((fr.radii > 0).sum().item())
    n_blend = int(fr.n_blend.item()) if fr.n_blend is not None else 0

    # ---- roofline of the dominant kernel (live CUDA events on the launching stream) -------------------
    px = N_VIEWS * H * W
    alg_bytes = {   # SURVEY §8d algorithmic bytes per launch
        "st3r_gs_project": 92 * N_GAUSS * N_VIEWS + 44 * n_vis,
        "st3r_gs_isect": 12 * n_isect,
        "st3r_radix_sort_pairs": 24 * n_isect,
        "st3r_gs_raster_fwd": 40 * n_isect + 20 * px,
        "st3r_gs_raster_bwd": 40 * n_isect + 24 * px + 36 * n_vis,
        "st3r_gs_project_bwd": 92 * N_GAUSS * N_VIEWS + 92 * N_GAUSS,
        "st3r_gs_loss_fwd": 24 * px, "st3r_gs_loss_bwd": 12 * px, "st3r_adam_step": 644 * N_GAUSS,
    }
    shares = {k: v[1] / max(v[0], 1) for k, v in prof.items()}
    top = max(shares, key=shares.get)
    top_ms = shares[top]
    achieved = alg_bytes.get(top, 0) / (top_ms * 1e-3) / 1e9
    # measured DRAM traffic / issue-slot utilisation of the same kernel from the committed ncu capture (profiles/)
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["kernels"].get(top, {})
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"kernel": top, "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 4), "traffic": ncu.get("dram_bytes"), "peak_kind": peak_kind,
                "note": "the blend kernels are instruction-issue bound, not HBM bound (SURVEY 8d): ncu issue-slot "
                        "utilisation is the figure of merit, the HBM fraction is reported because the schema asks for it",
                "ncu_issue_active_pct": ncu.get("issue_active_pct"), "ncu_sm_throughput_pct": ncu.get("sm_throughput_pct"),
                "kernel_ms": round(top_ms, 4), "alg_bytes": alg_bytes.get(top, 0),
                "step_breakdown_ms": {k: round(v, 4) for k, v in sorted(shares.items(), key=lambda kv: -kv[1])}}

    # ---- second sweep of SURVEY 8d: log-normal scales exp(N(-4, 0.5)) (6x larger splats, ~20x the blends) -----------
    sweep = None
    if world == 1:
        p2, s2, t2, c2 = make_workload(dev, seed=rank, scale_mode="rand")
        plan2 = gs.TrainPlan(N_GAUSS, N_VIEWS, W, H, dev)
        for i in range(4):
            gs.train_step(p2, s2, t2, c2, W, H, i + 1, plan=plan2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(5):
            _, fr2 = gs.train_step(p2, s2, t2, c2, W, H, 5 + i, plan=plan2, count_blends=(i == 4))
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 5
        sweep = {"workload": "same step, scales = exp(N(-4, 0.5)) instead of the 3e-3 initialisation", "ms_per_step": ms2,
                 "gaussians_per_sec": N_GAUSS * N_VIEWS / (ms2 * 1e-3), "intersections": fr2.n_isect,
                 "blends_per_frame": int(fr2.n_blend.item())}
        del p2, s2, t2, c2, plan2, fr2
        torch.cuda.empty_cache()

    # ---- end-to-end: truth images come from pinned host memory every step, the loss is read back -----
    # The upload of step i+1's images runs on a copy stream under step i's kernels (two device buffers); every
    # step's images are copied inside the timed region and every step's loss is read back (one step late, so the
    # read does not drain the queue).
    truth_host = truth.cpu().pin_memory()
    truth_dev = [torch.empty_like(truth), torch.empty_like(truth)]
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]     # upload of buffer b finished
    freed = [torch.cuda.Event(), torch.cuda.Event()]     # step that read buffer b finished

    def upload(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[b])
            truth_dev[b].copy_(truth_host, non_blocking=True)
            ready[b].record(copy_stream)

    barrier()
    for b in (0, 1):
        freed[b].record(main_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    copy_stream.wait_event(e0)
    upload(0)
    losses = []
    for i in range(args.steps):
        b = i & 1
        if i + 1 < args.steps:
            upload(1 - b)
        main_stream.wait_event(ready[b])
        loss, _ = step_fn(n_done + i, images=truth_dev[b])
        freed[b].record(main_stream)
        loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        losses.append(loss)
    e1.record()
    barrier()
    assert all(math.isfinite(x) for x in loss_host.tolist())
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    e2e = {"value": N_GAUSS * N_VIEWS * world / (e2e_ms * 1e-3), "unit": "Gaussians/s", "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": truth_host.numel() * 4, "d2h_bytes_per_step": 4}

    # ---- MATCH: 512x512 image pairs per second (each rank matches its own pairs) ---------------------
    A, B = synth.descriptor_pair(MATCH_HW, MATCH_HW, seed=rank, device=dev)
    A2, B2 = synth.descriptor_pair(MATCH_HW, MATCH_HW, seed=100 + rank, device=dev)
    q = [1 + 9 * torch.rand(MATCH_HW, MATCH_HW, device=dev) for _ in range(4)]
    feats = [A, B, B2, A2]
    for _ in range(3):
        match.extract_correspondences_device(feats, q, 8)
    barrier()
    n_pairs = max(4, args.steps)
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(n_pairs):
        out = match.extract_correspondences_device(feats, q, 8)
    m1.record()
    barrier()
    tm = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    pair_ms = tm.item() / n_pairs
    # e2e: descriptors + confidences from pinned host memory, correspondences read back
    hfe = [f.cpu().pin_memory() for f in feats]
    hq = [x.cpu().pin_memory() for x in q]
    # two device staging sets: pair p+1 is uploaded on the copy stream while pair p is matched; every pair's inputs are
    # copied and every pair's correspondences are read back inside the timed region
    stage = [([torch.empty_like(f) for f in feats], [torch.empty_like(x) for x in q]) for _ in range(2)]
    up_done = [torch.cuda.Event(), torch.cuda.Event()]
    use_done = [torch.cuda.Event(), torch.cuda.Event()]

    def upload_pair(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(use_done[b])
            for dst, src in zip(stage[b][0] + stage[b][1], hfe + hq):
                dst.copy_(src, non_blocking=True)
            up_done[b].record(copy_stream)

    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    for b in (0, 1):
        use_done[b].record(main_stream)
    m0.record()
    copy_stream.wait_event(m0)
    upload_pair(0)
    for p_i in range(n_pairs):
        b = p_i & 1
        if p_i + 1 < n_pairs:
            upload_pair(1 - b)
        main_stream.wait_event(up_done[b])
        xy1, xy2, conf = match.extract_correspondences(stage[b][0], stage[b][1], 8, device=dev)
        use_done[b].record(main_stream)
        xy1.cpu(), xy2.cpu(), conf.cpu()
    m1.record()
    barrier()
    tm = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    pair_e2e_ms = tm.item() / n_pairs
    clk.t1 = time.time()
    clk.__exit__()
    # algorithmic FLOPs of one pair = 2*M*N*24 summed over the NN calls the reference algorithm makes on this input
    rows = match_rows(A, B, A2, B2)
    flops = 2.0 * rows * MATCH_HW * MATCH_HW * 24
    n_corr = int(out[3].item())
    match_res = {"metric": "pairs_per_sec_matched_512x512", "value": 1000.0 / pair_ms * world, "unit": "pairs/s",
                 "ms_per_pair": pair_ms, "correspondences": n_corr,
                 "e2e": {"value": 1000.0 / pair_e2e_ms * world, "unit": "pairs/s",
                         "h2d_bytes_per_step": sum(f.numel() for f in hfe + hq) * 4, "d2h_bytes_per_step": n_corr * 36},
                 "roofline": {"kernel": "nn_tc_kernel (tcgen05 kind::tf32)", "bound": "tensor",
                              "achieved": round(flops / (pair_ms * 1e-3) / 1e12, 1), "peak": round(bf16_peak / 2, 1),
                              "unit": "TFLOP/s", "frac": round(flops / (pair_ms * 1e-3) / 1e12 / (bf16_peak / 2), 4),
                              "peak_kind": peak_kind + " bf16 / 2 (TF32 dense runs at half the bf16 rate)",
                              "alg_flops_per_pair": flops, "query_rows_per_pair": rows}}

    # ---- MATCH + ALIGN end to end (rank 0, N = 1 only): reconstruct_scene on BASELINE.json configs[1] ------------
    recon = None
    if world == 1 and rank == 0 and not args.no_align:
        try:
            recon = reconstruct_leg(dev)
        except Exception as e:      # noqa: BLE001 - an auxiliary leg must not take the headline line down with it
            recon = {"error": repr(e)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CPU baseline (oracle port, bounded sample) on rank 0 at N = 1 --------------------------------
    cpu = cpu_baseline(sample_views=1) if world == 1 and not args.no_cpu else None
    line = {"metric": "gaussians_per_sec_rasterized", "value": value, "unit": "Gaussians/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "views_per_gpu": N_VIEWS, "gaussians": N_GAUSS, "image": [H, W],
                       "parallelism": f"views sharded dp{world}, {exchange}" if world > 1 else "single GPU",
                       "l2": "256 MiB flush between timed steps",
                       # opt-in kernel variants (all 0 / False = the default kernels)
                       "kernel_variants": _kernel_variants()},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "blends_per_sec": n_blend * world / (shares.get("st3r_gs_raster_fwd", float("nan")) * 1e-3),
            "blends_per_frame": n_blend, "intersections": n_isect, "visible": n_vis, "loss": float(loss.item()),
            "sweep_lognormal_scales": sweep, "match": match_res, "reconstruct": recon}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


class _CachedNet:
    """Stands in for the (out-of-scope) MASt3R network: returns pre-computed synthetic predictions per image pair."""

    def __init__(self, net, imgs, dev):
        n = len(imgs)
        self.cache = {}
        for i in range(n):
            for j in range(i):          # make_pairs('complete') order: (i, j) with j < i, then the symmetric ones
                for a, b in ((i, j), (j, i)):
                    res = net.symmetric_inference({"idx": a}, {"idx": b})
                    self.cache[a, b] = tuple({k: v.to(dev) for k, v in r.items()} for r in res)

    def symmetric_inference(self, img1, img2, device=None):
        return self.cache[int(img1["idx"]), int(img2["idx"])]


def reconstruct_leg(dev):
    """MATCH + ALIGN on BASELINE.json configs[1] (8 views 512x512): starster.reconstruct_scene = 28 image pairs matched
    (extract_correspondences), canonical views, MST, the 500 + 200 iteration sparse global alignment, dense points
    and clean_pointcloud.  The network predictions are synthetic and pre-computed (the network is out of scope)."""
    from starst3r_b200 import reconstruct as rc
    from starst3r_b200 import synth
    n = N_VIEWS
    net = synth.SyntheticMast3r(n, W, H, seed=0, device="cpu", arc_deg=120.0)
    imgs = net.images()
    model = _CachedNet(net, imgs, dev)
    out = {}
    for rep_i in range(2):              # first pass warms the kernels / allocator, second is reported
        rc._MEMO.clear()
        torch.cuda.synchronize()
        t0 = time.time()
        scene, _ = rc.reconstruct_scene(model, imgs, [f"{i}.png" for i in range(n)], dev)
        pts, _, confs = scene.get_dense_pts3d(clean_depth=True)
        torch.cuda.synchronize()
        out = {"seconds": time.time() - t0}
    # the matcher alone on this scene's descriptor maps: smooth fields (like real MASt3R descriptors) keep ~100 columns
    # per query row inside the TF32 error band, which the kernel resolves exactly on the spot; the headline `match`
    # figure uses the random descriptors SURVEY 8d defines, where the band holds 1-3 columns
    from starst3r_b200 import match
    res = model.cache[1, 0]
    feats = [r["desc"][0].float().contiguous() for r in res]
    qonfs = [r["desc_conf"][0].float().contiguous() for r in res]
    for _ in range(2):
        match.extract_correspondences_device(feats, qonfs, 8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        o = match.extract_correspondences_device(feats, qonfs, 8)
    e1.record()
    torch.cuda.synchronize()
    ms_pair = e0.elapsed_time(e1) / 8
    out["match_smooth_descriptors"] = {"ms_per_pair": ms_pair, "pairs_per_s": 1000.0 / ms_pair,
                                       "correspondences": int(o[3].item())}
    n_pairs = n * (n - 1) // 2
    out.update({"workload": f"{n} views {W}x{H}: {n_pairs} pairs matched + sparse global alignment (500 + 200 iterations) + "
                            "dense points + clean_pointcloud; synthetic network predictions pre-computed on the device",
                "pairs": n_pairs, "align_iterations": 700, "dense_points": int(sum(p.shape[0] for p in pts)),
                "views_per_s": n / out["seconds"]})
    return out


def match_rows(A, B, A2, B2):
    """Query rows the reference algorithm issues for one image pair (4 seeded searches, fast_nn.py:152-168)."""
    from starst3r_b200 import match
    return sum(sum(match.recip_query_rows(P1, P2, 8)) for P1, P2 in ((A, B), (B, A), (A2, B2), (B2, A2)))


def cpu_baseline(sample_views=1):
    """The oracle (CPU port of the reference algorithm: PyTorch fp32 restatement of gsplat 1.4 + torch Adam) timed on
    the host cores on a bounded sample: `sample_views` of the 8 views, all 200k Gaussians, one full train step."""
    from oracle import gs_oracle as go
    from starst3r_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    viewmats, Ks = synth.look_at_cameras(N_VIEWS, W, H)
    sp = synth.random_splats(N_GAUSS, seed=0, scale_mode="init")
    params = {k: v.clone() for k, v in sp.items()}
    states = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}
    truth = torch.rand(sample_views, H, W, 3)
    t0 = time.time()
    go.train_step(params, states, truth, viewmats[:sample_views], Ks[:sample_views], W, H, 1)
    dt = time.time() - t0
    return {"value": N_GAUSS * sample_views / dt, "unit": "Gaussians/s", "cores": cores, "kind": "port",
            "sample": f"{sample_views} of {N_VIEWS} views x {N_GAUSS} Gaussians, one train step "
                      f"(oracle/gs_oracle.py, torch {torch.__version__} CPU), {dt:.1f} s",
            "seconds": dt}


def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric.  gsplat has no CPU backend and is not
    installable here, so the oracle port (PyTorch restatement of the algorithm + torch Adam) stands in."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    vals = []
    for _ in range(max(1, min(args.steps, 2))):
        vals.append(cpu_baseline(sample_views=1))
    best = max(vals, key=lambda d: d["value"])
    line = {"impl": "reference", "metric": "gaussians_per_sec_rasterized", "value": best["value"],
            "unit": "Gaussians/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": len(vals), "warmup": 0,
            "ms_per_step": best["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "views_per_gpu": N_VIEWS, "gaussians": N_GAUSS, "image": [H, W],
                       "parallelism": "host cores (PyTorch CPU)", "sample": "each step = bounded sample of 1 of the 8 views"},
            "cpu_baseline": best,
            "e2e": {"value": best["value"], "unit": "Gaussians/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def _kernel_variants():
    from starst3r_b200 import gs, match
    from starst3r_b200 import reconstruct as rc
    return {"raster": int(gs.RASTER_VARIANT), "nn_split": match.NN_SPLIT, "nn_cooperative": match.NN_COOPERATIVE,
            "align": int(rc.ALIGN_VARIANT)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-align", action="store_true", help="skip the MATCH + ALIGN reconstruct leg")
    args = ap.parse_args()
    # >= 12 untimed steps: a one-off ~7 ms host stall (lazy driver / allocator initialisation, seen at the 10th
    # iteration of a process whatever the kernels are) would otherwise land inside a 10-step timed region
    args.warmup = max(args.warmup, 12) if args.impl == "ours" else args.warmup
    # Exactly ONE line goes to stdout (the JSON): libraries that chat on fd 1 (NCCL prints its version there) are
    # diverted to stderr until the result is ready.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    _emit.fd = saved
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    run_ours(args)


def _emit(line):
    sys.stdout.flush()
    os.write(_emit.fd, (json.dumps(line) + "\n").encode())


_emit.fd = 1


if __name__ == "__main__":
    main()
