"""RASTER + ADAM hot path — host-side mirror of starster/gs.py (init_3dgs, render_3dgs, render_3dgs_original,
run_3dgs_optim) and of the one gsplat entry point it calls (gsplat.rasterization, gs.py:76-87).

Everything numeric runs in hand-written sm_100a kernels behind the C ABI (include/starst3r_b200.h):
projection + SH -> tile count -> scan -> key emission -> radix sort -> tile offsets -> alpha blend, the blend /
projection / SH backward, the fused L1+SSIM loss and a fused multi-tensor Adam.  No gsplat, no torchmetrics,
no CPU fallback.  PyTorch only owns the memory and the stream.
"""
import ctypes
import math
import os

import torch
from tqdm import trange

from . import _lib

__all__ = ("init_3dgs", "render_3dgs", "render_3dgs_original", "render_3dgs_path", "run_3dgs_optim", "rasterization",
           "FusedAdam", "MCMCStrategy", "TrainPlan")

TILE = 16
# "fused": st3r_gs_bin_tiles (counting sort by tile + in-tile shared-memory sort); "radix": the generic
# st3r_gs_isect -> st3r_radix_sort_pairs -> st3r_gs_offsets chain.  Both produce identical arrays.
BINNING = "fused"
EPS2D, NEAR, FAR, RADIUS_CLIP = 0.3, 0.01, 1e10, 0.0
# Implementation of the blend kernels (st3r_gs_set_raster_variant): 0 = fragment-pool kernels (default), 1 = visit-list
# kernels (the first implementation, kept as an independent cross-check; tests/test_gs_gpu.py compares the two).
RASTER_VARIANT = int(os.environ.get("ST3R_RASTER_VARIANT", "0"))
# run_3dgs_optim under a torch.distributed process group: shard the views over the ranks (splat replicated, gradients
# summed over NVLink peer memory / NVSwitch multimem / NCCL).  Validated on 2 / 4 / 8 B200s (replicas bit-identical and
# equal to the single-GPU run: tests/test_dist_gpu.py, bench.py --gpus N), so it is the default whenever a process group
# of more than one rank exists; with the MCMC strategy on (enable_pruning=True: per-rank random draws would split the
# replicas) the ranks train as plain replicas instead.  ST3R_SHARD_VIEWS=0 / 1 forces it off / on (1 + pruning raises).
SHARD_VIEWS = {"0": False, "1": True}.get(os.environ.get("ST3R_SHARD_VIEWS", ""), "auto")


# Steady-state iterations of train_step (a sized TrainPlan, no strategy hook, one GPU) are captured once as a CUDA graph
# and replayed: the ~25 host calls of an iteration (14 kernel launches through ctypes + the torch plumbing) cost more
# host time than the 1.8 ms the kernels need, so the un-captured loop runs at the host's pace whenever the host hiccups.
# Sharded views on 4 and more GPUs are captured too (two graphs, one per parity of the alternating gradient buffers: the
# cross-device barriers, the in-switch / peer-memory reduce-scatter and Adam are all stream-ordered kernels); the fused
# P2P gradient sum + Adam of 2-GPU runs takes the step number as a launch argument and stays eager.
# ST3R_TRAIN_GRAPH=0 keeps every iteration eager (the cross-check: tests compare the two).
TRAIN_GRAPH = os.environ.get("ST3R_TRAIN_GRAPH", "1") == "1"


class _Prof:
    """Optional per-entry-point CUDA-event timing (bench.py's roofline leg): `gs.PROF = {}` turns it on; each C-ABI
    call then appends (start, end) events on the launching stream under its name."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PROF is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROF is not None:
            self.e1.record()
            PROF.setdefault(self.name, []).append((self.e0, self.e1))
        return False


PROF = None


def prof_summary():
    """{name: (calls, total_ms)} of the events collected so far (synchronises)."""
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (PROF or {}).items()}


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _f(x):
    return ctypes.c_float(float(x))


def make_cams(viewmats, Ks):
    """[C, 19] fp32 camera records: R (9) | t (3) | fx fy cx cy | camera centre (3)."""
    C = viewmats.shape[0]
    V = viewmats.detach().float()
    K = Ks.detach().float()
    pos = torch.linalg.inv(V.cpu())[:, :3, 3].to(V.device)     # C tiny inverses: host (no cuSOLVER initialisation)
    cams = torch.cat([V[:, :3, :3].reshape(C, 9), V[:, :3, 3], K[:, 0, 0:1], K[:, 1, 1:2], K[:, 0, 2:3], K[:, 1, 2:3],
                      pos], dim=1)
    return cams.contiguous()


class _Frame:
    """Device state of one forward pass (kept for the backward)."""
    _n_isect = None

    @property
    def n_isect(self):
        """Number of (Gaussian, tile) intersections of this frame (host int; read from the device on first use)."""
        if self._n_isect is None:
            self._n_isect = int(self.n_isect_dev.item())
        return self._n_isect


class TrainPlan:
    """Persistent device buffers for repeated frames of one shape: no allocation and NO host synchronisation per
    frame.  gsplat (and the un-planned path below) reads the intersection count back every frame to size the sort
    buffers exactly; here they have a capacity (HEADROOM x the last known count), the kernels clamp to it on the
    device, and the count is copied to pinned memory asynchronously and checked one frame later.  A frame can only be
    rendered truncated if the count grows by more than HEADROOM within the two frames the host runs ahead, so the plan
    keeps the count under watch: its first SYNC_FRAMES frames, and SYNC_FRAMES frames after any frame-to-frame growth
    above GROWTH_WATCH, read the count back BEFORE binning (the exact-size discipline of gsplat: the buffers are grown
    first, nothing is truncated, nothing wrong reaches the optimiser); a count above 80 % of the capacity grows the
    buffers for the next frame; an overflow that still gets through raises (loudly, never truncates silently)."""
    HEADROOM = 1.5
    SYNC_FRAMES = 8
    GROWTH_WATCH = 0.05

    def __init__(self, N, C, width, height, device):
        lib = _lib.load()
        self.N, self.C, self.W, self.H, self.dev = N, C, int(width), int(height), device
        self.tile_w, self.tile_h = (self.W + TILE - 1) // TILE, (self.H + TILE - 1) // TILE
        E = max(C * N, 1)
        i32, f32 = torch.int32, torch.float32
        self.radii = torch.empty(E, dtype=i32, device=device)
        self.geomA = torch.empty((E, 4), dtype=f32, device=device)
        self.geomB = torch.empty((E, 4), dtype=f32, device=device)
        self.rgb = torch.empty((E, 4), dtype=f32, device=device)
        self.tiles = torch.empty(E, dtype=i32, device=device)
        self.cum = torch.empty(E, dtype=i32, device=device)
        self.n_raw = torch.zeros(1, dtype=i32, device=device)        # true count of the current frame
        self.n_clamped = torch.zeros(1, dtype=i32, device=device)    # min(count, capacity): what the kernels use
        self.scan_ws = _ws(lib.st3r_scan_ws_bytes(C * N), device)
        self.offsets = torch.empty(max(C * self.tile_w * self.tile_h, 1), dtype=i32, device=device)
        self.render = torch.empty((C, self.H, self.W, 3), dtype=f32, device=device)
        self.alphas = torch.empty((C, self.H, self.W), dtype=f32, device=device)
        self.last_ids = torch.empty((C, self.H, self.W), dtype=i32, device=device)
        self.n_blend = torch.zeros(1, dtype=torch.int64, device=device)
        # backward / loss
        self.v_geom = torch.empty((3, E, 4), dtype=f32, device=device)
        self.grads = dict(means=torch.empty((N, 3), dtype=f32, device=device), quats=torch.empty((N, 4), dtype=f32, device=device),
                          scales=torch.empty((N, 3), dtype=f32, device=device), opacities=torch.empty((N,), dtype=f32, device=device),
                          sh=torch.empty((N, 4, 3), dtype=f32, device=device))
        self.dmaps = torch.empty((C, self.H, self.W, 3, 3), dtype=f32, device=device)
        self.acc = torch.zeros((2 * C + 2,), dtype=f32, device=device)    # [C,2] SSIM / L1 sums | 2 regulariser sums
        self.v_render = torch.empty((C, self.H, self.W, 3), dtype=f32, device=device)
        self.peer = None       # dist.PeerGradExchange when the views are sharded over GPUs (gradients live there)
        self.cap = 0
        self.keys = self.vals = self.keys_alt = self.vals_alt = self.sort_ws = None
        self._host = torch.zeros(64, dtype=i32).pin_memory()           # ring of asynchronously copied counts
        self._pending = []                                              # (event, slot, capacity at launch)
        self._slot = 0
        self.last_n_isect = None
        self._sync_left = self.SYNC_FRAMES                              # frames that still read the count synchronously
        # graph mode (train_step): captured iterations keyed by the tensors they touch, device-side Adam step counter
        self.steps_done = torch.zeros(1, dtype=i32, device=device)
        self._steps_done_host = None
        self.loss_buf = torch.zeros((), dtype=f32, device=device)
        self._graphs = {}
        self._epoch = 0                                                 # bumped when the intersection buffers move
        self._capturing = False
        self.graph_replays = 0

    def sync_mode(self):
        return self._sync_left > 0

    def _observe(self, n):
        """Book-keeping of one known count: growth watch + capacity for the next frame."""
        prev = self.last_n_isect
        if prev and n > (1.0 + self.GROWTH_WATCH) * prev:
            self._sync_left = max(self._sync_left, self.SYNC_FRAMES)
        self.last_n_isect = n

    def matches(self, N, C, width, height, device):
        return (self.N, self.C, self.W, self.H) == (N, C, int(width), int(height)) and self.dev == device

    def _alloc_isect(self, cap):
        lib = _lib.load()
        self.cap = int(cap)
        self.keys = torch.empty(self.cap, dtype=torch.int64, device=self.dev)
        self.vals = torch.empty(self.cap, dtype=torch.int32, device=self.dev)
        self.keys_alt = torch.empty(self.cap, dtype=torch.int64, device=self.dev)
        self.vals_alt = torch.empty(self.cap, dtype=torch.int32, device=self.dev)
        self.sort_ws = _ws(lib.st3r_radix_sort_ws_bytes(self.cap), self.dev)
        self.bin_ws = _ws(lib.st3r_gs_bin_ws_bytes(self.C, self.W, self.H, TILE, self.cap), self.dev)
        self.n_total = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._epoch += 1
        self._graphs.clear()            # captured iterations point into the old buffers

    def poll(self, wait_all=False):
        """Checks the counts of earlier frames (waits for all but the newest, so the host stays at most one frame
        ahead of the device).  Raises on overflow; grows the capacity for the NEXT frame when it runs tight."""
        if self._capturing:
            return
        keep = 0 if wait_all else 1
        while len(self._pending) > keep:
            ev, slot, cap = self._pending.pop(0)
            ev.synchronize()
            n = int(self._host[slot])
            self._observe(n)
            if n > 0.8 * self.cap:
                torch.cuda.current_stream().synchronize()
                self._alloc_isect(max(int(self.HEADROOM * n), 1024))
            if n > cap:
                raise RuntimeError(f"starst3r_b200.gs.TrainPlan: {n} tile intersections exceed the buffer capacity {cap}; "
                                   "that frame was rendered truncated - re-run it (the capacity has been raised)")

    def after_scan(self):
        """Called right after the scan of a frame: capacity bookkeeping without stalling the device."""
        if self.cap == 0 or self._sync_left > 0:   # watched frame: the true count decides BEFORE anything is binned
            n = int(self.n_raw.item())
            self._observe(n)
            self._sync_left = max(self._sync_left - 1, 0)
            if n > 0.8 * self.cap:
                self._alloc_isect(max(int(self.HEADROOM * n), 1024))
        else:
            slot = self._slot = (self._slot + 1) % self._host.numel()
            self._host[slot:slot + 1].copy_(self.n_raw, non_blocking=True)
            if not self._capturing:     # a captured iteration: graph_step records the event after every replay
                ev = torch.cuda.Event()
                ev.record()
                self._pending.append((ev, slot, self.cap))
        torch.clamp(self.n_raw, max=self.cap, out=self.n_clamped)

    def graph_step(self, params, states, truth, cams, width, height, step, hyper, loss_out):
        """One steady-state iteration as a CUDA-graph replay (captured on first use for this set of tensors).  Returns
        None when the iteration has to run eagerly (count under watch, buffers not sized yet)."""
        self.poll()
        if self.cap == 0 or self.sync_mode() or BINNING != "fused":
            return None
        names = ("means", "scales", "quats", "opacities", "shN")
        key = (self._epoch, int(RASTER_VARIANT), truth.data_ptr(), cams.data_ptr(), tuple(params[k].data_ptr() for k in names),
               tuple(t.data_ptr() for k in names for t in states[k]), tuple(params[k].shape for k in names), hyper,
               (id(self.peer), self.peer.step & 1) if self.peer is not None else None)   # gradient buffers alternate
        ent = self._graphs.get(key)
        lib = _lib.load()
        if ent is None:
            if len(self._graphs) >= 8:
                self._graphs.clear()
            lr, betas, eps, f_ssim, f_opac, f_scale = hyper
            graph = torch.cuda.CUDAGraph()
            n0 = lib.st3r_launch_count()
            self._capturing = True
            try:
                # (thread-local capture mode: NCCL's watchdog thread may query its events meanwhile)
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    _, fr = _train_step_eager(params, states, truth, cams, width, height, None, lr, betas, eps, f_ssim,
                                              f_opac, f_scale, False, None, self, self.loss_buf)
            finally:
                self._capturing = False
            ent = self._graphs[key] = (graph, fr, self._slot, int(lib.st3r_launch_count() - n0))
        graph, fr, slot, launches = ent
        if self._steps_done_host != step - 1:
            self.steps_done.fill_(step - 1)
        graph.replay()
        self._steps_done_host = step
        self.graph_replays += 1
        if self.peer is not None:
            self.peer.advance()
        lib.st3r_launch_count_add(launches)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((ev, slot, self.cap))
        fr._n_isect = None
        if loss_out is not None:
            loss_out.copy_(self.loss_buf)
            return loss_out, fr
        return self.loss_buf.clone(), fr


def _forward(means, quats, scales, opacities, shN, cams, width, height, count_blends=False, plan=None):
    lib = _lib.load()
    _lib.require_cuda(means, quats, scales, opacities, shN, cams)
    dev = means.device
    N, C = means.shape[0], cams.shape[0]
    assert lib.st3r_gs_cam_floats() == cams.shape[1]
    sh_coeffs = shN.shape[1]
    fr = _Frame()
    fr.N, fr.C, fr.W, fr.H, fr.sh_coeffs, fr.cams = N, C, int(width), int(height), sh_coeffs, cams
    fr.tile_w, fr.tile_h = (fr.W + TILE - 1) // TILE, (fr.H + TILE - 1) // TILE
    E = max(C * N, 1)
    if plan is not None:
        assert plan.matches(N, C, width, height, dev), "TrainPlan was built for another problem shape"
        plan.poll()
        fr.radii, fr.geomA, fr.geomB, fr.rgb, fr.tiles, fr.cum = plan.radii, plan.geomA, plan.geomB, plan.rgb, plan.tiles, plan.cum
        fr.n_isect_dev = plan.n_raw
    else:
        fr.radii = torch.empty(E, dtype=torch.int32, device=dev)
        fr.geomA = torch.empty((E, 4), dtype=torch.float32, device=dev)
        fr.geomB = torch.empty((E, 4), dtype=torch.float32, device=dev)
        fr.rgb = torch.empty((E, 4), dtype=torch.float32, device=dev)
        fr.tiles = torch.empty(E, dtype=torch.int32, device=dev)
        fr.cum = torch.empty(E, dtype=torch.int32, device=dev)
        fr.n_isect_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    st = _lib.stream_ptr()
    with torch.cuda.device(dev):
        with _Prof("st3r_gs_project"):
            _lib.check(lib.st3r_gs_project(_lib.ptr(means), _lib.ptr(quats), _lib.ptr(scales), _lib.ptr(opacities),
                                           _lib.ptr(shN), sh_coeffs, _lib.ptr(cams), N, C, fr.W, fr.H, TILE, _f(EPS2D),
                                           _f(NEAR), _f(FAR), _f(RADIUS_CLIP), _lib.ptr(fr.radii), _lib.ptr(fr.geomA),
                                           _lib.ptr(fr.geomB), _lib.ptr(fr.rgb), _lib.ptr(fr.tiles), st), "st3r_gs_project")
        # The per-entry scan (torch.cumsum of tiles_per_gauss in gsplat) gives the emission offsets of the generic chain
        # and the intersection total.  The fused binning does not need the offsets and reports the total itself, so a
        # sized plan skips the scan altogether.
        late_count = plan is not None and plan.cap > 0 and BINNING == "fused" and not plan.sync_mode()
        if not late_count:
            ws = plan.scan_ws if plan is not None else _ws(lib.st3r_scan_ws_bytes(C * N), dev)
            with _Prof("st3r_exclusive_scan_i32"):
                _lib.check(lib.st3r_exclusive_scan_i32(_lib.ptr(fr.tiles), _lib.ptr(fr.cum), C * N, _lib.ptr(fr.n_isect_dev),
                                                       _lib.ptr(ws), ws.numel(), st), "st3r_exclusive_scan_i32")
        if plan is not None:
            # capacity-sized intersection buffers, count clamped on the device, checked one frame later
            if not late_count:
                plan.after_scan()
            n_cap, n_dev = plan.cap, plan.n_clamped
            fr.keys, fr.vals, keys_alt, vals_alt, ws = plan.keys, plan.vals, plan.keys_alt, plan.vals_alt, plan.sort_ws
            bin_ws, n_total = plan.bin_ws, (plan.n_raw if late_count else plan.n_total)
            fr.n_isect_dev = n_dev
            fr.offsets, fr.render, fr.alphas, fr.last_ids = plan.offsets, plan.render, plan.alphas, plan.last_ids
            fr.n_blend = plan.n_blend.zero_() if count_blends else None
        else:
            # One host read per frame sizes the intersection buffers exactly (gsplat does the same).
            n_isect = int(fr.n_isect_dev.item())
            fr._n_isect = n_isect
            n_cap, n_dev = n_isect, fr.n_isect_dev
            fr.keys = torch.empty(max(n_cap, 1), dtype=torch.int64, device=dev)
            fr.vals = torch.empty(max(n_cap, 1), dtype=torch.int32, device=dev)
            fr.offsets = torch.empty(max(C * fr.tile_w * fr.tile_h, 1), dtype=torch.int32, device=dev)
            fr.render = torch.empty((C, fr.H, fr.W, 3), dtype=torch.float32, device=dev)
            fr.alphas = torch.empty((C, fr.H, fr.W), dtype=torch.float32, device=dev)
            fr.last_ids = torch.empty((C, fr.H, fr.W), dtype=torch.int32, device=dev)
            fr.n_blend = torch.zeros(1, dtype=torch.int64, device=dev) if count_blends else None
            if BINNING == "fused":
                bin_ws = _ws(lib.st3r_gs_bin_ws_bytes(C, fr.W, fr.H, TILE, n_cap), dev)
                n_total = torch.zeros(1, dtype=torch.int32, device=dev)
            else:
                keys_alt = torch.empty(max(n_cap, 1), dtype=torch.int64, device=dev)
                vals_alt = torch.empty(max(n_cap, 1), dtype=torch.int32, device=dev)
                ws = _ws(lib.st3r_radix_sort_ws_bytes(max(n_cap, 1)), dev)
        if BINNING == "fused":
            with _Prof("st3r_gs_bin_tiles"):
                _lib.check(lib.st3r_gs_bin_tiles(_lib.ptr(fr.radii), _lib.ptr(fr.geomA), N, C, fr.W, fr.H, TILE,
                                                 _lib.ptr(fr.offsets), _lib.ptr(n_total), _lib.ptr(fr.keys), _lib.ptr(fr.vals),
                                                 n_cap, _lib.ptr(bin_ws), bin_ws.numel(), st), "st3r_gs_bin_tiles")
            if late_count:
                plan.after_scan()           # the total just written by the binning: async copy + device-side clamp
        else:
            with _Prof("st3r_gs_isect"):
                _lib.check(lib.st3r_gs_isect(_lib.ptr(fr.radii), _lib.ptr(fr.geomA), _lib.ptr(fr.cum), N, C, fr.W, fr.H, TILE,
                                             _lib.ptr(fr.keys), _lib.ptr(fr.vals), n_cap, st), "st3r_gs_isect")
            bits = lib.st3r_gs_sort_bits(C, fr.W, fr.H, TILE)
            with _Prof("st3r_radix_sort_pairs"):
                _lib.check(lib.st3r_radix_sort_pairs(_lib.ptr(fr.keys), _lib.ptr(fr.vals), _lib.ptr(keys_alt),
                                                     _lib.ptr(vals_alt), _lib.ptr(n_dev), n_cap, 0, bits,
                                                     _lib.ptr(ws), ws.numel(), st), "st3r_radix_sort_pairs")
            with _Prof("st3r_gs_offsets"):
                _lib.check(lib.st3r_gs_offsets(_lib.ptr(fr.keys), _lib.ptr(n_dev), n_cap, C, fr.W, fr.H, TILE,
                                               _lib.ptr(fr.offsets), st), "st3r_gs_offsets")
        _lib.check(lib.st3r_gs_set_raster_variant(int(RASTER_VARIANT)), "st3r_gs_set_raster_variant")
        with _Prof("st3r_gs_raster_fwd"):
            _lib.check(lib.st3r_gs_raster_fwd(_lib.ptr(fr.offsets), _lib.ptr(n_dev), _lib.ptr(fr.vals),
                                              _lib.ptr(fr.geomA), _lib.ptr(fr.geomB), _lib.ptr(fr.rgb), C, fr.W, fr.H, TILE,
                                              _lib.ptr(fr.render), _lib.ptr(fr.alphas), _lib.ptr(fr.last_ids),
                                              _lib.ptr(fr.n_blend), st), "st3r_gs_raster_fwd")
    return fr


def _backward(fr, means, quats, scales, opacities, shN, v_render, v_alphas, reg_opac=0.0, reg_scale=0.0,
              reg_sums=None, plan=None):
    """Blend backward + projection/SH backward.  Returns (v_means, v_quats, v_scales, v_opacities, v_sh [N,4,3])."""
    lib = _lib.load()
    dev = means.device
    N, C = fr.N, fr.C
    E = max(C * N, 1)
    if plan is not None:
        v_geom = plan.v_geom.zero_()
        g = plan.peer.grads() if plan.peer is not None else plan.grads
        v_means, v_quats, v_scales, v_opac, v_sh = g["means"], g["quats"], g["scales"], g["opacities"], g["sh"]
    else:
        v_geom = torch.zeros((3, E, 4), dtype=torch.float32, device=dev)
        v_means = torch.empty((N, 3), dtype=torch.float32, device=dev)
        v_quats = torch.empty((N, 4), dtype=torch.float32, device=dev)
        v_scales = torch.empty((N, 3), dtype=torch.float32, device=dev)
        v_opac = torch.empty((N,), dtype=torch.float32, device=dev)
        v_sh = torch.empty((N, 4, 3), dtype=torch.float32, device=dev)
    st = _lib.stream_ptr()
    _lib.check(lib.st3r_gs_set_raster_variant(int(RASTER_VARIANT)), "st3r_gs_set_raster_variant")
    with torch.cuda.device(dev):
        with _Prof("st3r_gs_raster_bwd"):
            _lib.check(lib.st3r_gs_raster_bwd(_lib.ptr(fr.offsets), _lib.ptr(fr.n_isect_dev), _lib.ptr(fr.vals),
                                              _lib.ptr(fr.geomA), _lib.ptr(fr.geomB), _lib.ptr(fr.rgb), C, fr.W, fr.H, TILE,
                                              _lib.ptr(fr.alphas), _lib.ptr(fr.last_ids), _lib.ptr(v_render),
                                              _lib.ptr(v_alphas), _lib.ptr(v_geom[0]), _lib.ptr(v_geom[1]),
                                              _lib.ptr(v_geom[2]), st), "st3r_gs_raster_bwd")
        with _Prof("st3r_gs_project_bwd"):
            _lib.check(lib.st3r_gs_project_bwd(_lib.ptr(means), _lib.ptr(quats), _lib.ptr(scales), _lib.ptr(opacities),
                                               _lib.ptr(shN), fr.sh_coeffs, _lib.ptr(fr.cams), N, C, fr.W, fr.H, _f(EPS2D),
                                               _f(NEAR), _f(FAR), _f(RADIUS_CLIP), _lib.ptr(fr.radii), _lib.ptr(v_geom[0]),
                                               _lib.ptr(v_geom[1]), _lib.ptr(v_geom[2]), _f(reg_opac), _f(reg_scale),
                                               _lib.ptr(v_means), _lib.ptr(v_quats), _lib.ptr(v_scales), _lib.ptr(v_opac),
                                               _lib.ptr(v_sh), _lib.ptr(reg_sums), st), "st3r_gs_project_bwd")
    return v_means, v_quats, v_scales, v_opac, v_sh


def _info(fr, opacities):
    """The `info` dict of gsplat.rasterization (packed=True) assembled from the dense device state
    (index plumbing only)."""
    C, N = fr.C, fr.N
    vis = fr.radii[:C * N] > 0
    ids = torch.nonzero(vis).squeeze(1)
    packed_index = torch.cumsum(vis.to(torch.int32), 0, dtype=torch.int32) - 1
    gaussian_ids = ids % max(N, 1)
    n = fr.n_isect
    return {
        "camera_ids": ids // max(N, 1), "gaussian_ids": gaussian_ids, "radii": fr.radii[ids],
        "means2d": fr.geomA[ids, 0:2], "depths": fr.geomA[ids, 3], "conics": fr.geomB[ids, 0:3],
        "opacities": opacities.detach()[gaussian_ids], "tile_width": fr.tile_w, "tile_height": fr.tile_h,
        "tiles_per_gauss": fr.tiles[ids], "isect_ids": fr.keys[:n],
        "flatten_ids": packed_index[fr.vals[:n].long()], "isect_offsets": fr.offsets.reshape(C, fr.tile_h, fr.tile_w),
        "width": fr.W, "height": fr.H, "tile_size": TILE, "n_cameras": C, "last_ids": fr.last_ids,
    }


class _Rasterize(torch.autograd.Function):
    last_frame = None   # device state of the most recent forward (read by rasterization() to build `info`)

    @staticmethod
    def forward(ctx, means, quats, scales, opacities, colors, cams, width, height):
        args = [t.detach().float().contiguous() for t in (means, quats, scales, opacities, colors)]
        fr = _forward(*args, cams, width, height)
        ctx.fr, ctx.args = fr, args
        _Rasterize.last_frame = fr
        ctx.mark_non_differentiable(fr.last_ids)
        return fr.render, fr.alphas.unsqueeze(-1), fr.last_ids

    @staticmethod
    def backward(ctx, v_render, v_alpha, _):
        fr = ctx.fr
        v_render = v_render.float().contiguous()
        v_alpha = v_alpha.float().reshape(fr.C, fr.H, fr.W).contiguous()
        vm, vq, vs, vo, vsh = _backward(fr, *ctx.args, v_render, v_alpha)
        v_colors = torch.zeros((fr.N, fr.sh_coeffs, 3), dtype=torch.float32, device=vm.device)
        v_colors[:, :4] = vsh
        return vm, vq, vs, vo, v_colors, None, None, None


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height, sh_degree=1, **kw):
    """gsplat.rasterization as Starst3r calls it (gs.py:76-87): returns (render [C,H,W,3], alpha [C,H,W,1], info).
    Differentiable w.r.t. means / quats / scales / opacities / colors."""
    if sh_degree != 1:
        raise NotImplementedError("the Starst3r hot path renders with sh_degree=1 (gs.py:86); other degrees are "
                                  "outside the B200 path")
    if kw:
        raise NotImplementedError(f"unsupported gsplat.rasterization options {sorted(kw)}: Starst3r uses the defaults")
    _lib.require_cuda(means, quats, scales, opacities, colors)      # raises: there is no CPU fallback
    cams = make_cams(viewmats.to(means.device), Ks.to(means.device))
    render, alpha, last_ids = _Rasterize.apply(means, quats, scales, opacities, colors, cams, int(width), int(height))
    info = _info(_Rasterize.last_frame, opacities)
    return render, alpha, info


# ------------------------------------------------------------------------------------------- optimiser
class FusedAdam:
    """Optimizer-like handle for one splat tensor (what scene.optimizers[k] holds; gs.py:37).  Exposes
    param_groups / state like torch.optim.Adam; step() runs the fused CUDA Adam on `param.grad`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.param_groups = [dict(params=list(params), lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)]
        self.state = {}

    def _st(self, p):
        if p not in self.state:
            self.state[p] = dict(step=torch.zeros((), dtype=torch.float32), exp_avg=torch.zeros_like(p.data),
                                 exp_avg_sq=torch.zeros_like(p.data))
        return self.state[p]

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None if set_to_none else (p.grad.zero_() if p.grad is not None else None)

    def step(self):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                st = self._st(p)
                st["step"] += 1
                n = p.numel()
                adam_step([(p.data.view(-1), p.grad.contiguous().view(-1), st["exp_avg"].view(-1),
                            st["exp_avg_sq"].view(-1), 1, n, n, n)], g["lr"], g["betas"], g["eps"], int(st["step"].item()))


def adam_step(segments, lr, betas, eps, step, steps_done=None):
    """segments: list of (param, grad, exp_avg, exp_avg_sq, rows, cols, ld_param, ld_grad) device tensors.
    steps_done (device int32 tensor) instead of `step`: the step number lives on the device (st3r_adam_step_dev)."""
    lib = _lib.load()
    n = len(segments)
    PP = ctypes.c_void_p * n
    II = ctypes.c_int * n
    p = PP(*[s[0].data_ptr() for s in segments])
    g = PP(*[s[1].data_ptr() for s in segments])
    m = PP(*[s[2].data_ptr() for s in segments])
    v = PP(*[s[3].data_ptr() for s in segments])
    rows, cols = II(*[int(s[4]) for s in segments]), II(*[int(s[5]) for s in segments])
    ldp, ldg = II(*[int(s[6]) for s in segments]), II(*[int(s[7]) for s in segments])
    dev = segments[0][0].device
    with torch.cuda.device(dev), _Prof("st3r_adam_step"):
        if steps_done is not None:
            rc = lib.st3r_adam_step_dev(n, p, g, m, v, rows, cols, ldp, ldg, ctypes.c_double(lr), ctypes.c_double(betas[0]),
                                        ctypes.c_double(betas[1]), ctypes.c_double(eps), _lib.ptr(steps_done),
                                        _lib.stream_ptr())
        else:
            rc = lib.st3r_adam_step(n, p, g, m, v, rows, cols, ldp, ldg, ctypes.c_double(lr), ctypes.c_double(betas[0]),
                                    ctypes.c_double(betas[1]), ctypes.c_double(eps),
                                    int(step), _lib.stream_ptr())
    _lib.check(rc, "st3r_adam_step")


def adam_step_peers(segments, grad_offsets, peer_bases, lr, betas, eps, step):
    """Fused (sum of the gradients over all ranks' symmetric buffers) + Adam; see st3r_adam_step_peers."""
    lib = _lib.load()
    n, w = len(segments), len(peer_bases)
    PP, II, LL = ctypes.c_void_p * n, ctypes.c_int * n, ctypes.c_longlong * n
    p = PP(*[s[0].data_ptr() for s in segments])
    m = PP(*[s[2].data_ptr() for s in segments])
    v = PP(*[s[3].data_ptr() for s in segments])
    rows, cols = II(*[int(s[4]) for s in segments]), II(*[int(s[5]) for s in segments])
    ldp, ldg = II(*[int(s[6]) for s in segments]), II(*[int(s[7]) for s in segments])
    offs = LL(*[int(o) for o in grad_offsets])
    bases = (ctypes.c_void_p * w)(*[int(b) for b in peer_bases])
    dev = segments[0][0].device
    with torch.cuda.device(dev), _Prof("st3r_adam_step_peers"):
        rc = lib.st3r_adam_step_peers(n, p, offs, m, v, rows, cols, ldp, ldg, w, bases, ctypes.c_double(lr),
                                      ctypes.c_double(betas[0]), ctypes.c_double(betas[1]), ctypes.c_double(eps),
                                      int(step), _lib.stream_ptr())
    _lib.check(rc, "st3r_adam_step_peers")


class MCMCStrategy:
    """gsplat.MCMCStrategy as Starst3r drives it (gs.py:43-45 construction with defaults, :146-147
    step_pre_backward, :163-164 step_post_backward(..., lr=1e-3)); gsplat 1.4 strategy/mcmc.py + strategy/ops.py
    semantics.  The numeric work runs in csrc/gs_mcmc.cu; the random draws come from
    torch.multinomial / torch.randn_like exactly where gsplat takes them, so the RNG stream is the reference's.
    Like gsplat it assumes logit opacities and log scales (the reference renders them raw: quirk C-1, preserved)."""

    N_MAX = 51

    def __init__(self, cap_max=1_000_000, noise_lr=5e5, refine_start_iter=500, refine_stop_iter=25_000,
                 refine_every=100, min_opacity=0.005, verbose=False):
        self.cap_max, self.noise_lr = cap_max, noise_lr
        self.refine_start_iter, self.refine_stop_iter, self.refine_every = refine_start_iter, refine_stop_iter, refine_every
        self.min_opacity, self.verbose = min_opacity, verbose

    def check_sanity(self, params, optimizers):
        assert set(params.keys()) == set(optimizers.keys()), "params and optimizers must have the same keys"
        for k in ("means", "scales", "quats", "opacities"):
            assert k in params, f"{k} is required in params but missing."

    def initialize_state(self):
        binoms = torch.zeros((self.N_MAX, self.N_MAX))
        for n in range(self.N_MAX):
            for k in range(n + 1):
                binoms[n, k] = math.comb(n, k)
        return {"binoms": binoms}

    def step_pre_backward(self, params, optimizers, state, step, info):
        pass

    def step_post_backward(self, params, optimizers, state, step, info, lr):
        dev = _lib.require_cuda_device(params["means"].device, "starst3r_b200.gs.MCMCStrategy")   # no CPU fallback
        binoms = state["binoms"] = state["binoms"].to(dev).contiguous()
        if self.refine_start_iter < step < self.refine_stop_iter and step % self.refine_every == 0:
            n_relocated = self._relocate_gs(params, optimizers, binoms)
            n_new = self._add_new_gs(params, optimizers, binoms)
            if self.verbose:
                print(f"Step {step}: Relocated {n_relocated} GSs. Added {n_new} GSs. Now having {len(params['means'])} GSs.")
        inject_noise_to_position(params, lr * self.noise_lr)

    # ---- gsplat strategy/ops.py: relocate / sample_add
    def _relocate_gs(self, params, optimizers, binoms, sampled=None):
        lib = _lib.load()
        opac = params["opacities"].data
        dev, N = opac.device, opac.shape[0]
        dead = torch.empty(N, dtype=torch.int32, device=dev)
        alive = torch.empty(N, dtype=torch.int32, device=dev)
        probs = torch.empty(N, dtype=torch.float32, device=dev)
        alive_probs = torch.empty(N, dtype=torch.float32, device=dev)
        n_dead_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = _ws(lib.st3r_mcmc_partition_ws_bytes(N), dev)
        with torch.cuda.device(dev):
            _lib.check(lib.st3r_mcmc_partition(_lib.ptr(opac), N, _f(self.min_opacity), _lib.ptr(dead), _lib.ptr(alive),
                                               _lib.ptr(probs), _lib.ptr(alive_probs), _lib.ptr(n_dead_dev), _lib.ptr(ws),
                                               ws.numel(), _lib.stream_ptr()), "st3r_mcmc_partition")
        n_dead = int(n_dead_dev.item())      # gsplat reads dead_mask.sum().item() at the same place
        if n_dead == 0:
            return 0
        if n_dead == N:
            raise RuntimeError("MCMCStrategy: every Gaussian is dead (sigmoid(opacity) <= min_opacity); nothing to sample")
        if sampled is None:
            sampled = _multinomial_sample(alive_probs[:N - n_dead], n_dead)
        _relocate_rows(params, optimizers, sampled, alive, dead, 0, n_dead, N, binoms, self.min_opacity, True)
        _rewrap(params, optimizers, {k: v.data for k, v in params.items()}, None)
        return n_dead

    def _add_new_gs(self, params, optimizers, binoms, sampled=None):
        N = len(params["means"])
        n_target = min(self.cap_max, int(1.05 * N))
        n = max(0, n_target - N)
        if n == 0:
            return 0
        if sampled is None:
            sampled = _multinomial_sample(torch.sigmoid(params["opacities"].data.flatten()), n)
        grown = {}
        for k, p in params.items():
            g = torch.empty((N + n, *p.shape[1:]), dtype=p.dtype, device=p.device)
            g[:N].copy_(p.data)
            grown[k] = g
        tmp = {k: torch.nn.Parameter(v, requires_grad=False) for k, v in grown.items()}
        _relocate_rows(tmp, None, sampled, None, None, N, n, N, binoms, self.min_opacity, False)
        _rewrap(params, optimizers, grown, n)
        return n


# kept for callers of the round-1 name
MCMCStrategyStub = MCMCStrategy


def _multinomial_sample(weights, n, replacement=True):
    """gsplat ops._multinomial_sample: torch.multinomial up to 2**24 categories, numpy beyond."""
    if weights.shape[0] <= 2 ** 24:
        return torch.multinomial(weights, n, replacement=replacement)
    import numpy as np
    w = (weights / weights.sum()).detach().cpu().numpy()
    idx = np.random.choice(weights.shape[0], size=n, p=w, replace=replacement)
    return torch.from_numpy(idx).to(weights.device)


def _adam_state(opt, p):
    return opt.state.get(p, None)


def _relocate_rows(params, optimizers, sampled, alive, dst, dst_base, n, N, binoms, min_opacity, zero_moments):
    """One st3r_mcmc_relocate call over every parameter tensor (and, for relocate, every Adam moment tensor)."""
    lib = _lib.load()
    dev = params["means"].device
    sampled = sampled.to(device=dev, dtype=torch.int64).contiguous()
    assert sampled.numel() == n
    rows = [(k, p.data) for k, p in params.items() if k not in ("opacities", "scales")]
    moments = []
    if zero_moments and optimizers is not None:
        for k, p in params.items():
            st = _adam_state(optimizers[k], p)
            if st:
                moments += [st["exp_avg"], st["exp_avg_sq"]]
    for _, t in rows:
        assert t.is_contiguous() and t.dtype == torch.float32
    counts = torch.empty(max(N, 1), dtype=torch.int32, device=dev)

    def table(tensors):
        k = len(tensors)
        return (ctypes.c_void_p * max(k, 1))(*[t.data_ptr() for t in tensors]), \
               (ctypes.c_int * max(k, 1))(*[t[0].numel() for t in tensors])

    rp, rc = table([t for _, t in rows])
    mp, mc = table(moments)
    with torch.cuda.device(dev):
        _lib.check(lib.st3r_mcmc_relocate(_lib.ptr(params["opacities"].data), _lib.ptr(params["scales"].data),
                                          len(rows), rp, rc, len(moments), mp, mc, _lib.ptr(sampled), _lib.ptr(alive),
                                          _lib.ptr(dst), dst_base, n, N, _lib.ptr(binoms), binoms.shape[0],
                                          _f(min_opacity), _lib.ptr(counts), _lib.stream_ptr()), "st3r_mcmc_relocate")


def _rewrap(params, optimizers, new_data, n_added):
    """gsplat ops._update_param_with_optimizer: every entry of `params` becomes a fresh nn.Parameter and the
    optimiser state moves to it (moments grown by `n_added` zero rows for sample_add)."""
    for k in list(params.keys()):
        old = params[k]
        new = torch.nn.Parameter(new_data[k], requires_grad=old.requires_grad)
        opt = optimizers[k]
        st = opt.state.pop(old, None)
        if st is not None:
            if n_added:
                for key in ("exp_avg", "exp_avg_sq"):
                    z = torch.zeros((n_added, *st[key].shape[1:]), dtype=st[key].dtype, device=st[key].device)
                    st[key] = torch.cat([st[key], z])
            opt.state[new] = st
        for g in opt.param_groups:
            g["params"] = [new if q is old else q for q in g["params"]]
        params[k] = new


def inject_noise_to_position(params, scaler, noise=None):
    """gsplat ops.inject_noise_to_position (fused in st3r_mcmc_inject_noise); `noise` defaults to
    torch.randn_like(means), the draw gsplat makes."""
    lib = _lib.load()
    means = params["means"].data
    if noise is None:
        noise = torch.randn_like(means)
    with torch.cuda.device(means.device):
        _lib.check(lib.st3r_mcmc_inject_noise(_lib.ptr(means), _lib.ptr(params["quats"].data.contiguous()),
                                              _lib.ptr(params["scales"].data.contiguous()),
                                              _lib.ptr(params["opacities"].data.contiguous()), _lib.ptr(noise.contiguous()),
                                              means.shape[0], _f(scaler), _lib.stream_ptr()), "st3r_mcmc_inject_noise")


def compute_relocation(opacities, scales, ratios, binoms):
    """gsplat.relocation.compute_relocation (activated opacities [N], scales [N,3], ratios [N] int)."""
    lib = _lib.load()
    _lib.require_cuda(opacities, scales, ratios, binoms)
    N = opacities.shape[0]
    new_o = torch.empty_like(opacities, dtype=torch.float32)
    new_s = torch.empty_like(scales, dtype=torch.float32)
    with torch.cuda.device(opacities.device):
        _lib.check(lib.st3r_mcmc_compute_relocation(_lib.ptr(opacities.float().contiguous()), _lib.ptr(scales.float().contiguous()),
                                                    _lib.ptr(ratios.to(torch.int32).contiguous()),
                                                    _lib.ptr(binoms.float().contiguous()), binoms.shape[0], N,
                                                    _lib.ptr(new_o), _lib.ptr(new_s), _lib.stream_ptr()),
                   "st3r_mcmc_compute_relocation")
    return new_o, new_s


# ------------------------------------------------------------------------------------------- reference API
def init_3dgs(scene, init_scale=3e-3, lr=1e-3):
    """gs.py:14-45: splats from the dense MASt3R points (raw scales / opacities, wxyz identity quats, SH = 1 - colour)."""
    pts = scene.dense_pts_flat
    colors = scene.dense_cols_flat
    n = pts.shape[0]
    g = {
        "means": pts.detach().clone().float(),
        "scales": torch.full((n, 3), float(init_scale)),
        "quats": torch.zeros(n, 4),
        "opacities": torch.ones(n),
        "sh0": torch.zeros(n, 1, 3),
        "shN": torch.zeros(n, 24, 3),
    }
    g["quats"][:, 0] = 1.0
    col = (1 - colors).float().cpu()
    g["sh0"][:, 0] = col
    g["shN"][:] = col[:, None, :]
    scene.gaussians = {k: torch.nn.Parameter(v.to(scene.device).contiguous()) for k, v in g.items()}
    scene.optimizers = {k: FusedAdam([v], lr=lr) for k, v in scene.gaussians.items()}
    scene.ssim = None          # the SSIM term lives inside the fused loss kernel (st3r_gs_loss_fwd / _bwd)
    scene.strategy = MCMCStrategy()
    scene.strategy.check_sanity(scene.gaussians, scene.optimizers)
    scene.strategy_state = scene.strategy.initialize_state()
    scene._gs_truth = None


def render_3dgs(scene, w2c, intrinsics, width, height):
    """gs.py:47-88."""
    gz = scene.gaussians
    return rasterization(means=gz["means"], quats=gz["quats"], scales=gz["scales"], opacities=gz["opacities"],
                         colors=gz["shN"], viewmats=w2c, Ks=intrinsics, width=width, height=height, sh_degree=1)


def render_3dgs_original(scene, width, height):
    """gs.py:90-95."""
    return scene.render_3dgs(scene.w2c, scene.intrinsics, width, height)


def render_3dgs_path(scene, c2w_start, c2w_end, steps, intrinsics, width, height, chunk=32):
    """Novel-view fly-through (what the reference's demo renders frame by frame with
    utils.interp_se3_path + render_3dgs): the interpolated cameras are rendered `chunk` at a time as ONE batched
    rasterization call each (the kernels are batched over cameras).  Returns [steps, H, W, 3] on the device."""
    from .utils import interp_se3_path
    dev = scene.gaussians["means"].device
    path = interp_se3_path(c2w_start.detach().cpu().float(), c2w_end.detach().cpu().float(), steps)
    w2c = torch.linalg.inv(path).to(dev)
    K = intrinsics.to(dev).float()
    K = K.expand(steps, 3, 3) if K.dim() == 2 else K
    frames = []
    with torch.no_grad():
        for i in range(0, steps, chunk):
            img, _, _ = render_3dgs(scene, w2c[i:i + chunk], K[i:i + chunk], width, height)
            frames.append(img)
    return torch.cat(frames)


def _truth_images(scene, device):
    """Ground-truth images as one [C,H,W,3] device tensor (the reference re-uploads them every step, gs.py:151)."""
    cached = getattr(scene, "_gs_truth", None)
    if cached is None or cached.shape[0] != len(scene.imgs):
        imgs = [torch.as_tensor(im, dtype=torch.float32) for im in scene.imgs]
        cached = torch.stack(imgs).to(device).contiguous()
        scene._gs_truth = cached
    return cached


def train_step(params, states, truth, cams, width, height, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
               loss_ssim_fac=0.2, loss_opacity_fac=0.01, loss_scale_fac=0.01, count_blends=False, grad_hook=None,
               plan=None, loss_out=None):
    """One fused iteration of gs.py:143-161 (without the strategy hooks): render all C views, loss, backward, Adam.
    params: dict of contiguous fp32 device tensors (means, scales, quats, opacities, shN), updated in place;
    states: dict name -> (exp_avg, exp_avg_sq).  With a TrainPlan the iteration allocates nothing and never
    synchronises with the host, and once the plan's buffers are sized it is one CUDA-graph replay (TRAIN_GRAPH).
    Returns (loss tensor [] on device, frame)."""
    if (TRAIN_GRAPH and plan is not None and (plan.peer is None or plan.peer.scatter) and grad_hook is None
            and not count_blends and PROF is None and params["means"].is_cuda):
        hyper = (float(lr), (float(betas[0]), float(betas[1])), float(eps), float(loss_ssim_fac), float(loss_opacity_fac),
                 float(loss_scale_fac))
        out = plan.graph_step(params, states, truth, cams, width, height, int(step), hyper, loss_out)
        if out is not None:
            return out
    return _train_step_eager(params, states, truth, cams, width, height, step, lr, betas, eps, loss_ssim_fac,
                             loss_opacity_fac, loss_scale_fac, count_blends, grad_hook, plan, loss_out)


def _train_step_eager(params, states, truth, cams, width, height, step, lr, betas, eps, loss_ssim_fac, loss_opacity_fac,
                      loss_scale_fac, count_blends, grad_hook, plan, loss_out):
    """The iteration launch by launch.  step=None (graph capture): Adam reads the step from plan.steps_done."""
    lib = _lib.load()
    means, quats, scales, opac, shN = (params[k] for k in ("means", "quats", "scales", "opacities", "shN"))
    dev = means.device
    fr = _forward(means, quats, scales, opac, shN, cams, width, height, count_blends, plan)
    C, N, H, W = fr.C, fr.N, fr.H, fr.W
    if plan is not None:
        acc = plan.acc.zero_()
        dmaps, v_render = plan.dmaps, plan.v_render
    else:
        acc = torch.zeros((2 * C + 2,), dtype=torch.float32, device=dev)
        dmaps = torch.empty((C, H, W, 3, 3), dtype=torch.float32, device=dev)
        v_render = torch.empty_like(fr.render)
    sums, reg = acc[:2 * C], acc[2 * C:]
    st = _lib.stream_ptr()
    with torch.cuda.device(dev):
        with _Prof("st3r_gs_loss_fwd"):
            _lib.check(lib.st3r_gs_loss_fwd(_lib.ptr(fr.render), _lib.ptr(truth), C, H, W, _f(loss_ssim_fac),
                                            _lib.ptr(dmaps), _lib.ptr(sums), st), "st3r_gs_loss_fwd")
        with _Prof("st3r_gs_loss_bwd"):
            _lib.check(lib.st3r_gs_loss_bwd(_lib.ptr(fr.render), _lib.ptr(truth), _lib.ptr(dmaps), C, H, W,
                                            _f(loss_ssim_fac), _lib.ptr(v_render), st), "st3r_gs_loss_bwd")
    reg_o = C * loss_opacity_fac / max(N, 1)
    reg_s = C * loss_scale_fac / max(3 * N, 1)
    vm, vq, vs, vo, vsh = _backward(fr, means, quats, scales, opac, shN, v_render, None, reg_o, reg_s, reg, plan)
    fr.grads = dict(means=vm, quats=vq, scales=vs, opacities=vo, sh=vsh)
    if grad_hook is not None:      # multi-GPU: all-reduce of the per-Gaussian gradients (views are sharded)
        grad_hook(fr)
    segs = [(means, vm, *states["means"], N, 3, 3, 3), (scales, vs, *states["scales"], N, 3, 3, 3),
            (quats, vq, *states["quats"], N, 4, 4, 4), (opac, vo, *states["opacities"], N, 1, 1, 1),
            (shN, vsh, *states["shN"], N, 12, shN.shape[1] * 3, 12)]
    peer = plan.peer if plan is not None else None
    if peer is not None:
        # sharded views: one cross-device barrier, then the gradient sum over the ranks happens inside the Adam kernel
        with _Prof("peer_barrier"):
            peer.barrier()
        if peer.scatter:           # larger nodes: reduce-scatter + all-gather over peer memory, then a local Adam
            with _Prof("st3r_grad_reduce_scatter"):
                peer.reduce_scatter()
            with _Prof("peer_barrier"):
                peer.barrier()
            rd = peer.reduced
            segs = [(means, rd["means"], *states["means"], N, 3, 3, 3), (scales, rd["scales"], *states["scales"], N, 3, 3, 3),
                    (quats, rd["quats"], *states["quats"], N, 4, 4, 4), (opac, rd["opacities"], *states["opacities"], N, 1, 1, 1),
                    (shN, rd["sh"], *states["shN"], N, 12, shN.shape[1] * 3, 12)]
            adam_step(segs, lr, betas, eps, step, steps_done=plan.steps_done if step is None else None)
        else:
            adam_step_peers(segs, [peer.offsets[k] for k in ("means", "scales", "quats", "opacities", "sh")],
                            peer.peer_bases(), lr, betas, eps, step)
        if not plan._capturing:        # (a captured iteration: graph_step advances the parity after every replay)
            peer.advance()
    else:
        adam_step(segs, lr, betas, eps, step, steps_done=plan.steps_done if step is None else None)
    loss = loss_out if loss_out is not None else torch.empty((), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.st3r_gs_loss_finalize(_lib.ptr(sums), _lib.ptr(reg), C, H, W, _f(loss_ssim_fac), _f(reg_o), _f(reg_s),
                                             _lib.ptr(loss), st), "st3r_gs_loss_finalize")
    return loss, fr


def run_3dgs_optim(scene, iters, enable_pruning=False, loss_ssim_fac=0.2, loss_opacity_fac=0.01, loss_scale_fac=0.01,
                   verbose=False):
    """gs.py:97-166.  Returns the list of per-iteration losses (floats)."""
    gz = scene.gaussians
    dev = gz["means"].device
    if dev.type != "cuda":
        raise RuntimeError("starst3r_b200.gs.run_3dgs_optim needs the splats on a CUDA device (no CPU fallback)")
    height, width = scene.imgs[0].shape[:2]
    truth = _truth_images(scene, dev)
    cams = make_cams(scene.w2c.to(dev), scene.intrinsics.to(dev))
    names = ("means", "scales", "quats", "opacities", "shN")

    def opt_state(k):
        """Adam state of tensor k: FusedAdam's own record, or the `state` dict of a torch.optim.Adam the user swapped in
        (the reference lets users replace scene.optimizers[k], gs.py:37)."""
        opt = scene.optimizers[k]
        if isinstance(opt, FusedAdam):
            return opt._st(gz[k])
        st = opt.state[gz[k]]
        if not st:
            st.update(step=torch.zeros(()), exp_avg=torch.zeros_like(gz[k].data), exp_avg_sq=torch.zeros_like(gz[k].data))
        return st

    def bind():
        """Raw views of the splat tensors and their Adam moments (re-bound after the strategy replaced them)."""
        params = {k: gz[k].data for k in names}
        states = {}
        for k in names:
            st = opt_state(k)
            states[k] = (st["exp_avg"], st["exp_avg_sq"])
        return params, states

    params, states = bind()
    group = scene.optimizers["means"].param_groups[0]
    # the fused Adam kernel applies ONE (lr, betas, eps, step) to the five tensors of its launch
    for k in names:
        g_k = scene.optimizers[k].param_groups[0]
        if (g_k["lr"], tuple(g_k["betas"]), g_k["eps"]) != (group["lr"], tuple(group["betas"]), group["eps"]) or \
                g_k.get("weight_decay", 0) or g_k.get("amsgrad", False) or \
                int(opt_state(k)["step"]) != int(opt_state("means")["step"]):
            raise NotImplementedError(
                f"run_3dgs_optim: scene.optimizers[{k!r}] differs from scene.optimizers['means'] (lr / betas / eps / step "
                "count) or uses weight decay / amsgrad; the fused per-Gaussian Adam of this package applies one set of "
                "hyper-parameters to all splat tensors, as starster/gs.py:37 sets them up")
    losses_dev = torch.zeros(max(iters, 1), dtype=torch.float32, device=dev)
    # Sharded views: under a process group every rank renders the views i = rank mod G of the same,
    # replicated splat; the sum of the per-Gaussian gradients over the ranks is the full gradient (the loss is a sum
    # over views), so all replicas take the same Adam step.  Exchange: peer memory over NVLink fused with Adam
    # (dist.PeerGradExchange), NCCL all-reduce if symmetric memory cannot be set up.
    shard, hook = None, None
    if SHARD_VIEWS is True or (SHARD_VIEWS == "auto" and not enable_pruning):
        from . import dist as _sd
        rank, world_size = _sd.world()
        if world_size > 1 and cams.shape[0] >= world_size:
            if enable_pruning:
                raise NotImplementedError("run_3dgs_optim with sharded views keeps the splat replicated; the MCMC strategy "
                                          "draws random numbers per rank and would let the replicas diverge")
            shard = _sd.shard_indices(cams.shape[0], rank, world_size)
            truth = truth[shard].contiguous()
            cams = cams[shard].contiguous()
    n_views = cams.shape[0]

    def get_plan():
        plan = getattr(scene, "_gs_plan", None)
        if plan is None or not plan.matches(params["means"].shape[0], n_views, width, height, dev):
            plan = scene._gs_plan = TrainPlan(params["means"].shape[0], n_views, width, height, dev)
        return plan

    if shard is not None:
        # replicas must BE replicas: same number of Gaussians everywhere, rank 0's values and Adam moments on every rank
        # (the ranks may have arrived here through slightly different reconstructions)
        import torch.distributed as _dist
        n_here = torch.tensor([params["means"].shape[0], -params["means"].shape[0]], dtype=torch.int64, device=dev)
        _dist.all_reduce(n_here, op=_dist.ReduceOp.MAX)
        if int(n_here[0]) != -int(n_here[1]):
            raise RuntimeError(f"run_3dgs_optim with sharded views: the ranks hold different splats ({-int(n_here[1])} .. "
                               f"{int(n_here[0])} Gaussians); initialise the scene identically on every rank")
        for k in names:
            _dist.broadcast(params[k], 0)
            _dist.broadcast(states[k][0], 0)
            _dist.broadcast(states[k][1], 0)
        plan = get_plan()
        if plan.peer is None:
            try:
                plan.peer = _sd.PeerGradExchange(params["means"].shape[0], dev)
            except Exception as e:      # noqa: BLE001 - e.g. no symmetric memory on this node
                print(f"starst3r_b200.gs: symmetric memory unavailable ({e!r}); NCCL all-reduce of the gradients")
        if plan.peer is None:
            hook = lambda fr: _sd.allreduce_gradients(fr.grads)      # noqa: E731

    pbar = trange(iters, disable=not verbose)
    for step in pbar:
        if enable_pruning:
            scene.strategy.step_pre_backward(scene.gaussians, scene.optimizers, scene.strategy_state, step, None)
        opt_step = int(opt_state("means")["step"]) + 1
        loss, fr = train_step(params, states, truth, cams, width, height, opt_step, lr=group["lr"],
                              betas=group["betas"], eps=group["eps"], loss_ssim_fac=loss_ssim_fac,
                              loss_opacity_fac=loss_opacity_fac, loss_scale_fac=loss_scale_fac, plan=get_plan(),
                              loss_out=losses_dev[step], grad_hook=hook)
        for k in names:
            opt_state(k)["step"] += 1
        if verbose:
            pbar.set_description(f"Gsplat optimization: loss={loss.item()}")
        if enable_pruning:
            scene.strategy.step_post_backward(scene.gaussians, scene.optimizers, scene.strategy_state, step, None, 1e-3)
            if gz["means"].data.data_ptr() != params["means"].data_ptr() or gz["means"].shape[0] != params["means"].shape[0]:
                params, states = bind()
    if iters <= 0:
        return []
    get_plan().poll(wait_all=True)
    if shard is not None:               # the loss is a sum over views: add the ranks' shares
        import torch.distributed as _dist
        _dist.all_reduce(losses_dev)
    return losses_dev[:iters].cpu().tolist()
