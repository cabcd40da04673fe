"""MATCH + ALIGN pipeline — host-side mirror of starster/reconstruct.py (reconstruct_scene, run_sparse_ga,
sparse_scene_optimizer_slam) and of the mast3r/cloud_opt/sparse_ga.py functions it star-imports.

The numerics run in sm_100a kernels behind the C ABI: correspondences via st3r_extract_corres (match.py), the
700-iteration global alignment via st3r_align_optimize (align.cu), canonical views / dense points / point-cloud
cleaning via st3r_canonical_view, st3r_dense_points, st3r_clean_pointcloud (align_dense.cu).  Pair results stay in
HBM (no torch.save round trip); `tmpdir` / `cache_dir` are accepted for signature parity.
"""
import ctypes
import os
import math
from collections import namedtuple

import numpy as np
import torch

from . import _lib, match

__all__ = ("reconstruct_scene", "run_sparse_ga", "sparse_scene_optimizer_slam", "gamma_loss", "cosine_schedule",
           "linear_schedule", "make_pairs", "SparseGA", "PairOfSlices")

PairOfSlices = namedtuple("ImgPair", "img1, slice1, pix1, anchor_idxs1, img2, slice2, pix2, anchor_idxs2, confs, confs_sum")


# ------------------------------------------------------------------------------------------- small helpers
class gamma_loss:
    """cloud_opt/utils/losses.py:19-28 as a descriptor object: the fused kernel evaluates (d + o)^gamma - o^gamma
    itself, so only `gamma` is carried; calling the object evaluates the same formula with torch (API parity)."""

    def __init__(self, gamma, mul=1, offset=None, clip=np.inf):
        if mul != 1 or offset is not None or clip != np.inf:
            raise NotImplementedError("gamma_loss: only the default mul/offset/clip are used by Starst3r")
        self.gamma = float(gamma)
        self.offset = 0.0 if gamma == 1 else (1 / gamma) ** (1 / (gamma - 1))

    def __call__(self, x, y):
        d = torch.linalg.norm(x - y, dim=-1)
        return d if self.gamma == 1 else (d + self.offset) ** self.gamma - self.offset ** self.gamma


def linear_schedule(alpha, lr_base, lr_end=0):
    return (1 - alpha) * lr_base + alpha * lr_end


def cosine_schedule(alpha, lr_base, lr_end=0):
    return lr_end + (lr_base - lr_end) * (1 + np.cos(alpha * np.pi)) / 2


def _gamma_of(loss):
    if isinstance(loss, (int, float)):
        return float(loss)
    if hasattr(loss, "gamma"):
        return float(loss.gamma)
    raise NotImplementedError("the fused ALIGN kernel needs a gamma_loss(gamma) descriptor (or a number) as pixel loss")


def _sl(s):
    return s if isinstance(s, slice) else slice(s[1], s[2])


def make_pairs(imgs, scene_graph="complete", prefilter=None, symmetrize=True):
    """dust3r/image_pairs.py:11-59, 'complete' graph (the only one Starst3r uses, reconstruct.py:51)."""
    if scene_graph != "complete" or prefilter is not None:
        raise NotImplementedError("only scene_graph='complete', prefilter=None is on the Starst3r path")
    pairs = [(imgs[i], imgs[j]) for i in range(len(imgs)) for j in range(i)]
    if symmetrize:
        pairs += [(b, a) for a, b in pairs]
    return pairs


# ------------------------------------------------------------------------------------------- problem flattening
class _ImgConst(ctypes.Structure):
    _fields_ = [("W", ctypes.c_float), ("H", ctypes.c_float), ("base_focal", ctypes.c_float), ("median", ctypes.c_float),
                ("min_focal", ctypes.c_float), ("max_focal", ctypes.c_float), ("core_off", ctypes.c_int32),
                ("n_core", ctypes.c_int32)]


class _Problem(ctypes.Structure):
    _fields_ = [("n_img", ctypes.c_int32), ("img_const", ctypes.c_void_p), ("core", ctypes.c_void_p),
                ("n_core_total", ctypes.c_int32), ("root", ctypes.c_int32), ("edges", ctypes.c_void_p),
                ("n_anchor", ctypes.c_int32), ("anc_img", ctypes.c_void_p), ("anc_uv", ctypes.c_void_p),
                ("anc_k", ctypes.c_void_p), ("anc_off", ctypes.c_void_p),
                ("n3", ctypes.c_int32), ("e3_a1", ctypes.c_void_p), ("e3_a2", ctypes.c_void_p),
                ("e3_conf", ctypes.c_void_p), ("norm3", ctypes.c_float),
                ("n2", ctypes.c_int32), ("e2_img1", ctypes.c_void_p), ("e2_pix", ctypes.c_void_p),
                ("e2_a2", ctypes.c_void_p), ("e2_conf", ctypes.c_void_p), ("norm2", ctypes.c_float),
                ("nd", ctypes.c_int32), ("ed_a1", ctypes.c_void_p), ("ed_img2", ctypes.c_void_p),
                ("ed_tgt", ctypes.c_void_p), ("ed_conf", ctypes.c_void_p), ("normd", ctypes.c_float)]


def flatten_problem(imgs, imsizes, pps, base_focals, core_depth, anchors, corres, corres2d, preds_21, mst,
                    matching_conf_thr=5.0, device="cpu"):
    """condense_data output (sparse_ga.py:729-814) -> flat arrays for st3r_align_optimize (see St3rAlignProblem).
    Follows reconstruct.py:141-309: pp normalisation (:170), median-normalised core depth (:176-177), focal bounds
    (:203-205), matching_check (:283-290), loss3d / dust3r slice split and corres2d filtering (:291-309).
    Everything is assembled on `device` (the correspondence data is born there): the per-slice index ranges are expanded
    with one repeat_interleave per array instead of one arange per slice, and the host reads back twice - the per-image
    medians / per-slice maximum confidences that decide the layout, then the three confidence sums."""
    dev = torch.device(device)
    N = len(imgs)
    i32, i64 = torch.int32, torch.int64

    def f32(x):
        return torch.as_tensor(x).detach().to(dev, torch.float32)

    def cat(xs, dtype, shape_tail=()):
        if xs:
            return torch.cat(xs).to(dtype).contiguous()
        return torch.zeros((0,) + shape_tail, dtype=dtype, device=dev)

    def expand(starts, lens):
        """[starts[0], starts[0] + 1, ..., starts[0] + lens[0] - 1, starts[1], ...] as int32 on the device."""
        total = int(sum(lens))
        if total == 0:
            return torch.zeros(0, dtype=i32, device=dev)
        lens_t = torch.tensor(lens, dtype=i64, device=dev)
        base = torch.tensor(starts, dtype=i64, device=dev) - (torch.cumsum(lens_t, 0) - lens_t)
        return (torch.arange(total, device=dev) + torch.repeat_interleave(base, lens_t, output_size=total)).to(i32)

    def repeat(vals, lens):
        total = int(sum(lens))
        if total == 0:
            return torch.zeros(0, dtype=i32, device=dev)
        return torch.repeat_interleave(torch.tensor(vals, dtype=i32, device=dev), torch.tensor(lens, dtype=i64, device=dev),
                                       output_size=total)

    core = [f32(c).reshape(-1) for c in core_depth]
    med_t = torch.stack([c.median() for c in core])
    core_n = torch.cat([c / m for c, m in zip(core, med_t)]).contiguous()
    n_core = [int(c.numel()) for c in core]
    _, _, slices = corres
    smax = [f32(s[8]).max() for s in slices]
    # first read-back: medians, per-slice maximum confidence, image sizes, base focals
    head = torch.cat([med_t, torch.stack(smax) if smax else med_t[:0], f32(imsizes).reshape(-1), f32(base_focals).reshape(-1)]).cpu()
    S = len(slices)
    median, smax_h = head[:N].clone(), head[N:N + S]
    imsizes_f, base_f = head[N + S:N + S + 2 * N].reshape(N, 2).clone(), head[N + S + 2 * N:].clone()
    assert base_f.numel() == N
    diag = imsizes_f.norm(dim=1)
    img_names = list(imgs)
    ic_host = np.zeros(N, dtype=[("W", "f4"), ("H", "f4"), ("bf", "f4"), ("med", "f4"), ("minf", "f4"), ("maxf", "f4"),
                                 ("off", "i4"), ("n", "i4")])
    off = 0
    for i in range(N):
        ic_host[i] = (imsizes_f[i, 0], imsizes_f[i, 1], base_f[i], median[i], 0.25 * diag[i], 10 * diag[i], off, n_core[i])
        off += n_core[i]
    # anchors
    aoff = [0]
    for i in range(N):
        aoff.append(aoff[-1] + len(anchors[i][1]))
    n_anchor = aoff[-1]
    counts = [aoff[i + 1] - aoff[i] for i in range(N)]
    ok = {(s[0], s[4]): bool(smax_h[k] > matching_conf_thr) for k, s in enumerate(slices)}
    s3a, s3b, l3, e3c = [], [], [], []
    sda, ldd, vdi, edt, edc = [], [], [], [], []
    for s in slices:
        i1, sl1, i2, sl2, confs = s[0], _sl(s[1]), s[4], _sl(s[5]), s[8]
        if ok[i1, i2]:
            s3a.append(aoff[i1] + sl1.start); s3b.append(aoff[i2] + sl2.start); l3.append(sl1.stop - sl1.start)
            e3c.append(f32(confs))
        else:
            tgt, tc = preds_21[img_names[i2]][img_names[i1]]
            sda.append(aoff[i1]); ldd.append(counts[i1]); vdi.append(i2)
            edt.append(f32(tgt)); edc.append(f32(tc))
    v2i, s2a, l2, e2p, e2c = [], [], [], [], []
    for img1, pix1, confs, _, sls in corres2d:
        cur = 0
        for img2, sl2 in sls:
            sl2 = _sl(sl2)
            n = sl2.stop - sl2.start
            if ok[img1, img2]:
                v2i.append(img1); s2a.append(aoff[img2] + sl2.start); l2.append(n)
                e2p.append(f32(pix1[cur:cur + n])); e2c.append(f32(confs[cur:cur + n]))
            cur += n
    t = dict(
        img_const=torch.from_numpy(ic_host.view(np.uint8).reshape(N, -1).copy()).to(dev),
        core=core_n,
        edges=(torch.tensor([[int(a), int(b)] for a, b in mst[1]], dtype=i32).reshape(-1).contiguous()
               if len(mst[1]) else torch.zeros(0, dtype=i32)).to(dev),
        anc_img=repeat(list(range(N)), counts), anc_uv=cat([f32(anchors[i][0])[:, :2] for i in range(N)], torch.float32, (2,)),
        anc_k=cat([anchors[i][1].detach().to(dev) for i in range(N)], i32), anc_off=cat([f32(anchors[i][2]) for i in range(N)], torch.float32),
        e3_a1=expand(s3a, l3), e3_a2=expand(s3b, l3), e3_conf=cat(e3c, torch.float32),
        e2_img1=repeat(v2i, l2), e2_pix=cat(e2p, torch.float32, (2,)), e2_a2=expand(s2a, l2), e2_conf=cat(e2c, torch.float32),
        ed_a1=expand(sda, ldd), ed_img2=repeat(vdi, ldd), ed_tgt=cat(edt, torch.float32, (3,)), ed_conf=cat(edc, torch.float32))
    norms = torch.stack([t["e3_conf"].sum(), t["e2_conf"].sum(), t["ed_conf"].sum()]).tolist()      # second read-back
    meta = dict(N=N, root=int(mst[0]), n_anchor=n_anchor, aoff=aoff, n_core=n_core,
                norm3=float(norms[0]), norm2=float(norms[1]), normd=float(norms[2]),
                median=median, imsizes=imsizes_f, base_focals=base_f)
    return t, meta


def problem_struct(t, meta):
    p = _Problem()
    p.n_img, p.root, p.n_anchor = meta["N"], meta["root"], meta["n_anchor"]
    p.n_core_total = int(t["core"].numel())
    p.n3, p.n2, p.nd = int(t["e3_a1"].numel()), int(t["e2_img1"].numel()), int(t["ed_a1"].numel())
    p.norm3, p.norm2, p.normd = meta["norm3"], meta["norm2"], meta["normd"]
    for k in ("img_const", "core", "edges", "anc_img", "anc_uv", "anc_k", "anc_off", "e3_a1", "e3_a2", "e3_conf",
              "e2_img1", "e2_pix", "e2_a2", "e2_conf", "ed_a1", "ed_img2", "ed_tgt", "ed_conf"):
        setattr(p, k, t[k].data_ptr())
    return p


# Variants of the ALIGN kernels (bit mask of st3r_align_set_variant): bit 0 = segmented register accumulation +
# replicated gradient tables in the loss kernels and staged camera kernels, bit 1 = one thread-block cluster per image in
# the Weiszfeld focal kernel (B200, profiles/r02a_variants.json: reconstruct_scene 0.388 -> 0.356 s on 8 views 512 x 512),
# bit 2 = the whole optimisation loop of a phase as ONE cooperative launch (optimiser state replicated in every CTA's
# shared memory, entries packed once, one grid barrier per iteration; up to 64 images).  All on by default; 0 = the
# first implementation and 3 = the launch-per-iteration loop, kept as cross-checks (tests/test_align_gpu.py runs them).
# ST3R_ALIGN_VARIANT sets the initial value.
ALIGN_VARIANT = int(os.environ.get("ST3R_ALIGN_VARIANT", "7"))


def _optimize_phase(t, meta, params, mode, train_mask, gamma, lr_base, niter, schedule, dust3r_w, lossd_gamma,
                    want_grad=False):
    """One optimize_loop call (reconstruct.py:371-406) on the device."""
    lib = _lib.load()
    _lib.check(lib.st3r_align_set_variant(int(ALIGN_VARIANT)), "st3r_align_set_variant")
    dev = params["pps"].device
    N = meta["N"]
    prob = problem_struct(t, meta)
    lr = np.array([schedule(it / niter, lr_base, 0) for it in range(niter)], dtype=np.float32) if niter else \
        np.zeros(1, np.float32)
    adam_m = torch.zeros(N, 11, device=dev)
    adam_v = torch.zeros(N, 11, device=dev)
    loss_hist = torch.zeros(max(niter, 1), device=dev)
    camf = lib.st3r_align_cam_floats()
    cam = torch.empty(N, camf, device=dev)
    pts3d = torch.empty(max(meta["n_anchor"], 1), 3, device=dev)
    depth = torch.empty(max(int(t["core"].numel()), 1), device=dev)
    grad = torch.zeros(N, 11, device=dev) if want_grad else None
    ws = torch.empty(lib.st3r_align_ws_bytes_for(ctypes.byref(prob), niter), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.st3r_align_optimize(ctypes.byref(prob), _lib.ptr(params["pps"]), _lib.ptr(params["log_focals"]),
                                     _lib.ptr(params["quats"]), _lib.ptr(params["trans"]), _lib.ptr(params["log_sizes"]),
                                     _lib.ptr(adam_m), _lib.ptr(adam_v), mode, train_mask, ctypes.c_float(gamma),
                                     ctypes.c_float(lossd_gamma), ctypes.c_float(dust3r_w),
                                     lr.ctypes.data_as(ctypes.c_void_p), niter, ctypes.c_double(0.9),
                                     ctypes.c_double(0.9), ctypes.c_double(1e-8), _lib.ptr(loss_hist), _lib.ptr(cam),
                                     _lib.ptr(pts3d), _lib.ptr(depth), _lib.ptr(grad), _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr())
    _lib.check(rc, "st3r_align_optimize")
    K = torch.zeros(N, 3, 3, device=dev)
    K[:, 0, 0] = K[:, 1, 1] = cam[:, 12]
    K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = cam[:, 13], cam[:, 14], 1.0
    cam2w = torch.zeros(N, 4, 4, device=dev)
    cam2w[:, :3, :3] = cam[:, :9].reshape(N, 3, 3)
    cam2w[:, :3, 3] = cam[:, 9:12]
    cam2w[:, 3, 3] = 1.0
    aoff = meta["aoff"]
    coff = np.cumsum([0] + meta["n_core"])
    res = dict(intrinsics=K, cam2w=cam2w,
               depthmaps=[depth[coff[i]:coff[i + 1]] for i in range(N)],
               pts3d=[pts3d[aoff[i]:aoff[i + 1]] for i in range(N)])
    return res, loss_hist[:niter], grad


def _broadcast_alignment(params, res_coarse, res_fine, src=0):
    """Rank `src`'s optimiser state and results to every rank of the group: one flat fp32 broadcast."""
    import torch.distributed as dist
    tensors = [params[k] for k in ("pps", "log_focals", "quats", "trans", "log_sizes")]
    for res in (res_coarse, res_fine):
        if res is not None:
            tensors += [res["intrinsics"], res["cam2w"], *res["depthmaps"], *res["pts3d"]]
    flat = torch.cat([x.reshape(-1).float() for x in tensors]).contiguous()
    dist.broadcast(flat, src)
    off = 0
    for x in tensors:
        n = x.numel()
        x.copy_(flat[off:off + n].view_as(x))
        off += n


def sparse_scene_optimizer_slam(imgs, subsample, imsizes, pps, base_focals, core_depth, anchors, corres, corres2d,
                                preds_21, canonical_paths, mst, cache_path=None,
                                lr1=0.2, niter1=500, loss1=gamma_loss(1.1),
                                lr2=0.02, niter2=500, loss2=gamma_loss(0.4),
                                lossd=gamma_loss(1.1),
                                opt_pp=True, opt_depth=True,
                                schedule=cosine_schedule, depth_mode="add", exp_depth=False,
                                lora_depth=False, shared_intrinsics=False,
                                init={}, device="cuda", dtype=torch.float32,
                                matching_conf_thr=5., loss_dust3r_w=0.01,
                                verbose=True, dbg=(), prev_params=None):
    """starster/reconstruct.py:116-457 on the B200: same arguments, returns (imgs, res_coarse, res_fine, params_ret).
    Options Starst3r never enables (opt_depth=True, lora_depth, exp_depth, depth_mode='mul', shared_intrinsics,
    per-image `init`) are outside the fused kernel and raise NotImplementedError."""
    if opt_depth or lora_depth or exp_depth or depth_mode != "add" or shared_intrinsics or any(init.values()):
        raise NotImplementedError("sparse_scene_optimizer_slam on B200 implements the configuration Starst3r uses "
                                  "(reconstruct.py:56-70: opt_depth=False, depth_mode='add', no lora/exp depth, "
                                  "separate intrinsics, no per-image init)")
    dev = _lib.require_cuda_device(device, "sparse_scene_optimizer_slam")        # the fused optimiser has no CPU fallback
    assert len(mst[1]) == len(imgs) - 1
    N = len(imgs)
    t, meta = flatten_problem(imgs, imsizes, pps, base_focals, core_depth, anchors, corres, corres2d, preds_21, mst,
                              matching_conf_thr, dev)
    pps_n = (torch.as_tensor(pps).detach().float().cpu() / meta["imsizes"])
    params = dict(pps=pps_n.to(dev).contiguous(), log_focals=meta["base_focals"].log().to(dev).contiguous(),
                  quats=torch.tensor([[0.0, 0, 0, 1]]).repeat(N, 1).to(dev).contiguous(),
                  trans=torch.zeros(N, 3, device=dev), log_sizes=torch.zeros(N, device=dev))
    if verbose:
        print("init focals =", meta["base_focals"].numpy())
    if prev_params is not None:       # reconstruct.py:408-415 - warm start of the first len(prev) images
        for k in ("pps", "log_focals", "quats", "trans", "log_sizes"):
            prev = prev_params[k]
            n = min(len(prev), N)
            if n:
                params[k][:n] = torch.stack([torch.as_tensor(x).detach().reshape(-1) for x in prev[:n]]).to(dev).reshape(
                    params[k][:n].shape)
    # Under a process group the optimiser runs on rank 0 only and its result is broadcast (O(11 N) parameters,
    # strictly sequential iterations).  Replicas would NOT stay identical: the kernels sum losses and gradients with fp32
    # atomics, and Adam turns that rounding noise into an O(lr) walk along the gauge directions,
    # so cameras, dense points and the splat built from them could differ between ranks.
    group = _shard_group()
    lead = group is None or group[0] == 0
    n1, n2 = (niter1, niter2) if lead else (0, 0)          # the other ranks only shape the result buffers
    res_coarse, hist1, _ = _optimize_phase(t, meta, params, 0, 4 | 8 | 16, _gamma_of(loss1), lr1, n1, schedule,
                                           loss_dust3r_w, _gamma_of(lossd))
    if verbose and n1:
        print(f">> final loss = {float(hist1[-1])}")
    res_fine = None
    if niter2:
        mask = 4 | 8 | 16 | 2 | (1 if opt_pp else 0)
        res_fine, hist2, _ = _optimize_phase(t, meta, params, 1, mask, _gamma_of(loss2), lr2, n2, schedule,
                                             loss_dust3r_w, _gamma_of(lossd))
        if verbose and n2:
            print(f">> final loss = {float(hist2[-1])}")
    if group is not None:
        _broadcast_alignment(params, res_coarse, res_fine)
    if verbose:
        f = params["log_focals"].exp().clip(min=torch.as_tensor(0.25 * meta["imsizes"].norm(dim=1)).to(dev),
                                            max=torch.as_tensor(10 * meta["imsizes"].norm(dim=1)).to(dev))
        print("Final focals =", f.cpu().numpy())
    P = torch.nn.Parameter
    coff = np.cumsum([0] + meta["n_core"])
    params_ret = {
        "pps": [P(params["pps"][i].clone()) for i in range(N)],
        "log_focals": [P(params["log_focals"][i:i + 1].clone()) for i in range(N)],
        "quats": [P(params["quats"][i].clone()) for i in range(N)],
        "trans": [P(params["trans"][i].clone()) for i in range(N)],
        "log_sizes": [P(params["log_sizes"][i:i + 1].clone()) for i in range(N)],
        "core_depth": [P(t["core"][coff[i]:coff[i + 1]].clone(), requires_grad=False) for i in range(N)],
    }
    return imgs, res_coarse, res_fine, params_ret


# =================================================================================================================
# Pipeline around the optimiser: pair loop + matching, canonical views, MST, condense, SparseGA result object.
# Restated from mast3r/cloud_opt/sparse_ga.py; per-pair tensors are memoised in HBM instead of torch.save files.
# =================================================================================================================
_MEMO = {}   # cache_path -> {"fwd": {(a, b): (X11, C11, X21, C21)}, "corres": {(a, b): (score, (xy1, xy2, conf))},
             #                "canon": {img: ((canon, canon2, cconf), focal)}}


def _memo(cache_path):
    return _MEMO.setdefault(cache_path, {"fwd": {}, "corres": {}, "canon": {}})


# ---- optional disk mirror in the reference's own file formats -----------------------------------
# Default: everything streams from HBM.  With PERSIST_CACHE = True (and a cache_path) every memo entry is also
# written / looked up on disk exactly where and how the reference stores it, so a cache directory produced by
# either implementation can be consumed by the other:
#   forward/<md5 a>/<md5 b>.pth                     = (X11 [H,W,3], C11 [H,W], X21 [H,W,3], C21 [H,W])   sparse_ga.py:552
#   corres_conf=<desc_conf>_subsample=<S>/<a>-<b>.pth = ((conf_score, sum conf, n), (xy1, xy2, conf))    sparse_ga.py:561
#   canon_views/<md5 img>_subsample=<S>_kw=<kw>.pth  = ((canon, canon2, cconf), focal)                   sparse_ga.py:706
PERSIST_CACHE = False


def clear_cache(cache_path=None):
    """Releases the device-resident pair / canonical-view memo of `cache_path` (all of them when None).  The reference
    keeps these tensors in files under cache_dir (sparse_ga.py:533-561,643); here they live in HBM - about
    N (N - 1) 8 H W 4 bytes per reconstruction - until the owning Scene goes away or this is called."""
    if cache_path is None:
        _MEMO.clear()
    else:
        _MEMO.pop(cache_path, None)


def _md5(name):
    import hashlib
    return hashlib.md5(name.encode("utf-8")).hexdigest()


def _cache_file(cache_path, kind, a, b=None, desc_conf="desc_conf", subsample=8, kw=None):
    import os
    if kind == "fwd":
        return os.path.join(cache_path, "forward", _md5(a), _md5(b) + ".pth")
    if kind == "corres":
        return os.path.join(cache_path, f"corres_conf={desc_conf}_{subsample=}", f"{_md5(a)}-{_md5(b)}.pth")
    return os.path.join(cache_path, "canon_views", _md5(a) + f"_{subsample=}_{kw=}.pth")


def _disk_get(path, device):
    import os
    if not (PERSIST_CACHE and path and os.path.isfile(path)):
        return None
    return torch.load(path, map_location=device)


def _disk_put(path, obj):
    import os
    if not (PERSIST_CACHE and path):
        return

    def cpu(x):
        if isinstance(x, torch.Tensor):
            return x.detach().cpu()
        if isinstance(x, (tuple, list)):
            return type(x)(cpu(v) for v in x)
        return x
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(cpu(obj), path)


def convert_dust3r_pairs_naming(imgs, pairs_in):
    """sparse_ga.py:111-115."""
    for pair in pairs_in:
        for k in range(2):
            pair[k]["instance"] = imgs[pair[k]["idx"]]
    return pairs_in


def symmetric_inference(model, img1, img2, device):
    """sparse_ga.py:571-592: encoder once, decoder in both directions.  `model` is the (out-of-scope) MASt3R network;
    any object offering `symmetric_inference(img1, img2)` (e.g. synth.SyntheticMast3r) is used directly."""
    if hasattr(model, "symmetric_inference"):
        return model.symmetric_inference(img1, img2)
    shape1 = torch.from_numpy(img1["true_shape"]).to(device, non_blocking=True)
    shape2 = torch.from_numpy(img2["true_shape"]).to(device, non_blocking=True)
    im1 = img1["img"].to(device, non_blocking=True)
    im2 = img2["img"].to(device, non_blocking=True)
    feat1, feat2, pos1, pos2 = model._encode_image_pairs(im1, im2, shape1, shape2)

    def decoder(fa, fb, pa, pb, sa, sb):
        dec1, dec2 = model._decoder(fa, pa, fb, pb)
        with torch.autocast("cuda", enabled=False):
            return (model._downstream_head(1, [t.float() for t in dec1], sa),
                    model._downstream_head(2, [t.float() for t in dec2], sb))
    res11, res21 = decoder(feat1, feat2, pos1, pos2, shape1, shape2)
    res22, res12 = decoder(feat2, feat1, pos2, pos1, shape2, shape1)
    return res11, res21, res22, res12


def symmetric_inference_batch(model, imgs1, imgs2, device):
    """B image pairs through the network in ONE pass (sparse_ga.py:571-592 with a batch dimension): the
    images of the pairs are stacked along dim 0, the encoder / the two decoder directions / the heads run once, and the
    result is split back into the per-pair tuples (res11, res21, res22, res12) `symmetric_inference` returns, each with
    batch dimension 1.  A model that brings its own `symmetric_inference_batch(imgs1, imgs2)` (synth.SyntheticMast3r) is
    used directly; one that only has `symmetric_inference`, and pairs of different image sizes, go pair by pair."""
    assert len(imgs1) == len(imgs2)
    if hasattr(model, "symmetric_inference_batch"):
        return list(model.symmetric_inference_batch(imgs1, imgs2))
    same = all(tuple(i["img"].shape) == tuple(imgs1[0]["img"].shape) for i in list(imgs1) + list(imgs2))
    if hasattr(model, "symmetric_inference") or len(imgs1) == 1 or not same:
        return [symmetric_inference(model, a, b, device) for a, b in zip(imgs1, imgs2)]
    shape1 = torch.cat([torch.from_numpy(i["true_shape"]) for i in imgs1]).to(device, non_blocking=True)
    shape2 = torch.cat([torch.from_numpy(i["true_shape"]) for i in imgs2]).to(device, non_blocking=True)
    im1 = torch.cat([i["img"] for i in imgs1]).to(device, non_blocking=True)
    im2 = torch.cat([i["img"] for i in imgs2]).to(device, non_blocking=True)
    feat1, feat2, pos1, pos2 = model._encode_image_pairs(im1, im2, shape1, shape2)

    def decoder(fa, fb, pa, pb, sa, sb):
        dec1, dec2 = model._decoder(fa, pa, fb, pb)
        with torch.autocast("cuda", enabled=False):
            return (model._downstream_head(1, [t.float() for t in dec1], sa),
                    model._downstream_head(2, [t.float() for t in dec2], sb))
    res11, res21 = decoder(feat1, feat2, pos1, pos2, shape1, shape2)
    res22, res12 = decoder(feat2, feat1, pos2, pos1, shape2, shape1)
    return [tuple({k: v[i:i + 1] for k, v in r.items()} for r in (res11, res21, res22, res12)) for i in range(len(imgs1))]


def _pair_missing(memo, a, b):
    return not ((a, b) in memo["fwd"] and (b, a) in memo["fwd"] and (a, b) in memo["corres"])


def _pair_maps(memo, res, a, b, device, desc_conf):
    """Stores the four point / confidence maps of one pair; returns (descriptors, descriptor confidences, conf_score)."""
    X11, X21, X22, X12 = [r["pts3d"][0].to(device).float().contiguous() for r in res]
    C11, C21, C22, C12 = [r["conf"][0].to(device).float().contiguous() for r in res]
    descs = [r["desc"][0].to(device) for r in res]
    qonfs = [r[desc_conf][0].to(device) for r in res]
    memo["fwd"][a, b] = (X11, C11, X21, C21)
    memo["fwd"][b, a] = (X22, C22, X12, C12)
    conf_score = (C11.mean() * C12.mean() * C21.mean() * C22.mean()).sqrt().sqrt()
    return descs, qonfs, conf_score


def _compute_pair(memo, model, img1, img2, device, desc_conf, subsample):
    """Inference + matching of one unordered image pair (sparse_ga.py:541-561); fills the memo."""
    a, b = img1["instance"], img2["instance"]
    res = symmetric_inference(model, img1, img2, device)
    descs, qonfs, conf_score = _pair_maps(memo, res, a, b, device, desc_conf)
    corres = match.extract_correspondences(descs, qonfs, device=device, subsample=subsample)
    memo["corres"][a, b] = ((float(conf_score), float(corres[2].sum()), len(corres[2])), corres)


# SURVEY 8f-2 - the pair loop as a two-stream pipeline.  The reference (and _compute_pair) runs network -> matcher ->
# host read-back pair by pair (sparse_ga.py:541-561: `float(conf_score)`, `len(corres[2])` synchronise every pair).
# Here the network runs INFERENCE_BATCH pairs per pass on a side stream, one batch ahead of the matcher; the matcher
# (a replayed CUDA graph per pair, match.extract_correspondences_device) consumes the descriptor maps where the heads
# wrote them ([H, W, 24] fp32 = the [HW, 24] K-major layout the tcgen05 kernel's TMA descriptor reads), results are
# staged on the device and the host reads all counts and scores back ONCE, after the last pair.  Entries of the memo
# are bit-identical to the sequential loop (tests/test_match_gpu.py::test_pipelined_pair_loop_equals_sequential).
PIPELINE_PAIRS = True
INFERENCE_BATCH = 4
_SIDE_STREAMS = {}


def _compute_pairs(memo, model, todo, device, desc_conf, subsample):
    """Inference + matching of the image pairs `todo` = [(img1, img2), ...]; fills the memo."""
    dev = torch.device(device)
    if not (PIPELINE_PAIRS and dev.type == "cuda" and torch.cuda.is_available() and len(todo) > 1):
        for img1, img2 in todo:
            _compute_pair(memo, model, img1, img2, device, desc_conf, subsample)
        return
    # the first pair synchronously: its read-back is where the matcher picks its precision variant for this scene
    _compute_pair(memo, model, todo[0][0], todo[0][1], device, desc_conf, subsample)
    todo = todo[1:]
    cur = torch.cuda.current_stream(dev)
    side = _SIDE_STREAMS.get(dev)
    if side is None:
        side = _SIDE_STREAMS[dev] = torch.cuda.Stream(dev)
    side.wait_stream(cur)
    chunks = [todo[i:i + INFERENCE_BATCH] for i in range(0, len(todo), INFERENCE_BATCH)]

    def infer(chunk):
        with torch.cuda.stream(side):
            res = symmetric_inference_batch(model, [p[0] for p in chunk], [p[1] for p in chunk], device)
            ev = torch.cuda.Event()
            ev.record(side)
        return res, ev

    counts = torch.zeros(len(todo), dtype=torch.int32).pin_memory()
    staged, scores = [], []
    ahead = infer(chunks[0])
    for ci, chunk in enumerate(chunks):
        results, ev = ahead
        if ci + 1 < len(chunks):
            ahead = infer(chunks[ci + 1])                     # the network runs one batch ahead of the matcher
        cur.wait_event(ev)
        for (img1, img2), res in zip(chunk, results):
            a, b = img1["instance"], img2["instance"]
            descs, qonfs, conf_score = _pair_maps(memo, res, a, b, device, desc_conf)
            xy1, xy2, conf, n_out = match.extract_correspondences_device([d.float().contiguous() for d in descs],
                                                                         [q.float().contiguous() for q in qonfs], subsample)
            counts[len(staged):len(staged) + 1].copy_(n_out, non_blocking=True)
            staged.append((a, b, xy1.clone(), xy2.clone(), conf.clone()))    # (the plan's buffers serve the next pair)
            scores.append(conf_score)
    cur.synchronize()                                          # the one host synchronisation of the loop
    side.synchronize()
    match.adapt_variant()
    sums = []
    for k, (a, b, xy1, xy2, conf) in enumerate(staged):
        n = int(counts[k])
        staged[k] = (a, b, xy1[:n], xy2[:n], conf[:n])
        sums.append(conf[:n].sum())
    host = torch.stack(scores + sums).double().cpu().tolist()
    for k, (a, b, xy1, xy2, conf) in enumerate(staged):
        memo["corres"][a, b] = ((host[k], host[len(staged) + k], int(conf.numel())), (xy1, xy2, conf))


# Under an initialised torch.distributed group (one process per GPU) forward_mast3r computes only every G-th missing
# pair on this rank - image pairs are independent (sparse_ga.py:529 loop body) - and distributes every
# pair from the rank that computed it (_exchange_pair): what the replicated optimiser reads goes to everybody, an
# image's own full-resolution map only to the rank that owns the image.  No collective touches the matching itself.
SHARD_PAIRS = True


def _shard_group():
    """(rank, world size) when the work of this module is to be shared by the ranks of a process group, else None."""
    if not SHARD_PAIRS:
        return None
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return None


class _Sub:
    """A map this rank only holds at the optimiser's resolution: [::subsample, ::subsample] of the original."""

    def __init__(self, t, subsample):
        self.t, self.subsample = t, subsample


def _subsampled(x, subsample):
    if isinstance(x, _Sub):
        assert x.subsample == subsample, "map was exchanged at another subsampling"
        return x.t
    return x[::subsample, ::subsample]


def _exchange_pair(memo, a, b, ia, ib, owner, rank, world, device, subsample):
    """Distributes one computed pair from the rank that computed it (`owner`):
      * to everybody (the optimiser runs as replicas): the scores, the correspondence list and the two cross maps
        X21 / X12 at the optimiser's resolution ([::subsample, ::subsample]: all anybody reads of them, preds_21);
      * image a's own full-resolution map (X11, C11) only to the rank that owns image a (index mod G), image b's
        (X22, C22) only to the owner of b: they feed the canonical view of that image and nothing else.
    Per pair 8 HW floats travel to at most two ranks instead of 16 HW floats to all of them, and a rank keeps the
    full-resolution maps of its own images only (N / G images x 2 (N - 1) entries)."""
    import torch.distributed as dist
    head = torch.zeros(7, dtype=torch.float64, device=device)
    if rank == owner:
        score, (xy1, xy2, conf) = memo["corres"][a, b]
        X11, C11, X21, C21 = memo["fwd"][a, b]
        X22, C22, X12, C12 = memo["fwd"][b, a]
        head = torch.tensor([score[0], score[1], float(score[2]), X11.shape[0], X11.shape[1], X22.shape[0], X22.shape[1]],
                            dtype=torch.float64, device=device)
        small = torch.cat([_subsampled(t, subsample).reshape(-1) for t in (X21, C21, X12, C12)]).float().contiguous()
    dist.broadcast(head, owner)
    n, H1, W1, H2, W2 = (int(v) for v in head[2:].tolist())
    h1, w1 = len(range(0, H1, subsample)), len(range(0, W1, subsample))
    h2, w2 = len(range(0, H2, subsample)), len(range(0, W2, subsample))
    sizes = [3 * h2 * w2, h2 * w2, 3 * h1 * w1, h1 * w1]
    if rank != owner:
        small = torch.empty(sum(sizes), dtype=torch.float32, device=device)
        xy1 = torch.empty((n, 2), dtype=torch.int64, device=device)
        xy2 = torch.empty((n, 2), dtype=torch.int64, device=device)
        conf = torch.empty(n, dtype=torch.float32, device=device)
    xy = torch.cat([xy1.reshape(-1, 2), xy2.reshape(-1, 2)], dim=1).contiguous()        # [n, 4]
    conf = conf.contiguous()
    dist.broadcast(small, owner)
    if n > 0:                                   # (no zero-length collectives)
        dist.broadcast(xy, owner)
        dist.broadcast(conf, owner)
    # full-resolution own maps: point to point, owner of the pair -> owner of the image
    full = {}
    for tag, idx, H, W in (("a", ia, H1, W1), ("b", ib, H2, W2)):
        dst = idx % world
        if rank == owner:
            X, C = (X11, C11) if tag == "a" else (X22, C22)
            if dst == owner:
                full[tag] = (X, C)
            else:
                dist.send(torch.cat([X.reshape(-1), C.reshape(-1)]).float().contiguous(), dst)
        elif rank == dst:
            buf = torch.empty(4 * H * W, dtype=torch.float32, device=device)
            dist.recv(buf, owner)
            full[tag] = (buf[:3 * H * W].view(H, W, 3), buf[3 * H * W:].view(H, W))
    parts = torch.split(small, sizes)
    X21s, C21s = _Sub(parts[0].view(h2, w2, 3), subsample), _Sub(parts[1].view(h2, w2), subsample)
    X12s, C12s = _Sub(parts[2].view(h1, w1, 3), subsample), _Sub(parts[3].view(h1, w1), subsample)
    memo["fwd"][a, b] = full.get("a", (None, None)) + (X21s, C21s)
    memo["fwd"][b, a] = full.get("b", (None, None)) + (X12s, C12s)
    if rank != owner:
        memo["corres"][a, b] = ((float(head[0]), float(head[1]), n), (xy[:, :2].contiguous(), xy[:, 2:].contiguous(), conf))


def _forward_sharded(pairs, memo, model, device, desc_conf, subsample, cache_path=None):
    """Computes this rank's share of the missing pairs and exchanges all of them; returns the number computed here."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    todo, seen = [], set()
    for k, (img1, img2) in enumerate(pairs):
        a, b = img1["instance"], img2["instance"]
        if frozenset((a, b)) not in seen and _pair_missing(memo, a, b) and (b, a) not in memo["corres"]:
            seen.add(frozenset((a, b)))
            todo.append(k)
    box = [todo]
    dist.broadcast_object_list(box, src=0)        # rank 0's list decides (caches may differ between ranks)
    todo = box[0]
    fkw = dict(desc_conf=desc_conf, subsample=subsample)
    share = [pairs[k] for j, k in enumerate(todo) if j % world == rank]
    _compute_pairs(memo, model, share, device, desc_conf, subsample)
    for img1, img2 in share:
        a, b = img1["instance"], img2["instance"]
        if PERSIST_CACHE and cache_path:        # the rank that computed a pair is the one that still holds all of it
            _disk_put(_cache_file(cache_path, "fwd", a, b), memo["fwd"][a, b])
            _disk_put(_cache_file(cache_path, "fwd", b, a), memo["fwd"][b, a])
            _disk_put(_cache_file(cache_path, "corres", a, b, **fkw), memo["corres"][a, b])
    mine = len(share)
    for j, k in enumerate(todo):
        img1, img2 = pairs[k]
        _exchange_pair(memo, img1["instance"], img2["instance"], int(img1["idx"]), int(img2["idx"]), j % world, rank, world,
                       device, subsample)
    return mine


@torch.no_grad()
def forward_mast3r(pairs, model, cache_path, desc_conf="desc_conf", device="cuda", subsample=8, **matching_kw):
    """sparse_ga.py:524-568.  Returns ({(name1, name2): ((key1, key2), key_corres)}, cache_path) where the keys index
    the in-HBM memo (the reference returns file paths)."""
    memo = _memo(cache_path)
    res_paths = {}
    fkw = dict(desc_conf=desc_conf, subsample=subsample)
    if PERSIST_CACHE and cache_path:            # adopt entries another run (or the reference) left on disk
        for img1, img2 in pairs:
            a, b = img1["instance"], img2["instance"]
            for key, kind in (((a, b), "fwd"), ((b, a), "fwd"), ((a, b), "corres"), ((b, a), "corres")):
                store = memo["fwd" if kind == "fwd" else "corres"]
                if key not in store:
                    got = _disk_get(_cache_file(cache_path, kind, *key, **fkw), device)
                    if got is not None:
                        store[key] = tuple(got) if kind == "fwd" else (tuple(got[0]), tuple(got[1]))
    sharded = False
    if SHARD_PAIRS and model is not None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            _forward_sharded(pairs, memo, model, device, desc_conf, subsample, cache_path)
            sharded = True
    def mirror(a, b):
        if (b, a) in memo["corres"] and (a, b) not in memo["corres"]:
            score, (xy1, xy2, confs) = memo["corres"][b, a]
            memo["corres"][a, b] = (score, (xy2, xy1, confs))                        # :538-540
            if not sharded or dist.get_rank() == 0:
                _disk_put(cache_path and _cache_file(cache_path, "corres", a, b, **fkw), memo["corres"][a, b])

    # the pairs still missing, in the reference's loop order (a pair whose mirror image comes earlier in the list is
    # served by the mirror rule above, like in the reference), computed as one pipelined batch
    todo = []
    if model is not None:
        have_fwd, have_cor = set(memo["fwd"]), set(memo["corres"])
        for img1, img2 in pairs:
            a, b = img1["instance"], img2["instance"]
            if (b, a) in have_cor:
                have_cor.add((a, b))
            if not ((a, b) in have_fwd and (b, a) in have_fwd and (a, b) in have_cor):
                todo.append((img1, img2))
                have_fwd |= {(a, b), (b, a)}
                have_cor.add((a, b))
        _compute_pairs(memo, model, todo, device, desc_conf, subsample)
        if PERSIST_CACHE and cache_path:
            for img1, img2 in todo:
                a, b = img1["instance"], img2["instance"]
                _disk_put(_cache_file(cache_path, "fwd", a, b), memo["fwd"][a, b])
                _disk_put(_cache_file(cache_path, "fwd", b, a), memo["fwd"][b, a])
                _disk_put(_cache_file(cache_path, "corres", a, b, **fkw), memo["corres"][a, b])
    for img1, img2 in pairs:
        a, b = img1["instance"], img2["instance"]
        mirror(a, b)
        if _pair_missing(memo, a, b):
            continue                                  # (model is None: nothing to compute with)
        res_paths[a, b] = ((a, b), (b, a)), (a, b)
    return res_paths, cache_path


def canonical_view(ptmaps11, confs11, subsample, mode="avg-angle"):
    """sparse_ga.py:817-855 on the device (st3r_canonical_view)."""
    if mode != "avg-angle":
        raise NotImplementedError("only mode='avg-angle' is used by Starst3r (reconstruct.py:99)")
    assert len(ptmaps11) == len(confs11) > 0, "not a single view1 for this image"
    lib = _lib.load()
    P, H, W, _ = ptmaps11.shape
    dev = ptmaps11.device
    canon = torch.empty(H, W, 3, device=dev)
    canon2 = torch.empty(H, W, device=dev)
    cconf = torch.empty(H, W, device=dev)
    pt = ptmaps11.float().contiguous()
    cf = confs11.float().contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.st3r_canonical_view(_lib.ptr(pt), _lib.ptr(cf), P, H, W, subsample, _lib.ptr(canon),
                                           _lib.ptr(canon2), _lib.ptr(cconf), _lib.stream_ptr()), "st3r_canonical_view")
    return canon, canon2, cconf


def estimate_focal_knowing_depth(pts3d, pp=None, focal_mode="weiszfeld", min_focal=0., max_focal=np.inf):
    """dust3r/post_process.py:12-60, 'weiszfeld' with the principal point at the image centre (sparse_ga.py:696-698)."""
    if focal_mode != "weiszfeld":
        raise NotImplementedError("only focal_mode='weiszfeld' is on the Starst3r path")
    lib = _lib.load()
    B, H, W, _ = pts3d.shape
    out = torch.empty(B, device=pts3d.device)
    x = pts3d.float().contiguous()
    _lib.check(lib.st3r_align_set_variant(int(ALIGN_VARIANT)), "st3r_align_set_variant")
    with torch.cuda.device(x.device):
        _lib.check(lib.st3r_focal_weiszfeld(_lib.ptr(x), B, H, W, ctypes.c_float(min_focal),
                                            ctypes.c_float(min(max_focal, 3.0e38)), _lib.ptr(out), _lib.stream_ptr()),
                   "st3r_focal_weiszfeld")
    return out


def anchor_depth_offsets(canon_depth, pixels, subsample=8):
    """sparse_ga.py:858-886: block-quantised anchor index and depth ratio to the block's core pixel."""
    H1, W1 = canon_depth.shape
    W2 = len(range(subsample // 2, W1, subsample))
    core_idxs, core_offs = {}, {}
    for img2, (xy1, _confs) in pixels.items():
        px, py = xy1.long().T
        cy = (py // subsample) * subsample + subsample // 2
        cx = (px // subsample) * subsample + subsample // 2
        core_idxs[img2] = (py // subsample) * W2 + (px // subsample)
        core_offs[img2] = (canon_depth[py, px] / canon_depth[cy, cx]).detach()
    return core_idxs, core_offs


@torch.no_grad()
def prepare_canonical_data(imgs, tmp_pairs, subsample, order_imgs=False, min_conf_thr=0, cache_path=None,
                           device="cuda", **kw):
    """sparse_ga.py:633-714."""
    memo = _memo(cache_path)
    canonical_views, canonical_paths, preds_21 = {}, [], {}
    pairwise_scores = torch.zeros((len(imgs), len(imgs)), device=device)
    for img in imgs:
        canonical_paths.append((cache_path, img))
        cached = memo["canon"].get(img)
        if cached is None and PERSIST_CACHE and cache_path:
            got = _disk_get(_cache_file(cache_path, "canon", img, subsample=subsample, kw=kw), device)
            if got is not None:
                cached = memo["canon"][img] = (tuple(got[0]), got[1])
        pts, cfs, pixels = [], [], {}
        for (img1, img2), ((key1, key2), key_corres) in tmp_pairs.items():
            score = None
            if img == img1:
                X, C, X2, C2 = memo["fwd"][key1]
                score, (xy1, xy2, confs) = memo["corres"][key_corres]
                pixels[img2] = xy1, confs
                preds_21.setdefault(img, {})[img2] = (_subsampled(X2, subsample).reshape(-1, 3),
                                                      _subsampled(C2, subsample).ravel())
            if img == img2:
                X, C, X2, C2 = memo["fwd"][key2]
                score, (xy1, xy2, confs) = memo["corres"][key_corres]
                pixels[img1] = xy2, confs
                preds_21.setdefault(img, {})[img1] = (_subsampled(X2, subsample).reshape(-1, 3),
                                                      _subsampled(C2, subsample).ravel())
            if score is not None:
                i, j = imgs.index(img1), imgs.index(img2)
                pairwise_scores[i, j] = pairwise_scores[j, i] = score[2]
                if cached is None:
                    pts.append(X)
                    cfs.append(C)
        def compute_canon():
            if any(p is None for p in pts):
                raise RuntimeError(f"the point maps of {img} are not resident on this rank (forward_mast3r distributed "
                                   "them by image ownership); compute its canonical view on the owning rank")
            canon, canon2, cconf = canonical_view(torch.stack(pts), torch.stack(cfs), subsample, **kw)
            focal = estimate_focal_knowing_depth(canon[None], None, "weiszfeld", min_focal=0.5, max_focal=3.5)
            memo["canon"][img] = ((canon, canon2, cconf), focal)      # stays as computed on later add_images calls,
            _disk_put(cache_path and _cache_file(cache_path, "canon", img, subsample=subsample, kw=kw), memo["canon"][img])
        group = _shard_group()
        if group is None:
            if cached is None:
                compute_canon()
        else:
            # canonical views shard by image: image i belongs to rank i mod G, which streams its 2 (N - 1)
            # pair entries and broadcasts the 5 floats per pixel + the focal; the vote keeps the collectives aligned
            # when the ranks' memos differ
            rank, world = group
            import torch.distributed as dist
            need = torch.tensor([1 if cached is None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(need, op=dist.ReduceOp.MAX)
            if int(need.item()):
                owner = imgs.index(img) % world
                hw = torch.zeros(2, dtype=torch.int64, device=device)
                if rank == owner:
                    if cached is None:
                        compute_canon()
                    (canon, canon2, cconf), focal = memo["canon"][img]
                    hw = torch.tensor(canon2.shape, dtype=torch.int64, device=device)
                    flat = torch.cat([canon.reshape(-1), canon2.reshape(-1), cconf.reshape(-1), focal.reshape(-1)[:1]]).float().contiguous()
                dist.broadcast(hw, owner)
                Hc, Wc = (int(v) for v in hw.tolist())
                if rank != owner:
                    flat = torch.empty(5 * Hc * Wc + 1, dtype=torch.float32, device=device)
                dist.broadcast(flat, owner)
                if rank != owner:
                    n = Hc * Wc
                    memo["canon"][img] = ((flat[:3 * n].view(Hc, Wc, 3), flat[3 * n:4 * n].view(Hc, Wc),
                                           flat[4 * n:5 * n].view(Hc, Wc)), flat[5 * n:5 * n + 1].clone())
        (canon, canon2, cconf), focal = memo["canon"][img]            # like the reference's file cache (quirk C-6)
        H, W = canon.shape[:2]
        pp = torch.tensor([W / 2, H / 2], device=device)
        core_depth = canon[subsample // 2::subsample, subsample // 2::subsample, 2]
        idxs, offsets = anchor_depth_offsets(canon2, pixels, subsample=subsample)
        canonical_views[img] = (pp, (H, W), focal.view(1), core_depth, pixels, idxs, offsets)
    return tmp_pairs, pairwise_scores, canonical_views, canonical_paths, preds_21


def condense_data(imgs, tmp_paths, canonical_views, preds_21, dtype=torch.float32):
    """sparse_ga.py:729-814: per-image anchors (pixel, core index, depth offset) + per-pair slices into them."""
    set_imgs = set(imgs)
    pps, shapes, focals, core_depth, img_anchors, tmp_pixels = [], [], [], [], {}, {}
    for idx1, img1 in enumerate(imgs):
        pp, shape, focal, anchors, pixels_confs, idxs, offsets = canonical_views[img1]
        pps.append(pp); shapes.append(shape); focals.append(focal); core_depth.append(anchors)
        uv, ii, oo, cur = [], [], [], 0
        for img2, (pixels, match_confs) in pixels_confs.items():
            if img2 not in set_imgs:
                continue
            assert len(pixels) == len(idxs[img2]) == len(offsets[img2])
            uv.append(torch.cat((pixels, torch.ones_like(pixels[:, :1])), dim=-1))
            ii.append(idxs[img2]); oo.append(offsets[img2])
            tmp_pixels[img1, img2] = pixels.to(dtype), match_confs.to(dtype), slice(cur, cur + len(pixels))
            cur += len(pixels)
        img_anchors[idx1] = (torch.cat(uv), torch.cat(ii), torch.cat(oo))
    all_confs, imgs_slices = [], []
    corres2d = {i: [] for i in range(len(imgs))}
    for img1, img2 in tmp_paths:
        if (img1, img2) not in tmp_pixels or (img2, img1) not in tmp_pixels:
            continue
        pix1, confs1, slice1 = tmp_pixels[img1, img2]
        pix2, confs2, slice2 = tmp_pixels[img2, img1]
        i1, i2 = imgs.index(img1), imgs.index(img2)
        confs = (confs1 * confs2).sqrt()
        all_confs.append(confs)
        imgs_slices.append(PairOfSlices(i1, slice1, pix1, canonical_views[img1][5][img2], i2, slice2, pix2,
                                        canonical_views[img2][5][img1], confs, float(confs.sum())))
        corres2d[i1].append((pix1, confs, i2, slice2))
        corres2d[i2].append((pix2, confs, i1, slice1))
    all_confs = torch.cat(all_confs)
    corres = (all_confs, float(all_confs.sum()), imgs_slices)

    def aggreg(i, ms):
        pix, confs, j, sl = zip(*ms)
        c = torch.cat(confs).to(dtype)
        return i, torch.cat(pix).to(dtype), c, float(c.sum()), list(zip(j, sl))
    corres2d = [aggreg(i, m) for i, m in corres2d.items()]
    imsizes = torch.tensor([(W, H) for H, W in shapes], device=pps[0].device)
    sub = {}
    for imk in preds_21:
        sub[imk] = {}
        for im2k, (pred, conf) in preds_21[imk].items():
            ia = img_anchors[imgs.index(im2k)][1]
            sub[imk][im2k] = (pred[ia], conf[ia])
    return imsizes, torch.stack(pps), torch.cat(focals), core_depth, img_anchors, corres, corres2d, sub


def compute_min_spanning_tree(pws):
    """sparse_ga.py:984-1009 (host, scipy): MST on -score, rooted at the node farthest from any leaf, BFS edge order."""
    from scipy import sparse as sp
    n = pws.shape[0]
    g = sp.dok_array((n, n))
    for i, j in pws.nonzero().cpu().tolist():
        g[i, j] = -float(pws[i, j])
    msp = sp.csgraph.minimum_spanning_tree(g)

    def bfs_ranks(start):
        order, _ = sp.csgraph.breadth_first_order(msp, start, directed=False)
        ranks = np.arange(len(order))
        ranks[order] = ranks.copy()
        return ranks
    r1 = bfs_ranks(0)
    r2 = bfs_ranks(r1.argmax())
    r1 = bfs_ranks(r2.argmax())
    root = int(np.minimum(r1, r2).argmax())
    order, pred = sp.csgraph.breadth_first_order(msp, root, directed=False)
    return root, [(int(pred[i]), int(i)) for i in order[1:]]


class SparseGA:
    """Result object with the surface of sparse_ga.py:33-108."""

    def __init__(self, img_paths, pairs_in, res_fine, anchors, canonical_paths=None, subsample=8):
        def fetch(name):
            for im1, im2 in pairs_in:
                for im in (im1, im2):
                    if im["instance"] == name:
                        return (im["img"][0].permute(1, 2, 0).cpu().numpy() * .5 + .5).clip(min=0., max=1.)
        self.canonical_paths, self.img_paths = canonical_paths, img_paths
        self.imgs = [fetch(n) for n in img_paths]
        self.intrinsics, self.cam2w = res_fine["intrinsics"], res_fine["cam2w"]
        self.depthmaps, self.pts3d = res_fine["depthmaps"], res_fine["pts3d"]
        self.working_device = self.cam2w.device
        self.pts3d_colors = []
        for i, im in enumerate(self.imgs):
            x, y = anchors[i][0][..., :2].detach().cpu().long().numpy().T
            self.pts3d_colors.append(im[y, x])
            assert self.pts3d_colors[-1].shape == tuple(self.pts3d[i].shape)
        self.n_imgs = len(self.imgs)
        self.subsample = subsample

    def get_focals(self):
        return self.intrinsics[:, 0, 0].clone()

    def get_principal_points(self):
        return self.intrinsics[:, :2, 2].clone()

    def get_im_poses(self):
        return self.cam2w

    def get_sparse_pts3d(self):
        return self.pts3d

    def get_pts3d_colors(self):
        return self.pts3d_colors

    def get_depthmaps(self):
        return self.depthmaps

    def get_masks(self):
        return [slice(None, None) for _ in range(len(self.imgs))]

    def get_dense_pts3d(self, clean_depth=True, subsample=8):
        """sparse_ga.py:70-93 + clean_pointcloud (dust3r/cloud_opt/base_opt.py:369-405) on the device."""
        assert self.canonical_paths, "cache_path is required for dense 3d points"
        lib = _lib.load()
        dev = self.cam2w.device
        pts3d, depths, confs = [], [], []
        cam_h = self.cam2w.detach().float().cpu().contiguous()
        K_h = self.intrinsics.detach().float().cpu().contiguous()
        for i, (cache_path, img) in enumerate(self.canonical_paths):
            (canon, canon2, conf), focal = _memo(cache_path)["canon"][img]
            H, W = conf.shape
            p = torch.empty(H * W, 3, device=dev)
            d = torch.empty(H * W, device=dev)
            core = self.depthmaps[i].reshape(-1).float().contiguous()
            c2 = canon2.contiguous()
            with torch.cuda.device(dev):
                _lib.check(lib.st3r_dense_points(_lib.ptr(c2), _lib.ptr(core), cam_h[i].numpy().ctypes.data_as(ctypes.c_void_p),
                                                 K_h[i].numpy().ctypes.data_as(ctypes.c_void_p), ctypes.c_float(float(focal)),
                                                 H, W, subsample, _lib.ptr(p), _lib.ptr(d), _lib.stream_ptr()),
                           "st3r_dense_points")
            pts3d.append(p); depths.append(d); confs.append(conf.clone())
        if clean_depth:
            # N tiny 4x4 inverses: on the host (the first cuSOLVER call on a device costs ~0.3 s of initialisation)
            w2c = torch.linalg.inv(self.cam2w.detach().float().cpu()).to(dev)
            confs = clean_pointcloud(confs, self.intrinsics, w2c, depths, pts3d)
        return pts3d, depths, confs


def clean_pointcloud(im_confs, K, cams, depthmaps, all_pts3d, tol=0.001, bad_conf=0, dbg=()):
    """dust3r/cloud_opt/base_opt.py:369-405 (all views must share one size, which Starst3r guarantees per scene)."""
    assert len(im_confs) == len(cams) == len(K) == len(depthmaps) == len(all_pts3d)
    assert 0 <= tol < 1
    lib = _lib.load()
    N = len(im_confs)
    H, W = im_confs[0].shape
    if any(c.shape != (H, W) for c in im_confs):
        raise NotImplementedError("clean_pointcloud on B200 expects equally sized views")
    dev = im_confs[0].device
    conf = torch.stack([c.reshape(-1) for c in im_confs]).float().contiguous()
    depth = torch.stack([d.reshape(-1) for d in depthmaps]).float().contiguous()
    pts = torch.stack([p.reshape(-1, 3) for p in all_pts3d]).float().contiguous()
    cm = torch.cat([cams[:, :3, :3].reshape(N, 9), cams[:, :3, 3], K.reshape(N, 9)], dim=1).float().contiguous()
    assert cm.shape[1] == lib.st3r_clean_cam_floats()
    with torch.cuda.device(dev):
        _lib.check(lib.st3r_clean_pointcloud(_lib.ptr(pts), _lib.ptr(conf), _lib.ptr(depth), _lib.ptr(cm), N, H, W,
                                             ctypes.c_float(tol), ctypes.c_float(bad_conf), _lib.stream_ptr()),
                   "st3r_clean_pointcloud")
    return [conf[i].reshape(H, W) for i in range(N)]


def run_sparse_ga(imgs, pairs_in, cache_path, model, subsample=8, desc_conf="desc_conf", device="cuda",
                  dtype=torch.float32, shared_intrinsics=False, optim_params=None, **kw):
    """starster/reconstruct.py:75-113."""
    pairs_in = convert_dust3r_pairs_naming(imgs, pairs_in)
    pairs, cache_path = forward_mast3r(pairs_in, model, cache_path=cache_path, subsample=subsample,
                                       desc_conf=desc_conf, device=device)
    tmp_pairs, pairwise_scores, canonical_views, canonical_paths, preds_21 = prepare_canonical_data(
        imgs, pairs, subsample, cache_path=cache_path, mode="avg-angle", device=device)
    mst = compute_min_spanning_tree(pairwise_scores)
    imsizes, pps, base_focals, core_depth, anchors, corres, corres2d, preds_21 = condense_data(
        imgs, tmp_pairs, canonical_views, preds_21, dtype)
    imgs, res_coarse, res_fine, optim_params = sparse_scene_optimizer_slam(
        imgs, subsample, imsizes, pps, base_focals, core_depth, anchors, corres, corres2d, preds_21, canonical_paths,
        mst, shared_intrinsics=shared_intrinsics, cache_path=cache_path, device=device, dtype=dtype,
        prev_params=optim_params, **kw)
    return SparseGA(imgs, pairs_in, res_fine or res_coarse, anchors, canonical_paths, subsample), optim_params


from .image import prepare_images_for_mast3r  # noqa: E402  (starster/image.py:112-139)


def reconstruct_scene(model, imgs, filelist, device, optim_params=None, tmpdir=None):
    """starster/reconstruct.py:19-72: MASt3R inference (out of scope, any `model`) + matching + global alignment."""
    import tempfile
    imgs = prepare_images_for_mast3r(imgs)
    pairs = make_pairs(imgs, scene_graph="complete", prefilter=None, symmetrize=True)
    private = tmpdir is None
    if private:
        tmpdir = tempfile.mkdtemp()
    scene, optim_params = run_sparse_ga(filelist, pairs, tmpdir, model, lr1=0.07, niter1=500, lr2=0.014, niter2=200,
                                        device=device, opt_depth=False, matching_conf_thr=5, shared_intrinsics=False,
                                        optim_params=optim_params)
    if private:     # nobody else can name this cache: its device-resident memo goes when the result object goes
        import weakref
        weakref.finalize(scene, clear_cache, tmpdir)
    return scene, optim_params
