"""MATCH hot path — host-side mirror of the reference operator interface.

Same names / argument meaning / error behaviour as
  mast3r/mast3r/fast_nn.py            (bruteforce_reciprocal_nns, cdistMatcher, merge_corres,
                                        fast_reciprocal_NNs)
  mast3r/mast3r/cloud_opt/sparse_ga.py:595-630 (extract_correspondences)
but every call runs hand-written sm_100a kernels through the C ABI of
libstarst3r_b200.so (include/starst3r_b200.h).  There is no CPU fallback.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

__all__ = ("nn_argmax", "bruteforce_reciprocal_nns", "cdistMatcher", "merge_corres", "fast_reciprocal_NNs",
           "extract_correspondences")

_IMPL = {"auto": 0, "simt": 1, "tcgen05": 2}


def _dev(device):
    return _lib.require_cuda_device(device, "starst3r_b200.match")


def _f32(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=torch.float32).contiguous()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def nn_argmax(Q, DB, impl="auto", return_score=False):
    """idx[i] = argmax_j <Q[i], DB[j]> (fp32 FMA chain, ties -> lowest j).  Device tensors in/out (int32)."""
    lib = _lib.load()
    _lib.require_cuda(Q, DB)
    Q = Q.float().contiguous()
    DB = DB.float().contiguous()
    M, d = Q.shape
    N = DB.shape[0]
    assert DB.shape[1] == d
    idx = torch.empty(M, dtype=torch.int32, device=Q.device)
    best = torch.empty(M, dtype=torch.float32, device=Q.device) if return_score else None
    if M == 0:
        return (idx, best) if return_score else idx
    if N == 0:
        raise ValueError("nn_argmax: empty database")
    _apply_variant(lib)        # the workspace size depends on the split-precision switch
    ws = _ws(lib.st3r_nn_argmax_ws_bytes(M, N, d), Q.device)
    with torch.cuda.device(Q.device):
        rc = lib.st3r_nn_argmax(_lib.ptr(Q), M, _lib.ptr(DB), N, d, _lib.ptr(idx), _lib.ptr(best), _lib.ptr(ws),
                                ws.numel(), _IMPL[impl], _lib.stream_ptr())
    _lib.check(rc, "st3r_nn_argmax")
    return (idx, best) if return_score else idx


@torch.no_grad()
def bruteforce_reciprocal_nns(A, B, device="cuda", block_size=None, dist="l2", impl="auto"):
    """fast_nn.py:16-70.  `block_size` is accepted for signature parity; the kernels never materialise the
    score matrix, so it has no effect on memory or on the result (ties keep the lowest index either way)."""
    if dist == "l2":
        raise NotImplementedError("dist='l2' is outside the B200 hot path (Starst3r only matches with dist='dot', "
                                  "sparse_ga.py:604)")
    if dist != "dot":
        raise ValueError(f"Unknown {dist=}")
    device = _dev(device)
    A = _f32(A, device)
    B = _f32(B, device)
    nn_A = nn_argmax(A, B, impl=impl)
    nn_B = nn_argmax(B, A, impl=impl)
    return nn_A.cpu().numpy().astype(np.int64), nn_B.cpu().numpy().astype(np.int64)


class cdistMatcher:
    """fast_nn.py:73-84."""

    def __init__(self, db_pts, device="cuda"):
        self.device = _dev(device)
        self.db_pts = _f32(db_pts, self.device)

    def query(self, queries, k=1, dist="dot", block_size=None, impl="auto", **kw):
        assert k == 1
        if queries.numel() == 0:
            return None, []
        if dist != "dot":
            raise NotImplementedError("only dist='dot' is on the B200 hot path")
        nnA = nn_argmax(_f32(queries, self.device), self.db_pts, impl=impl)
        return None, nnA.cpu().numpy().astype(np.int64)


def merge_corres(idx1, idx2, shape1=None, shape2=None, ret_xy=True, ret_index=False, device="cuda"):
    """fast_nn.py:87-106 on the device: radix sort of packed (idx1, idx2) + ordered unique."""
    assert idx1.dtype == idx2.dtype == np.int32
    lib = _lib.load()
    device = _dev(device)
    n = len(idx1)
    hw1 = int(shape1[0] * shape1[1]) if shape1 else (int(idx1.max()) + 1 if n else 1)
    hw2 = int(shape2[0] * shape2[1]) if shape2 else (int(idx2.max()) + 1 if n else 1)
    d1 = torch.from_numpy(np.ascontiguousarray(idx1)).to(device)
    d2 = torch.from_numpy(np.ascontiguousarray(idx2)).to(device)
    o1 = torch.empty(max(n, 1), dtype=torch.int32, device=device)
    o2 = torch.empty_like(o1)
    oi = torch.empty_like(o1)
    n_out = torch.zeros(1, dtype=torch.int32, device=device)
    ws = _ws(lib.st3r_merge_corres_ws_bytes(n), device)
    with torch.cuda.device(device):
        rc = lib.st3r_merge_corres(_lib.ptr(d1), _lib.ptr(d2), n, hw1, hw2, _lib.ptr(o1), _lib.ptr(o2), _lib.ptr(oi),
                                   _lib.ptr(n_out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "st3r_merge_corres")
    k = int(n_out.item())
    xy1 = o1[:k].cpu().numpy()
    xy2 = o2[:k].cpu().numpy()
    indices = oi[:k].cpu().numpy().astype(np.int64)
    if ret_xy:
        assert shape1 and shape2
        xy1 = _unravel(xy1, shape1, ret_xy)
        xy2 = _unravel(xy2, shape2, ret_xy)
    if ret_index:
        return xy1, xy2, indices
    return xy1, xy2


def _unravel(idx, shape, ret_xy=True):
    y, x = np.unravel_index(idx, shape)
    if ret_xy == "y_x":
        return (y, x)
    return np.stack([x, y], axis=1).astype(np.int64)


def _recip_device(P1, P2, subsample, seeds, max_iter, impl):
    """One seeded reciprocal search on the device; returns int32 device tensors (idx1, idx2) (sorted unique)."""
    lib = _lib.load()
    H1, W1, d = P1.shape
    H2, W2, _ = P2.shape
    dev = P1.device
    if seeds is None:
        nseed = lib.st3r_recip_seed_count(H1, W1, subsample)
        seeds_t = None
    else:
        seeds_t = torch.from_numpy(np.ascontiguousarray(seeds, dtype=np.int32)).to(dev)
        nseed = seeds_t.numel()
    cap = max(nseed, 1)
    o1 = torch.empty(cap, dtype=torch.int32, device=dev)
    o2 = torch.empty(cap, dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = _ws(lib.st3r_recip_nn_ws_bytes(nseed, nseed, max_iter), dev)
    with torch.cuda.device(dev):
        rc = lib.st3r_recip_nn(_lib.ptr(P1), H1, W1, _lib.ptr(P2), H2, W2, d, int(subsample) if seeds is None else 1,
                               _lib.ptr(seeds_t), nseed if seeds is not None else 0, max_iter, _lib.ptr(o1),
                               _lib.ptr(o2), _lib.ptr(n_out), _lib.ptr(ws), ws.numel(), _IMPL[impl],
                               _lib.stream_ptr())
    _lib.check(rc, "st3r_recip_nn")
    k = int(n_out.item())
    return o1[:k], o2[:k]


def fast_reciprocal_NNs(pts1, pts2, subsample_or_initxy1=8, ret_xy=True, pixel_tol=0, ret_basin=False,
                        device="cuda", impl="auto", **matcher_kw):
    """fast_nn.py:109-188.  The default call (integer subsample, pixel_tol=0, ret_basin=False — the only form
    Starst3r uses, sparse_ga.py:614-615) runs as one device-resident launch chain; the other forms drive the
    same NN kernel from a host loop that follows the reference statement by statement."""
    H1, W1, DIM1 = pts1.shape
    H2, W2, DIM2 = pts2.shape
    assert DIM1 == DIM2
    dist = matcher_kw.get("dist", "dot")
    if dist != "dot":
        raise NotImplementedError("only dist='dot' is on the B200 hot path (KDTree / l2 branches are CPU code "
                                  "in the reference)")
    device = _dev(device)
    P1 = _f32(pts1, device)
    P2 = _f32(pts2, device)

    if isinstance(subsample_or_initxy1, int) and pixel_tol == 0 and not ret_basin:
        o1, o2 = _recip_device(P1, P2, subsample_or_initxy1, None, 10, impl)
        xy1, xy2 = o1.cpu().numpy(), o2.cpu().numpy()
        if ret_xy:
            xy1, xy2 = _unravel(xy1, (H1, W1), ret_xy), _unravel(xy2, (H2, W2), ret_xy)
        return xy1, xy2

    # ---- general form (pixel_tol > 0, explicit seeds, ret_basin) -------------------------------------
    F1 = P1.reshape(-1, DIM1)
    F2 = P2.reshape(-1, DIM2)
    if isinstance(subsample_or_initxy1, int) and pixel_tol == 0:
        S = subsample_or_initxy1
        y1, x1 = np.mgrid[S // 2:H1:S, S // 2:W1:S].reshape(2, -1)
        max_iter = 10
    else:
        x1, y1 = subsample_or_initxy1
        x1 = x1.cpu().numpy() if isinstance(x1, torch.Tensor) else np.asarray(x1)
        y1 = y1.cpu().numpy() if isinstance(y1, torch.Tensor) else np.asarray(y1)
        max_iter = 1
    xy1 = np.int32(np.unique(x1 + W1 * y1))
    xy2 = np.full_like(xy1, -1)
    old_xy1, old_xy2 = xy1.copy(), xy2.copy()
    notyet = np.ones(len(xy1), dtype=bool)
    basin = np.full((H1 * W1 + 1,), -1, dtype=np.int32) if ret_basin else None

    def query(F_src, sel, F_db):
        if len(sel) == 0:
            return np.zeros(0, np.int32)
        q = F_src[torch.from_numpy(sel.astype(np.int64)).to(device)]
        return nn_argmax(q, F_db, impl=impl).cpu().numpy()

    niter = 0
    while notyet.any():
        xy2[notyet] = query(F1, xy1[notyet], F2)
        if not ret_basin:
            notyet &= (old_xy2 != xy2)
        xy1[notyet] = query(F2, xy2[notyet], F1)
        if ret_basin:
            basin[old_xy1[notyet]] = xy1[notyet]
        notyet &= (old_xy1 != xy1)
        niter += 1
        if niter >= max_iter:
            break
        old_xy2[:] = xy2
        old_xy1[:] = xy1

    if pixel_tol > 0:
        old_yx1 = np.stack(np.unravel_index(old_xy1, (H1, W1)), axis=1)
        new_yx1 = np.stack(np.unravel_index(xy1, (H1, W1)), axis=1)
        dis = np.linalg.norm(old_yx1 - new_yx1, axis=-1)
        converged = dis < pixel_tol
        if not isinstance(subsample_or_initxy1, int):
            xy1 = old_xy1
    else:
        converged = ~notyet
    xy1, xy2 = merge_corres(xy1[converged], xy2[converged], (H1, W1), (H2, W2), ret_xy=ret_xy, device=device)
    if ret_basin:
        return xy1, xy2, basin
    return xy1, xy2


def recip_query_rows(pts1, pts2, subsample=8, max_iter=10, impl="auto"):
    """Diagnostic: the number of query rows each NN call of fast_reciprocal_NNs issues on this input (the sizes
    of the not-yet-converged set, fast_nn.py:152-168), measured by driving the CUDA NN kernel from the host loop.
    bench.py uses the sum as the algorithmic FLOP count (2 * rows * H*W * d)."""
    device = pts1.device
    H1, W1, D = pts1.shape
    F1, F2 = pts1.reshape(-1, D).float().contiguous(), pts2.reshape(-1, D).float().contiguous()
    S = subsample
    y1, x1 = np.mgrid[S // 2:H1:S, S // 2:W1:S].reshape(2, -1)
    xy1 = torch.from_numpy(np.unique(x1 + W1 * y1).astype(np.int64)).to(device)
    xy2 = torch.full_like(xy1, -1)
    old1, old2 = xy1.clone(), xy2.clone()
    notyet = torch.ones_like(xy1, dtype=torch.bool)
    rows = []
    for _ in range(max_iter):
        if not bool(notyet.any()):
            break
        rows.append(int(notyet.sum()))
        xy2[notyet] = nn_argmax(F1[xy1[notyet]], F2, impl=impl).long()
        notyet &= old2 != xy2
        rows.append(int(notyet.sum()))
        if bool(notyet.any()):
            xy1[notyet] = nn_argmax(F2[xy2[notyet]], F1, impl=impl).long()
        notyet &= old1 != xy1
        old2.copy_(xy2)
        old1.copy_(xy1)
    return rows


class _ExtractPlan:
    """Persistent buffers + a captured CUDA graph of one st3r_extract_corres launch chain (~60 small launches: the four
    searches advance in lock-step, 20 half-iterations x {batched NN, update}, + sort + unique).  The chain is launch-latency
    bound (measured: ~50 % of the pair time is gaps between tiny tail kernels), and its shape only depends on the
    map sizes, so it is captured once per (H1, W1, H2, W2, d, subsample, impl) and replayed; inputs are copied into
    the plan's staging buffers (8 device-to-device copies, ~100 MB at 512x512, ~35 us)."""

    def __init__(self, lib, dev, H1, W1, H2, W2, d, subsample, impl, max_iter):
        self.key = (H1, W1, H2, W2, d, subsample, impl, max_iter)
        f = lambda h, w: torch.empty(h, w, d, dtype=torch.float32, device=dev)   # noqa: E731
        q = lambda h, w: torch.empty(h, w, dtype=torch.float32, device=dev)      # noqa: E731
        self.feats = [f(H1, W1), f(H2, W2), f(H2, W2), f(H1, W1)]                # feat11, feat21, feat22, feat12
        self.qonfs = [q(H1, W1), q(H2, W2), q(H2, W2), q(H1, W1)]
        cap = max(lib.st3r_extract_corres_cap(H1, W1, H2, W2, subsample), 1)
        self.xy1 = torch.empty((cap, 2), dtype=torch.int64, device=dev)
        self.xy2 = torch.empty((cap, 2), dtype=torch.int64, device=dev)
        self.conf = torch.empty(cap, dtype=torch.float32, device=dev)
        self.n_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ws_by_split = {}       # split-precision switch -> workspace (the split variant needs tf32 head / tail maps)
        self.ws = None
        self.graphs = {}            # matcher variant (cooperative rare path, split precision) -> captured graph
        self.lib, self.dev = lib, dev

    def select(self, variant):
        """Workspace for the matcher variant `variant` = (cooperative, split) that _apply_variant just switched on."""
        split = variant[1]
        if split not in self.ws_by_split:
            H1, W1, H2, W2, d, subsample, impl, max_iter = self.key
            self.ws_by_split[split] = _ws(self.lib.st3r_extract_corres_ws_bytes(H1, W1, H2, W2, subsample, max_iter),
                                          self.dev)
        self.ws = self.ws_by_split[split]

    def _launch(self):
        H1, W1, H2, W2, d, subsample, impl, max_iter = self.key
        f, q = self.feats, self.qonfs
        rc = self.lib.st3r_extract_corres(_lib.ptr(f[0]), _lib.ptr(f[1]), _lib.ptr(f[2]), _lib.ptr(f[3]), _lib.ptr(q[0]),
                                          _lib.ptr(q[1]), _lib.ptr(q[2]), _lib.ptr(q[3]), H1, W1, H2, W2, d, subsample,
                                          max_iter, _lib.ptr(self.xy1), _lib.ptr(self.xy2), _lib.ptr(self.conf),
                                          _lib.ptr(self.n_out), _lib.ptr(self.ws), self.ws.numel(), _IMPL[impl],
                                          _lib.stream_ptr())
        _lib.check(rc, "st3r_extract_corres")

    def run(self, feats, qonfs):
        with torch.cuda.device(self.dev):
            for dst, src in zip(self.feats + self.qonfs, list(feats) + list(qonfs)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            coop = _apply_variant(self.lib)
            self.select(coop)
            if coop not in self.graphs:
                self._launch()                      # eager warm-up (sets kernel attributes, loads modules)
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch()
                self.graphs[coop] = g
            self.graphs[coop].replay()
        return self.xy1, self.xy2, self.conf, self.n_out


_PLANS = {}
USE_CUDA_GRAPHS = True

# The tcgen05 matcher scores with TF32 tensor-core products and re-scores, exactly, every column inside the error band
# of the running maximum.  How many columns that is depends on the data, and two switches adapt the kernel to it - the
# results are identical in every combination (the exact fp32 re-score decides):
#   * st3r_nn_tc_set_split: split precision (hi.hi + hi.lo + lo.hi, 3 x the tensor work, ~40x narrower band).  Measured
#     on the B200 (profiles/r02n_nn_margin.json, r02a_variants.json): smooth descriptor fields - real MASt3R maps, the
#     synthetic scene - 8.0 -> 4.5 ms per 512 x 512 pair; random descriptors 2.5 -> 3.3 ms.  The band keeps a 256x
#     margin over the largest error the tensor core ever showed and 4x over the worst-case bound.
#   * st3r_nn_tc_set_cooperative: warp-cooperative resolution of full candidate lists (2x on smooth fields when the
#     split is off, ~14 % slower on random data).
# "auto" follows the statistics the kernels collect (exact list resolutions per scanned query row): random-like
# descriptors resolve ~never and run the plain per-thread variant; fields that resolve more than once per two rows
# switch to split precision (and back when its own, ~40x smaller, ratio says the data changed).  True / False pin a
# switch (ST3R_NN_SPLIT=0/1 in the environment pins the split).
NN_COOPERATIVE = "auto"
NN_SPLIT = {"0": False, "1": True}.get(os.environ.get("ST3R_NN_SPLIT", ""), "auto")
_variant = {"on": False, "split": False, "probe_in": 0}
# Resolutions per query row of the PLAIN kernel: above -> split on, below -> split off.  Measured (B200,
# profiles/r02o_nn_margin.json): random descriptors 0.0, the synthetic scene 5.4, smooth fields 244; the split kernel itself
# resolves ~never on anything but the smoothest fields, so its own statistics cannot tell when to switch back: while
# the split is on, every SPLIT_PROBE_EVERY-th call runs the plain kernel once as a probe.
SPLIT_ON_RATIO, SPLIT_OFF_RATIO, SPLIT_PROBE_EVERY = 0.5, 0.1, 64


def _apply_variant(lib):
    """Selects the matcher variant for the next launches; returns the key (cooperative, split) it stands for."""
    split = _variant["split"] if NN_SPLIT == "auto" else bool(NN_SPLIT)
    if NN_SPLIT == "auto" and split:
        _variant["probe_in"] -= 1
        if _variant["probe_in"] <= 0:
            split = False                        # probe: one call with the plain kernel, _adapt_variant reads its ratio
    on = (_variant["on"] and not split) if NN_COOPERATIVE == "auto" else bool(NN_COOPERATIVE)
    lib.st3r_nn_tc_set_cooperative(int(on))
    lib.st3r_nn_tc_set_split(int(split))
    _variant["ran_split"] = split
    return (on, split)


def _adapt_variant(lib):
    """Called where the host has just synchronised anyway (the correspondence count was read back)."""
    if NN_COOPERATIVE != "auto" and NN_SPLIT != "auto":
        return
    st = (ctypes.c_ulonglong * 2)()
    _lib.check(lib.st3r_nn_tc_stats(st, 1), "st3r_nn_tc_stats")
    rows, resolves = int(st[0]), int(st[1])
    if rows < 256 or _variant.get("ran_split"):
        return                       # (the split kernel's own ratio says nothing about the plain band)
    ratio = resolves / rows          # exact list resolutions per query row of the plain kernel
    if NN_SPLIT == "auto":
        if ratio > SPLIT_ON_RATIO:
            _variant["split"] = True
            _variant["probe_in"] = SPLIT_PROBE_EVERY
        elif ratio < SPLIT_OFF_RATIO:
            _variant["split"] = False
        elif _variant["split"]:
            _variant["probe_in"] = SPLIT_PROBE_EVERY
    if NN_COOPERATIVE == "auto":
        if ratio > 0.5:
            _variant["on"] = True
        elif ratio < 0.1:
            _variant["on"] = False


def adapt_variant():
    """Lets the matcher re-pick its precision variant from the statistics of the calls since the last read-out.  For
    callers of extract_correspondences_device, at a point where they have synchronised anyway (it reads two device
    counters)."""
    _adapt_variant(_lib.load())


def extract_correspondences_device(feats, qonfs, subsample=8, impl="auto", max_iter=10):
    """Device-resident form: returns (xy1 [cap,2] i64, xy2 [cap,2] i64, conf [cap] f32, n [1] i32) without
    synchronising; rows >= n are undefined and the buffers are reused by the next call with the same shapes.
    Used by the pair-sharded pipeline and by bench.py."""
    lib = _lib.load()
    f11, f21, f22, f12 = feats
    q11, q21, q22, q12 = qonfs
    assert f11.shape[:2] == f12.shape[:2] == q11.shape == q12.shape
    assert f21.shape[:2] == f22.shape[:2] == q21.shape == q22.shape
    _lib.require_cuda(f11, f21, f22, f12, q11, q21, q22, q12)
    dev = f11.device
    H1, W1, d = f11.shape
    H2, W2, _ = f22.shape
    key = (dev.index, H1, W1, H2, W2, d, subsample, impl, max_iter)
    plan = _PLANS.get(key)
    if plan is None:
        plan = _PLANS[key] = _ExtractPlan(lib, dev, H1, W1, H2, W2, d, subsample, impl, max_iter)
    if not USE_CUDA_GRAPHS:
        for dst, src in zip(plan.feats + plan.qonfs, list(feats) + list(qonfs)):
            dst.copy_(src, non_blocking=True)
        with torch.cuda.device(dev):
            plan.select(_apply_variant(lib))
            plan._launch()
        return plan.xy1, plan.xy2, plan.conf, plan.n_out
    return plan.run([x.float() for x in feats], [x.float() for x in qonfs])


def extract_correspondences(feats, qonfs, subsample=8, device=None, ptmap_key="pred_desc", impl="auto"):
    """sparse_ga.py:595-630: (xy1, xy2, conf) torch tensors on `device` for one image pair."""
    if "3d" in ptmap_key:
        raise NotImplementedError("ptmap_key with '3d' selects the CPU KDTree matcher in the reference "
                                  "(sparse_ga.py:601-602); it is outside the B200 hot path")
    device = _dev(device if device is not None else feats[0].device)
    feats = [_f32(f, device) for f in feats]
    qonfs = [_f32(q, device) for q in qonfs]
    xy1, xy2, conf, n_out = extract_correspondences_device(feats, qonfs, subsample, impl)
    n = int(n_out.item())
    _adapt_variant(_lib.load())
    return xy1[:n].clone(), xy2[:n].clone(), conf[:n].clone()      # the plan's buffers are reused by the next pair
