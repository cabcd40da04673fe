"""Golden fixtures for the ALIGN path, produced by the UNMODIFIED reference (mast3r/cloud_opt/sparse_ga.py) in the
build container.  Only the network is replaced: `symmetric_inference` is monkey-patched with the scene-consistent
synthetic generator (starst3r_b200/synth.py), because the MASt3R checkpoint is not available offline and the network
is outside the hot path.  Everything downstream (forward_mast3r's pair loop + extract_correspondences,
prepare_canonical_data, canonical_view, anchor_depth_offsets, compute_min_spanning_tree, condense_data,
sparse_scene_optimizer, SparseGA.get_dense_pts3d, clean_pointcloud) is the reference's own code.
TEST INFRASTRUCTURE ONLY."""
import os
import tempfile

import numpy as np
import torch


def plain(x):
    """Reference structures -> plain python containers of CPU tensors (PairOfSlices -> tuple, slice -> (a,b))."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().clone()
    if isinstance(x, slice):
        return ("slice", x.start, x.stop)
    if isinstance(x, dict):
        return {k: plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return tuple(plain(v) for v in x)
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    return x


def run_reference(sparse_ga, n_views, W, H, niter1, niter2, seed=0, low_conf=False, lr1=0.07, lr2=0.014, pts_noise=0.0):
    from dust3r.image_pairs import make_pairs
    from starst3r_b200 import synth
    model = synth.SyntheticMast3r(n_views, W, H, seed=seed, low_conf=low_conf, pts_noise=pts_noise, arc_deg=90.0)
    sparse_ga.symmetric_inference = lambda model, img1, img2, device: model.symmetric_inference(img1, img2)
    raw = model.images()
    imgs = [dict(img=im[None], true_shape=np.int32([im.shape[-2:]]), idx=i, instance=str(i)) for i, im in enumerate(raw)]
    filelist = [f"{i}.png" for i in range(n_views)]
    pairs_in = make_pairs(imgs, scene_graph="complete", prefilter=None, symmetrize=True)
    pairs_in = sparse_ga.convert_dust3r_pairs_naming(filelist, pairs_in)
    cache = tempfile.mkdtemp()
    pairs, cache = sparse_ga.forward_mast3r(pairs_in, model, cache_path=cache, subsample=8, desc_conf="desc_conf",
                                            device="cpu")
    tmp_pairs, pairwise_scores, canonical_views, canonical_paths, preds_21 = sparse_ga.prepare_canonical_data(
        filelist, pairs, 8, cache_path=cache, mode="avg-angle", device="cpu")
    mst = sparse_ga.compute_min_spanning_tree(pairwise_scores)
    imsizes, pps, base_focals, core_depth, anchors, corres, corres2d, preds_21c = sparse_ga.condense_data(
        filelist, tmp_pairs, canonical_views, preds_21, torch.float32)
    inputs = plain(dict(imgs=filelist, imsizes=imsizes, pps=pps, base_focals=base_focals, core_depth=core_depth,
                        anchors=anchors, corres=corres, corres2d=corres2d, preds_21=preds_21c,
                        mst=(int(mst[0]), [(int(a), int(b)) for a, b in mst[1]])))
    out = {}
    for tag, (n1, n2) in {"init": (0, 0), "short": (niter1, niter2), "full": (300, 200)}.items():
        a = plain(dict(imsizes=imsizes, pps=pps, base_focals=base_focals, core_depth=core_depth))   # fresh copies:
        imsz, pp_, bf_, cd_ = a["imsizes"], a["pps"].clone(), a["base_focals"].clone(), [c.clone() for c in a["core_depth"]]
        _, res_c, res_f = sparse_ga.sparse_scene_optimizer(
            filelist, 8, imsz, pp_, bf_, cd_, anchors, corres, corres2d, preds_21c, canonical_paths, mst,
            cache_path=cache, lr1=lr1, niter1=n1, lr2=lr2, niter2=n2, device="cpu", opt_depth=False,
            shared_intrinsics=False, matching_conf_thr=5.0, verbose=False)
        out[tag] = plain(dict(coarse=res_c, fine=res_f))
    # canonical view of image 0 and the dense point cloud (+ clean_pointcloud) of the optimised scene
    res = out["short"]["fine"] or out["short"]["coarse"]
    scene = sparse_ga.SparseGA(filelist, pairs_in, {k: (list(v) if isinstance(v, tuple) else v) for k, v in res.items()},
                               anchors, canonical_paths)
    pts3d, depthmaps, confs = scene.get_dense_pts3d(clean_depth=True)
    _, _, confs_raw = scene.get_dense_pts3d(clean_depth=False)
    canon = [plain(torch.load(p)) for p in canonical_paths]
    # inputs of canonical_view for image 0: its view-1 point maps in tmp_pairs order (sparse_ga.py:655-693)
    pt, cf = [], []
    for (img1, img2), ((path1, path2), path_corres) in tmp_pairs.items():
        if img1 == filelist[0]:
            X, C, _, _ = torch.load(path1)
            pt.append(X); cf.append(C)
        if img2 == filelist[0]:
            X, C, _, _ = torch.load(path2)
            pt.append(X); cf.append(C)
    dense = plain(dict(pts3d=pts3d, depthmaps=depthmaps, confs=confs, confs_raw=confs_raw, canon=canon,
                       canon_in_pts=torch.stack(pt), canon_in_conf=torch.stack(cf),
                       pairwise_scores=pairwise_scores))
    return dict(n_views=n_views, W=W, H=H, seed=seed, low_conf=low_conf, niter=(niter1, niter2), lr=(lr1, lr2),
                inputs=inputs, out=out, dense=dense)


def generate(sparse_ga, out_dir):
    torch.manual_seed(0)
    fx = run_reference(sparse_ga, 3, 64, 48, niter1=30, niter2=20, seed=0, pts_noise=0.02)
    torch.save(fx, os.path.join(out_dir, "align_match3.pt"))
    fx = run_reference(sparse_ga, 3, 64, 48, niter1=30, niter2=20, seed=1, low_conf=True)
    fx["dense"] = {k: v for k, v in fx["dense"].items() if k in ("pairwise_scores",)}
    torch.save(fx, os.path.join(out_dir, "align_dust3r3.pt"))
