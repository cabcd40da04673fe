"""Shared body of the index-plumbing parity tests: the repository's own
forward_mast3r -> prepare_canonical_data -> compute_min_spanning_tree -> condense_data, run on the synthetic model the
fixtures were generated from, against what the UNMODIFIED reference produced on the same model
(oracle/gen_golden_align.py -> tests/golden/align_*.pt: fx["inputs"], fx["dense"]["pairwise_scores" | "canon"]).
tests/test_align_gpu.py runs it on the B200, tests/test_align_lib_emu_host.py on the CPU emulator of the library."""
import os

import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# the arguments oracle/gen_golden_align.py::generate passed to run_reference (pts_noise is not stored in the fixture)
MODEL_KW = {"align_match3.pt": dict(pts_noise=0.02), "align_dust3r3.pt": dict(pts_noise=0.0)}


def plain(x):
    """Same normal form as oracle/gen_golden_align.py::plain (PairOfSlices -> tuple, slice -> ("slice", a, b))."""
    if isinstance(x, torch.Tensor):
        return x.detach().cpu()
    if isinstance(x, slice):
        return ("slice", x.start, x.stop)
    if isinstance(x, dict):
        return {k: plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return tuple(plain(v) for v in x)
    return x


def assert_same(got, want, path, atol=2e-5, rtol=1e-5):
    """Integers, index tensors and structure exactly; floats within atol / rtol."""
    if isinstance(want, torch.Tensor):
        assert isinstance(got, torch.Tensor) and got.shape == want.shape, (path, getattr(got, "shape", type(got)), want.shape)
        if want.dtype.is_floating_point:
            assert torch.allclose(got.to(want.dtype), want, atol=atol, rtol=rtol), (path, (got - want).abs().max().item())
        else:
            assert torch.equal(got.to(want.dtype), want), path
    elif isinstance(want, dict):
        assert set(got) == set(want), (path, sorted(got), sorted(want))
        for k in want:
            assert_same(got[k], want[k], f"{path}[{k!r}]", atol, rtol)
    elif isinstance(want, tuple):
        assert isinstance(got, tuple) and len(got) == len(want), (path, len(got) if isinstance(got, tuple) else type(got), len(want))
        for i, (g, w) in enumerate(zip(got, want)):
            assert_same(g, w, f"{path}[{i}]", atol, rtol)
    elif isinstance(want, float):
        assert abs(float(got) - want) <= atol + 1e-4 * abs(want), (path, got, want)
    else:
        assert got == want, (path, got, want)


def run_pipeline(device, name):
    """Returns (fixture, what the repository's pipeline produced for the same model), both in plain form."""
    from starst3r_b200 import reconstruct as rc
    from starst3r_b200 import synth
    from starst3r_b200.image import prepare_images_for_mast3r
    fx = torch.load(os.path.join(GOLD, name), weights_only=False)
    n, W, H = fx["n_views"], fx["W"], fx["H"]
    model = synth.SyntheticMast3r(n, W, H, seed=fx["seed"], low_conf=fx["low_conf"], arc_deg=90.0, **MODEL_KW[name])
    filelist = [f"{i}.png" for i in range(n)]
    imgs = prepare_images_for_mast3r(model.images())
    pairs_in = rc.convert_dust3r_pairs_naming(filelist, rc.make_pairs(imgs, "complete", None, True))
    cache = f"pipeline-parity-{name}-{device}"
    rc._MEMO.pop(cache, None)
    pairs, _ = rc.forward_mast3r(pairs_in, model, cache_path=cache, subsample=8, desc_conf="desc_conf", device=device)
    tmp_pairs, pairwise_scores, canonical_views, canonical_paths, preds_21 = rc.prepare_canonical_data(
        filelist, pairs, 8, cache_path=cache, mode="avg-angle", device=device)
    mst = rc.compute_min_spanning_tree(pairwise_scores)
    imsizes, pps, base_focals, core_depth, anchors, corres, corres2d, preds_21c = rc.condense_data(
        filelist, tmp_pairs, canonical_views, preds_21, torch.float32)
    memo = rc._memo(cache)
    got = dict(
        pair_keys=list(pairs), pairwise_scores=pairwise_scores,
        canon=[memo["canon"][f] for f in filelist],
        matching_score={k: memo["corres"][k][0] for k in pairs},
        inputs=dict(imgs=filelist, imsizes=imsizes, pps=pps, base_focals=base_focals, core_depth=core_depth,
                    anchors=anchors, corres=corres, corres2d=corres2d, preds_21=preds_21c,
                    mst=(int(mst[0]), [(int(a), int(b)) for a, b in mst[1]])))
    rc._MEMO.pop(cache, None)
    return fx, plain(got)


def check_pipeline_vs_reference(device, name):
    fx, got = run_pipeline(device, name)
    n = fx["n_views"]
    # a6 forward_mast3r: one entry per ordered pair, in make_pairs order (sparse_ga.py:529,562)
    assert len(got["pair_keys"]) == n * (n - 1) and len(set(got["pair_keys"])) == n * (n - 1)
    for (a, b), (conf_score, conf_sum, n_corr) in got["matching_score"].items():
        assert n_corr > 0 and conf_sum > 0 and conf_score > 0              # matching_score tuple, sparse_ga.py:558-559
        assert got["matching_score"][b, a][2] == n_corr                    # the reverse pair reuses the matches (:538-540)
    # a8 pairwise_scores[i, j] = number of correspondences (sparse_ga.py:680): integers, so exactly the reference's
    assert torch.equal(got["pairwise_scores"], fx["dense"]["pairwise_scores"])
    # a8 canonical views + Weiszfeld focals of every image (only the first fixture stores them)
    if "canon" in fx["dense"]:
        assert_same(got["canon"], fx["dense"]["canon"], "canon", atol=2e-5, rtol=2e-5)
    # a9 minimum spanning tree: root and BFS edge order
    assert got["inputs"]["mst"] == (fx["inputs"]["mst"][0], [tuple(e) for e in fx["inputs"]["mst"][1]]) or \
        plain(got["inputs"]["mst"]) == plain(fx["inputs"]["mst"])
    # a8 / a10 prepare_canonical_data + condense_data: sizes, principal points, focals, core depths, anchors (pixels and
    # block indices exactly, depth-ratio offsets to fp32 rounding), correspondence slices, corres2d, sub-sampled preds_21
    for key in ("imgs", "imsizes", "pps", "base_focals", "core_depth", "anchors", "corres", "corres2d", "preds_21"):
        assert_same(got["inputs"][key], plain(fx["inputs"][key]), f"inputs[{key!r}]", atol=2e-5, rtol=2e-5)
    return fx, got
