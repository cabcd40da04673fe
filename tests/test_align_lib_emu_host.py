"""CPU: the ALIGN path of the C ABI (st3r_align_optimize, st3r_canonical_view, st3r_focal_weiszfeld, st3r_dense_points,
st3r_clean_pointcloud) compiled for the host by tests/host/build_emu_lib.py - kernel launches rewritten onto the SIMT
emulator - and driven through the PRODUCT's own Python glue (starst3r_b200.reconstruct) on CPU tensors.  The glue
refuses anything but CUDA by design, so this test (and only the test) swaps the library handle, the stream getter and
torch.cuda.device for host stand-ins; the checks are the ones tests/test_align_gpu.py makes on the B200, against the
fixtures produced by the unmodified reference (oracle/gen_golden_align.py): canonical view, Weiszfeld focal, dense
points, clean_pointcloud, loss + gradients vs autograd, and the 30 + 20 iteration trajectory - for the default kernels
for the default kernels (the optimisation loop as one cooperative launch: the emulator runs its grid as one cluster of
fibers, the grid barrier as a cluster barrier), the launch-per-iteration loop and the first implementation."""
import pytest
import torch

CPU = torch.device("cpu")


@pytest.fixture(params=[7, 3, 0], ids=["default-kernels", "launch-per-iteration", "first-implementation"])
def backend(request, emu_backend, monkeypatch):
    from starst3r_b200 import reconstruct as rc
    monkeypatch.setattr(rc, "ALIGN_VARIANT", request.param)
    yield emu_backend
    emu_backend.st3r_align_set_variant(0)


def test_canonical_view_focal_dense_clean_vs_reference(backend):
    import test_align_gpu as t
    t.test_canonical_view_focal_dense_clean_vs_reference(CPU)


@pytest.mark.parametrize("name,mode", [("align_match3.pt", 0), ("align_match3.pt", 1), ("align_dust3r3.pt", 0)])
def test_loss_and_gradients_vs_autograd(backend, name, mode):
    import test_align_gpu as t
    t.test_kernel_loss_and_gradients_vs_autograd(CPU, name, mode)


@pytest.mark.parametrize("name", ["align_match3.pt", "align_dust3r3.pt"])
def test_optimizer_trajectory_vs_reference(backend, name):
    """The fixture's 30 coarse + 20 fine iterations through st3r_align_optimize: same state as the unmodified reference
    (gauge-invariant comparison)."""
    import test_align_gpu as t
    from starst3r_b200 import reconstruct as rc
    f = t.fx(name)
    inp = f["inputs"]
    tt, meta = rc.flatten_problem(list(inp["imgs"]), inp["imsizes"], inp["pps"].clone(), inp["base_focals"].clone(),
                                  [c.clone() for c in inp["core_depth"]], inp["anchors"], inp["corres"], inp["corres2d"],
                                  inp["preds_21"], inp["mst"], 5.0, CPU)
    N = meta["N"]
    params = dict(pps=(inp["pps"].detach().float() / meta["imsizes"]).contiguous(),
                  log_focals=meta["base_focals"].log().contiguous(),
                  quats=torch.tensor([[0.0, 0, 0, 1]]).repeat(N, 1).contiguous(), trans=torch.zeros(N, 3),
                  log_sizes=torch.zeros(N))
    n1, n2 = f["niter"]
    res_c, hist1, _ = rc._optimize_phase(tt, meta, params, 0, 4 | 8 | 16, 1.1, f["lr"][0], n1, rc.cosine_schedule, 0.01, 1.1)
    res_f, hist2, _ = rc._optimize_phase(tt, meta, params, 1, 4 | 8 | 16 | 2 | 1, 0.4, f["lr"][1], n2, rc.cosine_schedule, 0.01,
                                         1.1)
    assert torch.isfinite(hist1).all() and torch.isfinite(hist2).all() and hist1[-1] < hist1[0]
    t.assert_same_up_to_gauge(t.cpu(res_c), f["out"]["short"]["coarse"], 1e-4)
    t.assert_same_up_to_gauge(t.cpu(res_f), f["out"]["short"]["fine"], 1e-4)


@pytest.mark.parametrize("name", ["align_match3.pt", "align_dust3r3.pt"])
def test_pipeline_index_plumbing_vs_reference(emu_backend, monkeypatch, name):
    """SURVEY 8 rows a6 / a8 / a9 / a10 on the CPU: the repository's forward_mast3r -> prepare_canonical_data ->
    compute_min_spanning_tree -> condense_data (product code + emulated library) on the fixture's synthetic model against
    the arrays the unmodified reference produced (tests/pipeline_parity.py)."""
    from pipeline_parity import check_pipeline_vs_reference
    from starst3r_b200 import match
    from starst3r_b200 import reconstruct as rc
    monkeypatch.setattr(match, "USE_CUDA_GRAPHS", False)
    monkeypatch.setattr(rc, "SHARD_PAIRS", False)
    check_pipeline_vs_reference(CPU, name)


def test_reconstruct_scene_end_to_end(emu_backend, monkeypatch):
    """starster.reconstruct_scene on a synthetic 3-view scene, every stage through the product code and the emulated
    library: matching (exact SIMT nearest neighbour, reciprocal search, merge), canonical views, MST, the sparse global
    alignment, dense points + clean_pointcloud.  The recovered relative poses and focals match the scene's."""
    import math
    from starst3r_b200 import match, synth
    from starst3r_b200 import reconstruct as rc
    monkeypatch.setattr(match, "USE_CUDA_GRAPHS", False)
    monkeypatch.setattr(rc, "SHARD_PAIRS", False)
    W, H, n = 64, 48, 3
    model = synth.SyntheticMast3r(n, W, H, seed=0, device="cpu", arc_deg=70.0)
    imgs = model.images()
    real = rc.run_sparse_ga
    monkeypatch.setattr(rc, "run_sparse_ga", lambda *a, **kw: real(*a, **{**kw, "niter1": 150, "niter2": 60}))
    scene, params = rc.reconstruct_scene(model, imgs, [f"{i}.png" for i in range(n)], CPU)
    pts, depth, confs = scene.get_dense_pts3d(clean_depth=True)
    assert len(pts) == n and pts[0].shape == (H * W, 3) and torch.isfinite(pts[0]).all()
    gt = torch.linalg.inv(model.viewmats)
    est = scene.cam2w.cpu()
    rel_e = torch.linalg.inv(est[0:1]) @ est
    rel_g = torch.linalg.inv(gt[0:1]) @ gt
    for i in range(1, n):
        cosang = ((rel_e[i, :3, :3].T @ rel_g[i, :3, :3]).trace().item() - 1) / 2
        assert math.degrees(math.acos(max(-1.0, min(1.0, cosang)))) < 4.0
        assert torch.nn.functional.cosine_similarity(rel_e[i, :3, 3], rel_g[i, :3, 3], dim=0).item() > 0.98
    assert (scene.intrinsics[:, 0, 0].cpu() / (1.2 * max(W, H)) - 1).abs().max().item() < 0.1
