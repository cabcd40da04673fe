"""Kernel variants that were written after the round's GPU budget was spent: they compile for sm_100a but have not
run on a B200 yet, so they are OFF by default and their parity tests only run with ST3R_EXPERIMENTAL=1
(`ST3R_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -m gpu`).  Each test re-runs an existing parity
test of the default kernels with the variant switched on: same oracle, same fixtures, same tolerances."""
import os

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("ST3R_EXPERIMENTAL") != "1",
                                 reason="unmeasured kernel variants; set ST3R_EXPERIMENTAL=1 to run them")]


@pytest.fixture
def align_variant_3(monkeypatch):
    from starst3r_b200 import reconstruct as rc
    monkeypatch.setattr(rc, "ALIGN_VARIANT", 3)     # segmented loss kernels + clustered Weiszfeld
    yield


@pytest.mark.parametrize("name,mode", [("align_match3.pt", 0), ("align_match3.pt", 1), ("align_dust3r3.pt", 0)])
def test_align_segmented_loss_and_gradients(cuda_device, align_variant_3, name, mode):
    import test_align_gpu as t
    t.test_kernel_loss_and_gradients_vs_autograd(cuda_device, name, mode)


@pytest.mark.parametrize("name", ["align_match3.pt", "align_dust3r3.pt"])
def test_align_segmented_optimizer_vs_reference(cuda_device, align_variant_3, name):
    import test_align_gpu as t
    t.test_optimizer_vs_reference(cuda_device, name)
    t.test_optimizer_long_schedule(cuda_device, name)


def test_align_clustered_weiszfeld_and_pipeline(cuda_device, align_variant_3):
    import test_align_gpu as t
    t.test_canonical_view_focal_dense_clean_vs_reference(cuda_device)
    t.test_scene_add_images_end_to_end(cuda_device)


def test_align_variants_agree(cuda_device):
    """Variant 0 and variant 3 on the same problem: same loss history to fp32 summation-order noise."""
    import torch
    from starst3r_b200 import reconstruct as rc
    from test_align_gpu import fx, run_slam
    f = fx("align_match3.pt")
    out = []
    for v in (0, 3):
        rc.ALIGN_VARIANT = v
        try:
            _, res_c, _, _ = run_slam(f, cuda_device, 30, 0)
        finally:
            rc.ALIGN_VARIANT = 0
        out.append(res_c)
    assert torch.allclose(out[0]["intrinsics"], out[1]["intrinsics"], rtol=1e-4, atol=1e-3)
    for a, b in zip(out[0]["depthmaps"], out[1]["depthmaps"]):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-4)


# ---- split-precision tcgen05 matcher (st3r_nn_tc_set_split, match.NN_SPLIT) ------------------------------------
@pytest.fixture
def nn_split(monkeypatch):
    from starst3r_b200 import match
    monkeypatch.setattr(match, "NN_SPLIT", True)
    yield
    from starst3r_b200 import _lib
    _lib.load().st3r_nn_tc_set_split(0)


@pytest.mark.parametrize("M,N", [(1, 1), (2, 5), (7, 129), (128, 128), (129, 4097), (300, 20000), (1000, 66000)])
def test_split_nn_argmax_vs_oracle(cuda_device, nn_split, M, N):
    import test_match_gpu as t
    t.test_nn_argmax_vs_oracle(cuda_device, "tcgen05", M, N)


def test_split_goldens(golden, cuda_device, nn_split):
    import test_match_gpu as t
    t.test_nn_argmax_golden(golden, cuda_device, "tcgen05")
    t.test_nn_argmax_unnormalised_and_negative(cuda_device, "tcgen05")
    t.test_extract_correspondences_golden(golden, cuda_device, "tcgen05")


def test_split_full_size_properties(cuda_device, nn_split):
    import test_match_gpu as t
    t.test_full_size_properties(cuda_device, "tcgen05")


def test_split_identical_and_resolves_rarely_on_smooth_fields(cuda_device):
    """Same correspondences with and without the split on random and on smooth descriptor fields, and on the smooth
    ones the exact list resolutions per query row drop by an order of magnitude (that is the point of the variant)."""
    import ctypes
    import torch
    from starst3r_b200 import _lib, match, synth
    lib = _lib.load()
    H = W = 256
    g = torch.Generator().manual_seed(0)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    X = torch.stack([xx / W, yy / H, (xx + yy) / (W + H)], -1)
    freq = torch.randn(24, 3, generator=g) * 2.5
    smooth = lambda salt: torch.nn.functional.normalize(                                       # noqa: E731
        torch.cos(X @ freq.T + salt) + 0.003 * torch.randn(H, W, 24, generator=g), dim=-1).to(cuda_device)
    A, B = synth.descriptor_pair(H, W, seed=3, device=cuda_device)
    cases = {"random": [A, B, B, A], "smooth": [smooth(0.0), smooth(0.01), smooth(0.02), smooth(0.03)]}
    q = [torch.ones(H, W, device=cuda_device) for _ in range(4)]

    def stats():
        st = (ctypes.c_ulonglong * 2)()
        _lib.check(lib.st3r_nn_tc_stats(st, 1), "stats")
        return int(st[1]) / max(int(st[0]), 1)
    out, ratio = {}, {}
    try:
        match.NN_COOPERATIVE = False
        for split in (False, True):
            match.NN_SPLIT = split
            for name, feats in cases.items():
                stats()
                out[name, split] = [t.cpu() for t in match.extract_correspondences(feats, q, 8, device=cuda_device)]
                ratio[name, split] = stats()
        for name in cases:
            for a, b in zip(out[name, False], out[name, True]):
                assert torch.equal(a, b), name
        assert ratio["smooth", True] < 0.1 * ratio["smooth", False], ratio
    finally:
        match.NN_SPLIT = False
        match.NN_COOPERATIVE = "auto"
        lib.st3r_nn_tc_set_split(0)


# ---- queue variant of the blend backward (st3r_gs_set_raster_variant, gs.RASTER_BWD_VARIANT) -------------------
@pytest.fixture
def raster_queue(monkeypatch):
    from starst3r_b200 import gs
    monkeypatch.setattr(gs, "RASTER_BWD_VARIANT", 1)
    yield
    from starst3r_b200 import _lib
    _lib.load().st3r_gs_set_raster_variant(0)


def test_raster_queue_backward_vs_autograd(cuda_device, raster_queue):
    import test_gs_gpu as t
    t.test_rasterization_backward_vs_autograd(cuda_device)
    t.test_train_steps_vs_oracle(cuda_device)
    t.test_train_plan_matches_unplanned(cuda_device)


@pytest.mark.parametrize("scale_mult", [1.0, 8.0, 40.0])
def test_raster_queue_equals_default_gradients(cuda_device, scale_mult):
    """Both variants on the same frame, from splats a fraction of a pixel wide (every visit goes through the queues)
    to splats that cover whole tiles (every visit takes the dense path): gradients agree to fp32 summation order."""
    import torch
    from starst3r_b200 import gs, synth
    sp = synth.random_splats(20_000, seed=4)
    sp["scales"] = sp["scales"] * scale_mult
    viewmats, Ks = synth.look_at_cameras(3, 200, 136)
    dev = cuda_device
    args = [sp[k].to(dev) for k in ("means", "quats", "scales", "opacities", "shN")]
    g = torch.Generator().manual_seed(1)
    v_render = torch.randn(3, 136, 200, 3, generator=g).to(dev)
    v_alpha = torch.randn(3, 136, 200, 1, generator=g).to(dev)
    grads = {}
    for variant in (0, 1):
        gs.RASTER_BWD_VARIANT = variant
        try:
            leaves = [a.clone().requires_grad_(True) for a in args]
            render, alpha, _ = gs.rasterization(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], viewmats.to(dev),
                                                Ks.to(dev), 200, 136, sh_degree=1)
            ((render * v_render).sum() + (alpha * v_alpha).sum()).backward()
            grads[variant] = [x.grad.clone() for x in leaves]
        finally:
            gs.RASTER_BWD_VARIANT = 0
    for a, b in zip(grads[0], grads[1]):
        assert torch.isfinite(b).all()
        assert (a - b).abs().max().item() <= 2e-4 * max(a.abs().max().item(), 1e-12)
