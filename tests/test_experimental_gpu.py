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


# ---- Scene.run_3dgs_optim with sharded views (gs.SHARD_VIEWS), 2 GPUs ------------------------------------------
def _shard_views_worker(rank, world, port, out_dir):
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import starst3r_b200 as st
    from starst3r_b200 import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    W, H, C, N = 64, 48, 4, 3000
    viewmats, Ks = synth.look_at_cameras(C, W, H)
    sp = synth.random_splats(N, seed=3, scale_mode="rand")
    sp["scales"] = sp["scales"] * 6
    d = {k: v.to(dev) for k, v in sp.items()}
    with torch.no_grad():
        target, _, _ = st.gs.rasterization(d["means"], d["quats"], d["scales"], d["opacities"], d["shN"], viewmats.to(dev),
                                           Ks.to(dev), W, H)
    dist.broadcast(target, 0)        # identical ground truth on every rank

    def make_scene():
        scene = st.Scene(device=dev)
        scene.imgs = [t.clamp(0, 1).cpu().numpy() for t in target]
        scene.c2w = torch.linalg.inv(viewmats).to(dev)
        scene.intrinsics = Ks.to(dev)
        g = torch.Generator().manual_seed(0)
        scene.dense_pts = [(sp["means"] + 0.01 * torch.randn(N, 3, generator=g)).to(dev)]
        scene.dense_cols = [torch.rand(N, 3, generator=g)]
        scene.init_3dgs(init_scale=2e-2)
        return scene
    st.gs.SHARD_VIEWS = True
    sharded = make_scene()
    losses_s = sharded.run_3dgs_optim(10)
    st.gs.SHARD_VIEWS = False
    res = {"losses_s": losses_s, "params_s": {k: v.detach().cpu() for k, v in sharded.gaussians.items()}}
    if rank == 0:
        full = make_scene()
        res["losses_f"] = full.run_3dgs_optim(10)
        res["params_f"] = {k: v.detach().cpu() for k, v in full.gaussians.items()}
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_run_3dgs_optim_sharded_views_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_shard_views_worker, args=(2, 29741, str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(tmp_path, f"rank{k}.pt")) for k in range(2)]
    for k in r[0]["params_s"]:
        assert torch.equal(r[0]["params_s"][k], r[1]["params_s"][k]), f"replicas diverged: {k}"
        assert torch.allclose(r[0]["params_s"][k], r[0]["params_f"][k], rtol=1e-4, atol=4e-4), k
    assert r[0]["losses_s"] == r[1]["losses_s"]
    for a, b in zip(r[0]["losses_s"], r[0]["losses_f"]):
        assert abs(a - b) <= 1e-3 * abs(b), (a, b)
