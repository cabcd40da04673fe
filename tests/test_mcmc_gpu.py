"""GPU parity of the MCMC strategy kernels (csrc/gs_mcmc.cu through the C ABI) vs oracle/gs_oracle.py mcmc_* on
the same random draws.  Reference call sites: starster/gs.py:43-45,146-147,163-164 (gsplat.MCMCStrategy).
Index plumbing (dead / alive lists, which rows move) is bit-exact; values go through expf / logf / powf whose CUDA
and glibc implementations differ by a few ulp: rtol 2e-5."""
import numpy as np
import pytest
import torch

from oracle import gs_oracle as go

pytestmark = pytest.mark.gpu
RTOL = 2e-5


def make(n, seed=0, n_dead=5, dev="cpu"):
    g = torch.Generator().manual_seed(seed)
    p = {"means": torch.randn(n, 3, generator=g), "scales": torch.randn(n, 3, generator=g) * 0.3 - 3.0,
         "quats": torch.randn(n, 4, generator=g), "opacities": torch.randn(n, generator=g),
         "sh0": torch.randn(n, 1, 3, generator=g), "shN": torch.randn(n, 24, 3, generator=g)}
    if n_dead:
        dead = torch.randperm(n, generator=g)[:n_dead]
        p["opacities"][dead] = -8.0
    m = {k: (torch.rand(v.shape, generator=g), torch.rand(v.shape, generator=g)) for k, v in p.items()}
    return p, m


def device_scene(p, m, dev):
    from starst3r_b200 import gs
    params = {k: torch.nn.Parameter(v.clone().to(dev)) for k, v in p.items()}
    opts = {k: gs.FusedAdam([v]) for k, v in params.items()}
    for k, v in params.items():
        st = opts[k]._st(v)
        st["exp_avg"].copy_(m[k][0])
        st["exp_avg_sq"].copy_(m[k][1])
    return params, opts


@pytest.mark.parametrize("n", [1, 7, 1000, 100_003])
def test_compute_relocation_vs_oracle(cuda_device, n):
    from starst3r_b200 import gs
    g = torch.Generator().manual_seed(n)
    o = torch.rand(n, generator=g) * 0.98 + 0.01
    s = torch.rand(n, 3, generator=g) + 0.01
    r = torch.randint(0, 70, (n,), generator=g, dtype=torch.int32)       # includes 0 and > n_max: clamped
    b = go.mcmc_binoms()
    sub = slice(0, min(n, 300))                                          # the oracle is a python loop
    ro, rs = go.mcmc_compute_relocation(o[sub], s[sub], r[sub], b)
    no, ns = gs.compute_relocation(o.to(cuda_device), s.to(cuda_device), r.to(cuda_device), b.to(cuda_device))
    assert torch.allclose(no.cpu()[sub], ro, rtol=RTOL, atol=1e-7)
    assert torch.allclose(ns.cpu()[sub], rs, rtol=5e-5)
    assert torch.isfinite(no).all() and torch.isfinite(ns).all()


@pytest.mark.parametrize("n,n_dead", [(64, 5), (5000, 250), (5000, 0)])
def test_relocate_vs_oracle(cuda_device, n, n_dead):
    from starst3r_b200 import gs
    p, m = make(n, seed=n, n_dead=n_dead)
    params, opts = device_scene(p, m, cuda_device)
    strat = gs.MCMCStrategy()
    binoms = strat.initialize_state()["binoms"].to(cuda_device)
    g = torch.Generator().manual_seed(1)
    sampled = torch.randint(0, n - n_dead, (n_dead,), generator=g)
    if n_dead > 2:
        sampled[1] = sampled[0]                                          # a source drawn twice
    old_keys = {k: v for k, v in params.items()}
    got = strat._relocate_gs(params, opts, binoms, sampled=sampled.to(cuda_device))
    assert got == n_dead
    go.mcmc_relocate(p, m, sampled, go.mcmc_binoms())
    for k in p:
        assert torch.allclose(params[k].detach().cpu(), p[k], rtol=RTOL, atol=1e-6), k
        st = opts[k].state[params[k]]
        assert torch.equal(st["exp_avg"].cpu(), m[k][0]) and torch.equal(st["exp_avg_sq"].cpu(), m[k][1]), k
        if n_dead:
            assert params[k] is not old_keys[k] and old_keys[k] not in opts[k].state      # re-wrapped like gsplat
            assert opts[k].param_groups[0]["params"][0] is params[k]
    # rows other than opacities / scales are moved bit-exactly
    for k in ("means", "quats", "sh0", "shN"):
        assert torch.equal(params[k].detach().cpu(), p[k]), k


def test_partition_matches_nonzero(cuda_device):
    from starst3r_b200 import _lib
    lib = _lib.load()
    n = 300_001
    g = torch.Generator().manual_seed(3)
    raw = (torch.randn(n, generator=g) * 4).to(cuda_device)
    dead = torch.empty(n, dtype=torch.int32, device=cuda_device)
    alive = torch.empty_like(dead)
    probs = torch.empty(n, device=cuda_device)
    ap = torch.empty(n, device=cuda_device)
    nd = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    ws = torch.empty(lib.st3r_mcmc_partition_ws_bytes(n), dtype=torch.uint8, device=cuda_device)
    import ctypes
    _lib.check(lib.st3r_mcmc_partition(_lib.ptr(raw), n, ctypes.c_float(0.005), _lib.ptr(dead), _lib.ptr(alive),
                                       _lib.ptr(probs), _lib.ptr(ap), _lib.ptr(nd), _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr()), "partition")
    k = int(nd.item())
    mask = probs <= 0.005
    assert k == int(mask.sum().item()) and 0 < k < n
    assert torch.equal(dead[:k].long(), mask.nonzero(as_tuple=True)[0])
    assert torch.equal(alive[:n - k].long(), (~mask).nonzero(as_tuple=True)[0])
    assert torch.equal(ap[:n - k], probs[~mask])
    assert torch.allclose(probs.cpu(), torch.sigmoid(raw.cpu()), rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("n", [10, 4000])
def test_sample_add_vs_oracle(cuda_device, n):
    from starst3r_b200 import gs
    p, m = make(n, seed=n + 1, n_dead=0)
    params, opts = device_scene(p, m, cuda_device)
    strat = gs.MCMCStrategy()
    binoms = strat.initialize_state()["binoms"].to(cuda_device)
    n_new = int(1.05 * n) - n
    g = torch.Generator().manual_seed(2)
    sampled = torch.randint(0, n, (max(n_new, 0),), generator=g)
    got = strat._add_new_gs(params, opts, binoms, sampled=sampled.to(cuda_device))
    assert got == n_new
    if n_new == 0:
        assert params["means"].shape[0] == n
        return
    out_p, out_m = go.mcmc_sample_add(p, m, sampled, go.mcmc_binoms())
    for k in p:
        assert params[k].shape == out_p[k].shape
        assert torch.allclose(params[k].detach().cpu(), out_p[k], rtol=RTOL, atol=1e-6), k
        st = opts[k].state[params[k]]
        assert torch.equal(st["exp_avg"].cpu(), out_m[k][0]) and torch.equal(st["exp_avg_sq"].cpu(), out_m[k][1]), k
    for k in ("means", "quats", "sh0", "shN"):
        assert torch.equal(params[k].detach().cpu(), out_p[k]), k


def test_cap_max_limits_growth(cuda_device):
    from starst3r_b200 import gs
    p, m = make(100, n_dead=0)
    params, opts = device_scene(p, m, cuda_device)
    strat = gs.MCMCStrategy(cap_max=102)
    binoms = strat.initialize_state()["binoms"].to(cuda_device)
    assert strat._add_new_gs(params, opts, binoms) == 2 and params["means"].shape[0] == 102
    assert strat._add_new_gs(params, opts, binoms) == 0


@pytest.mark.parametrize("n", [1, 1000, 200_000])
def test_inject_noise_vs_oracle(cuda_device, n):
    from starst3r_b200 import gs
    p, _ = make(n, seed=5, n_dead=0)
    p["opacities"] = p["opacities"] * 4 - 4          # a spread of gates, many fully open
    noise = torch.randn(n, 3, generator=torch.Generator().manual_seed(9))
    ref = go.mcmc_inject_noise(p, noise, scaler=1e-3 * 5e5)
    params = {k: torch.nn.Parameter(v.clone().to(cuda_device)) for k, v in p.items()}
    gs.inject_noise_to_position(params, 1e-3 * 5e5, noise=noise.to(cuda_device))
    d_ref = ref - p["means"]
    d = params["means"].detach().cpu() - p["means"]
    assert d_ref.abs().max() > 1e-4                  # the test moves something
    # the displacement is a 3x3 product of exp() terms scaled by 500: compare it relative to its own magnitude
    assert torch.allclose(d, d_ref, rtol=1e-3, atol=1e-5 * float(d_ref.abs().max()))
    closed = (1 / (1 + torch.exp(-100 * ((1 - torch.sigmoid(p["opacities"])) - 0.995)))) < 1e-12
    assert torch.equal(params["means"].detach().cpu()[closed], p["means"][closed])      # closed gate: bit-identical


def test_run_3dgs_optim_with_pruning_fires_refine(cuda_device):
    """End to end through the reference API (Scene.run_3dgs_optim(enable_pruning=True), gs.py:146-147,163-164) with
    the refine window opened early (gsplat's defaults only open it after step 500): dead Gaussians are relocated,
    the splat grows by 5 % per refine, optimiser state follows, training continues on the grown splat."""
    import starst3r_b200 as st
    from starst3r_b200 import gs, synth
    W, H, C, N = 64, 48, 3, 2000
    viewmats, Ks = synth.look_at_cameras(C, W, H)
    sp = synth.random_splats(N, seed=3, scale_mode="rand")
    sp["scales"] = sp["scales"] * 6
    d = {k: v.to(cuda_device) for k, v in sp.items()}
    with torch.no_grad():
        target, _, _ = st.gs.rasterization(d["means"], d["quats"], d["scales"], d["opacities"], d["shN"],
                                           viewmats.to(cuda_device), Ks.to(cuda_device), W, H)
    scene = st.Scene(device=cuda_device)
    scene.imgs = [t.clamp(0, 1).cpu().numpy() for t in target]
    scene.c2w = torch.linalg.inv(viewmats).to(cuda_device)
    scene.intrinsics = Ks.to(cuda_device)
    g = torch.Generator().manual_seed(0)
    scene.dense_pts = [(sp["means"] + 0.01 * torch.randn(N, 3, generator=g)).to(cuda_device)]
    scene.dense_cols = [torch.rand(N, 3, generator=g)]
    scene.init_3dgs(init_scale=2e-2)
    assert isinstance(scene.strategy, gs.MCMCStrategy) and scene.strategy.refine_start_iter == 500
    assert scene.strategy_state["binoms"].shape == (51, 51)
    scene.strategy = gs.MCMCStrategy(refine_start_iter=2, refine_every=4, cap_max=2150)
    with torch.no_grad():
        scene.gaussians["opacities"][:100] = -9.0          # dead for the strategy (sigmoid <= 0.005)
    torch.manual_seed(0)
    losses = scene.run_3dgs_optim(10, enable_pruning=True)  # refines at steps 4 and 8
    assert len(losses) == 10 and all(np.isfinite(losses))
    n = scene.gaussians["means"].shape[0]
    assert n == 2150                                        # 2000 -> 2100 -> min(cap_max, 2205)
    for k, v in scene.gaussians.items():
        assert isinstance(v, torch.nn.Parameter) and v.shape[0] == n
        stt = scene.optimizers[k].state.get(v)
        if k != "sh0":
            assert stt["exp_avg"].shape == v.shape and int(stt["step"].item()) == 10
    assert (torch.sigmoid(scene.gaussians["opacities"]) <= 0.005).sum().item() == 0      # all relocated
    r, a, info = scene.render_3dgs_original(W, H)
    assert r.shape == (C, H, W, 3) and torch.isfinite(r).all()
