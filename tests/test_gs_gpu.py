"""GPU parity of the RASTER + ADAM path: CUDA kernels (through the C ABI) vs the CPU oracle (oracle/gs_oracle.py).
Tile / bin indices must be bit-exact; rendered RGB, gradients and optimiser state within the stated tolerances."""
import copy
import math

import numpy as np
import pytest
import torch

from oracle import gs_oracle as go
from starst3r_b200 import synth

pytestmark = pytest.mark.gpu


def small_scene(n=500, C=3, W=80, H=48, seed=0, scale_mult=8.0):
    sp = synth.random_splats(n, seed=seed, scale_mode="rand")
    sp["scales"] = sp["scales"] * scale_mult
    viewmats, Ks = synth.look_at_cameras(C, W, H)
    return sp, viewmats, Ks, W, H


def to(dev, d):
    return {k: v.to(dev) for k, v in d.items()}


def psnr(a, b):
    mse = ((a - b) ** 2).mean().item()
    return 10 * math.log10(1.0 / max(mse, 1e-20))


# ------------------------------------------------------------------------------------------ primitives
@pytest.mark.parametrize("n", [0, 1, 5, 4095, 4096, 4097, 100_000, 3_000_001])
def test_exclusive_scan(cuda_device, n):
    from starst3r_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 9, (max(n, 1),), generator=g, dtype=torch.int32)[:n].to(cuda_device)
    out = torch.empty_like(x)
    total = torch.full((1,), -1, dtype=torch.int32, device=cuda_device)
    ws = torch.empty(lib.st3r_scan_ws_bytes(n), dtype=torch.uint8, device=cuda_device)
    _lib.check(lib.st3r_exclusive_scan_i32(_lib.ptr(x), _lib.ptr(out), n, _lib.ptr(total), _lib.ptr(ws), ws.numel(),
                                           _lib.stream_ptr()), "scan")
    ref = torch.cumsum(x.long(), 0) - x.long()
    assert torch.equal(out.long(), ref)
    assert total.item() == int(x.long().sum().item())


@pytest.mark.parametrize("n,bits", [(1, 64), (1000, 64), (4096, 40), (4097, 45), (250_000, 47), (2_000_003, 45)])
def test_radix_sort_pairs_stable(cuda_device, n, bits):
    from starst3r_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(n)
    hi = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64)
    lo = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64)
    keys = ((hi << 32) | lo) & ((1 << min(bits, 62)) - 1)
    keys[::7] = keys[0]                     # many duplicates -> stability is observable through the values
    keys = keys.to(cuda_device)
    vals = torch.arange(n, dtype=torch.int32, device=cuda_device)
    k, v = keys.clone(), vals.clone()
    ka, va = torch.empty_like(k), torch.empty_like(v)
    nptr = torch.tensor([n], dtype=torch.int32, device=cuda_device)
    ws = torch.empty(lib.st3r_radix_sort_ws_bytes(n), dtype=torch.uint8, device=cuda_device)
    _lib.check(lib.st3r_radix_sort_pairs(_lib.ptr(k), _lib.ptr(v), _lib.ptr(ka), _lib.ptr(va), _lib.ptr(nptr), n, 0,
                                         bits, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "sort")
    rk, order = torch.sort(keys, stable=True)
    assert torch.equal(k, rk)
    assert torch.equal(v.long(), order)


# ------------------------------------------------------------------------------------------ forward
def test_rasterization_indices_bit_exact_and_rgb(cuda_device):
    from starst3r_b200 import gs
    sp, viewmats, Ks, W, H = small_scene()
    r_ref, a_ref, info_ref = go.rasterization(sp["means"], sp["quats"], sp["scales"], sp["opacities"], sp["shN"],
                                              viewmats, Ks, W, H)
    d = to(cuda_device, sp)
    with torch.no_grad():
        r, a, info = gs.rasterization(d["means"], d["quats"], d["scales"], d["opacities"], d["shN"],
                                      viewmats.to(cuda_device), Ks.to(cuda_device), W, H, sh_degree=1)
    assert r.shape == (3, H, W, 3) and a.shape == (3, H, W, 1)
    assert len(info_ref["isect_ids"]) > 2000
    for k in ("camera_ids", "gaussian_ids", "radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert np.array_equal(info[k].cpu().numpy(), np.asarray(info_ref[k])), k
    assert info["isect_ids"].dtype == torch.int64 and info["flatten_ids"].dtype == torch.int32
    assert info["tile_width"] == 5 and info["tile_height"] == 3 and info["n_cameras"] == 3
    # projection floats are produced by the same single-rounded operation sequence -> identical
    assert torch.equal(info["means2d"].cpu(), info_ref["means2d"])
    assert torch.equal(info["depths"].cpu(), info_ref["depths"])
    assert torch.equal(info["conics"].cpu(), info_ref["conics"])
    assert np.array_equal(info["last_ids"].cpu().numpy(), info_ref["last_ids"])
    # tolerance: blend uses ex2.approx (__expf) and a different summation order than the oracle
    assert torch.allclose(r.cpu(), r_ref, atol=2e-5, rtol=1e-4)
    assert torch.allclose(a.cpu(), a_ref, atol=2e-5, rtol=1e-4)
    assert psnr(r.cpu(), r_ref) > 80.0


def test_rasterization_ragged_image_and_empty(cuda_device):
    """Image size not a multiple of the tile size; a camera that sees nothing; zero Gaussians."""
    from starst3r_b200 import gs
    sp, viewmats, Ks, W, H = small_scene(n=300, C=2, W=70, H=37, seed=5)
    viewmats[1, :3, 3] += torch.tensor([0.0, 0.0, -50.0])       # camera 1 looks away: everything behind it
    r_ref, a_ref, info_ref = go.rasterization(sp["means"], sp["quats"], sp["scales"], sp["opacities"], sp["shN"],
                                              viewmats, Ks, W, H)
    d = to(cuda_device, sp)
    with torch.no_grad():
        r, a, info = gs.rasterization(d["means"], d["quats"], d["scales"], d["opacities"], d["shN"],
                                      viewmats.to(cuda_device), Ks.to(cuda_device), W, H)
    for k in ("camera_ids", "gaussian_ids", "radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert np.array_equal(info[k].cpu().numpy(), np.asarray(info_ref[k])), k
    assert torch.allclose(r.cpu(), r_ref, atol=2e-5, rtol=1e-4)
    assert r[1].abs().max().item() == 0.0
    e = {k: v[:0].contiguous() for k, v in d.items()}
    with torch.no_grad():
        r0, a0, info0 = gs.rasterization(e["means"], e["quats"], e["scales"], e["opacities"], e["shN"],
                                         viewmats.to(cuda_device), Ks.to(cuda_device), W, H)
    assert r0.abs().max().item() == 0.0 and info0["isect_ids"].numel() == 0


# ------------------------------------------------------------------------------------------ backward
def test_rasterization_backward_vs_autograd(cuda_device):
    from starst3r_b200 import gs
    sp, viewmats, Ks, W, H = small_scene(n=400, seed=2)
    g = torch.Generator().manual_seed(0)
    wr = torch.randn(3, H, W, 3, generator=g)
    wa = torch.randn(3, H, W, 1, generator=g)
    ref = {k: v.clone().requires_grad_(True) for k, v in sp.items()}
    r_ref, a_ref, _ = go.rasterization(ref["means"], ref["quats"], ref["scales"], ref["opacities"], ref["shN"],
                                       viewmats, Ks, W, H)
    ((r_ref * wr).sum() + (a_ref * wa).sum()).backward()
    d = {k: v.to(cuda_device).requires_grad_(True) for k, v in sp.items()}
    r, a, _ = gs.rasterization(d["means"], d["quats"], d["scales"], d["opacities"], d["shN"],
                               viewmats.to(cuda_device), Ks.to(cuda_device), W, H)
    ((r * wr.to(cuda_device)).sum() + (a * wa.to(cuda_device)).sum()).backward()
    for k in ("means", "quats", "scales", "opacities", "shN"):
        got, want = d[k].grad.cpu(), ref[k].grad
        scale = want.abs().max().item() + 1e-12
        err = (got - want).abs().max().item()
        assert err / scale < 5e-3, (k, err, scale)   # fp32 atomics + __expf vs exp; typical error is ~1e-5
        assert ((got - want).abs().mean() / (want.abs().mean() + 1e-12)).item() < 1e-3, k


def test_visit_list_kernels_vs_oracle(cuda_device, monkeypatch):
    """The oracle parity tests above ran the fragment-pool kernels (the default); the same checks for the visit-list
    pair (st3r_gs_set_raster_variant(1)), which serves as the independent cross-check of the default kernels."""
    from starst3r_b200 import gs
    monkeypatch.setattr(gs, "RASTER_VARIANT", 1)
    test_rasterization_indices_bit_exact_and_rgb(cuda_device)
    test_rasterization_ragged_image_and_empty(cuda_device)
    test_rasterization_backward_vs_autograd(cuda_device)


@pytest.mark.parametrize("scale_mult", [1.0, 8.0, 40.0, 150.0])
def test_blend_kernel_pairs_agree(cuda_device, scale_mult):
    """The two independent implementations of the blend (fragment-pool and visit-list kernels) on the same frame, from
    splats a fraction of a pixel wide (pool batches) over a mix to splats that cover whole tiles (dense batches, pool
    overflow): forward outputs identical bit for bit (same per-pixel arithmetic in the same order), gradients equal up
    to fp32 summation order."""
    from starst3r_b200 import gs
    sp = synth.random_splats(20_000, seed=4, scale_mode="rand")
    sp["scales"] = sp["scales"] * scale_mult / 8.0
    viewmats, Ks = synth.look_at_cameras(3, 200, 136)
    dev = cuda_device
    args = [sp[k].to(dev) for k in ("means", "quats", "scales", "opacities", "shN")]
    g = torch.Generator().manual_seed(1)
    v_render = torch.randn(3, 136, 200, 3, generator=g).to(dev)
    v_alpha = torch.randn(3, 136, 200, 1, generator=g).to(dev)
    out = {}
    for variant in (0, 1):
        gs.RASTER_VARIANT = variant
        try:
            leaves = [a.clone().requires_grad_(True) for a in args]
            render, alpha, info = gs.rasterization(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], viewmats.to(dev),
                                                   Ks.to(dev), 200, 136, sh_degree=1)
            ((render * v_render).sum() + (alpha * v_alpha).sum()).backward()
            out[variant] = (render.detach(), alpha.detach(), info["last_ids"], [x.grad.clone() for x in leaves])
        finally:
            gs.RASTER_VARIANT = 0
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    assert out[0][0].abs().max().item() > 0.1
    for name, a, b in zip(("means", "quats", "scales", "opacities", "shN"), out[0][3], out[1][3]):
        assert torch.isfinite(a).all() and torch.isfinite(b).all()
        assert (a - b).abs().max().item() <= 1e-3 * b.abs().max().item(), (name, scale_mult)
        assert ((a - b).abs().mean() / (b.abs().mean() + 1e-20)).item() < 1e-4, (name, scale_mult)


# ------------------------------------------------------------------------------------------ loss
def test_loss_forward_backward_vs_oracle(cuda_device):
    from starst3r_b200 import _lib
    lib = _lib.load()
    C, H, W = 2, 45, 52
    g = torch.Generator().manual_seed(4)
    truth = torch.rand(C, H, W, 3, generator=g)
    render = (truth + 0.2 * torch.randn(C, H, W, 3, generator=g)).clamp(0, 1.2).requires_grad_(True)
    zero = torch.zeros(1)
    loss_ref = sum(go.compute_loss(truth[i], render[i], zero - 1e9, zero - 1e9, 0.2, 0.0, 0.0) for i in range(C))
    loss_ref.backward()
    rd, td = render.detach().to(cuda_device).contiguous(), truth.to(cuda_device).contiguous()
    dmaps = torch.empty(C, H, W, 3, 3, device=cuda_device)
    sums = torch.zeros(C, 2, device=cuda_device)
    v = torch.empty_like(rd)
    _lib.check(lib.st3r_gs_loss_fwd(_lib.ptr(rd), _lib.ptr(td), C, H, W, 0.2, _lib.ptr(dmaps), _lib.ptr(sums),
                                    _lib.stream_ptr()), "loss_fwd")
    _lib.check(lib.st3r_gs_loss_bwd(_lib.ptr(rd), _lib.ptr(td), _lib.ptr(dmaps), C, H, W, 0.2, _lib.ptr(v),
                                    _lib.stream_ptr()), "loss_bwd")
    l1 = sums[:, 1] / (3.0 * H * W)
    ss = sums[:, 0] / (3.0 * (H - 10) * (W - 10))
    loss = (0.8 * l1 + 0.2 * (1 - ss)).sum().item()
    assert abs(loss - loss_ref.item()) < 2e-6 * max(1.0, abs(loss_ref.item()))
    want = render.grad
    assert torch.allclose(v.cpu(), want, atol=2e-7 + 1e-4 * want.abs().max().item(), rtol=1e-3)


# ------------------------------------------------------------------------------------------ Adam
def test_fused_adam_vs_torch(cuda_device):
    from starst3r_b200 import gs
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, 7, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt_ref = torch.optim.Adam([ref], lr=1e-3)
    mine = torch.nn.Parameter(p0.clone().to(cuda_device))
    opt = gs.FusedAdam([mine], lr=1e-3)
    for step in range(5):
        gr = torch.randn(1000, 7, generator=g) * (10.0 ** (step - 2))
        ref.grad = gr.clone()
        opt_ref.step()
        mine.grad = gr.to(cuda_device)
        opt.step()
    assert torch.allclose(mine.data.cpu(), ref.data, atol=1e-7, rtol=1e-6)
    st = opt.state[mine]
    assert torch.allclose(st["exp_avg"].cpu(), opt_ref.state[ref]["exp_avg"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(st["exp_avg_sq"].cpu(), opt_ref.state[ref]["exp_avg_sq"], rtol=1e-5, atol=1e-9)


def test_adam_step_dev_equals_host_form(cuda_device):
    """st3r_adam_step_dev (step number on the device, bias corrections evaluated there: the form a captured CUDA graph
    replays) == st3r_adam_step, step by step, and the counter advances."""
    from starst3r_b200 import gs
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(777, 5, generator=g)
    pa, pb = p0.clone().to(cuda_device), p0.clone().to(cuda_device)
    ma, va, mb, vb = (torch.zeros_like(pa) for _ in range(4))
    done = torch.full((1,), 0, dtype=torch.int32, device=cuda_device)
    for step in list(range(1, 7)) + [1000, 1001]:
        if step == 1000:
            done.fill_(999)
        gr = (torch.randn(777, 5, generator=g) * (10.0 ** ((step % 5) - 2))).to(cuda_device)
        gs.adam_step([(pa, gr, ma, va, 777, 5, 5, 5)], 1e-3, (0.9, 0.999), 1e-8, step)
        gs.adam_step([(pb, gr, mb, vb, 777, 5, 5, 5)], 1e-3, (0.9, 0.999), 1e-8, None, steps_done=done)
        assert int(done.item()) == step
        assert torch.equal(ma, mb) and torch.equal(va, vb)
        assert (pa - pb).abs().max().item() <= 1e-9, step


def test_adam_layouts_bit_identical(cuda_device):
    """Every memory layout the kernel distinguishes (flat float4 + scalar tail, strided rows moved as float4 - the shN
    [N, 24, 3] segment of gs.py:37 -, strided scalar rows, unaligned bases) gives the bits of the plain contiguous
    update."""
    from starst3r_b200 import gs
    g = torch.Generator().manual_seed(11)
    hp = (1e-3, (0.9, 0.999), 1e-8)

    def run(rows, cols, ld_p, ld_g, off=0):
        grad = torch.randn(rows, cols, generator=g)
        p0 = torch.randn(rows, cols, generator=g)
        # reference: contiguous tensors
        pr, mr, vr = p0.clone().to(cuda_device), torch.zeros(rows, cols, device=cuda_device), torch.zeros(rows, cols, device=cuda_device)
        gr = grad.to(cuda_device)
        # layout under test: rows inside wider buffers, optionally shifted by `off` floats
        P = torch.full((rows * ld_p + off + 8,), 7.0, device=cuda_device)
        M, V = torch.zeros_like(P), torch.zeros_like(P)
        G = torch.full((rows * ld_g + off + 8,), 9.0, device=cuda_device)
        pv = P[off:off + rows * ld_p].view(rows, ld_p)
        gv = G[off:off + rows * ld_g].view(rows, ld_g)
        pv[:, :cols] = pr
        gv[:, :cols] = gr
        for step in (1, 2, 3):
            gs.adam_step([(pr, gr, mr, vr, rows, cols, cols, cols)], *hp, step)
            gs.adam_step([(P[off:], G[off:], M[off:], V[off:], rows, cols, ld_p, ld_g)], *hp, step)
        mv, vv = M[off:off + rows * ld_p].view(rows, ld_p), V[off:off + rows * ld_p].view(rows, ld_p)
        assert torch.equal(pv[:, :cols], pr) and torch.equal(mv[:, :cols], mr) and torch.equal(vv[:, :cols], vr)
        if ld_p > cols:                      # the padding of every row is untouched
            assert bool((pv[:, cols:] == 7.0).all()) and bool((mv[:, cols:] == 0).all())
        assert bool((P[off + rows * ld_p:] == 7.0).all())

    run(1001, 3, 3, 3)            # flat, 3003 floats: float4 body + 3-element tail
    run(500, 12, 72, 12)          # the shN segment: 12 of 72 floats per row, float4 rows
    run(333, 3, 5, 3)             # strided scalar rows
    run(257, 4, 4, 4, off=1)      # base addresses not 16-byte aligned
    run(64, 8, 12, 8, off=2)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_gradient_kernels_on_one_device(cuda_device, world):
    """The multi-GPU exchange kernels with every rank's buffer on ONE device (the kernels only see addresses):
    st3r_adam_step_peers == st3r_adam_step on the rank-ordered sum, bit for bit, for the float4 rows, the strided shN rows
    and an offset that is not a multiple of 4 floats (scalar peer loads); st3r_grad_reduce_scatter run once per rank
    leaves that same sum in every rank's `reduced` buffer."""
    import ctypes
    from starst3r_b200 import _lib, gs
    lib = _lib.load()
    g = torch.Generator().manual_seed(17 + world)
    N = 501
    shapes = [(N, 3, 3, 3), (N, 4, 4, 4), (N, 12, 72, 12), (N, 1, 1, 1)]            # rows, cols, ld_param, ld_grad
    offs, L = [], 0
    for k, (r, c, ldp, ldg) in enumerate(shapes):
        if k == 3:
            L += 1                              # the last segment starts at an odd float: the scalar path
        offs.append(L)
        L += (r * ldg + 3) // 4 * 4
    L = (L + 3) // 4 * 4
    peers = [torch.randn(L, generator=g).to(cuda_device) for _ in range(world)]
    gsum = peers[0].clone()
    for k in range(1, world):
        gsum = gsum + peers[k]                   # rank order, like the kernels
    hp = (1e-3, (0.9, 0.999), 1e-8)

    def state():
        gg = torch.Generator().manual_seed(3)
        out = []
        for r, c, ldp, ldg in shapes:
            out.append((torch.randn(r * ldp, generator=gg).to(cuda_device), torch.zeros(r * ldp, device=cuda_device),
                        torch.zeros(r * ldp, device=cuda_device)))
        return out
    ref, got = state(), state()
    for step in (1, 2):
        gs.adam_step([(p, gsum[o:], m, v, r, c, ldp, ldg) for (p, m, v), o, (r, c, ldp, ldg) in zip(ref, offs, shapes)], *hp, step)
        gs.adam_step_peers([(p, None, m, v, r, c, ldp, ldg) for (p, m, v), (r, c, ldp, ldg) in zip(got, shapes)], offs,
                           [t.data_ptr() for t in peers], *hp, step)
    for (pa, ma, va), (pb, mb, vb) in zip(ref, got):
        assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb)
    # reduce-scatter + all-gather form: rank r sums its slice and stores it into every rank's `reduced` buffer
    reduced = [torch.full((L,), float("nan"), device=cuda_device) for _ in range(world)]
    gp = (ctypes.c_void_p * world)(*[t.data_ptr() for t in peers])
    rp = (ctypes.c_void_p * world)(*[t.data_ptr() for t in reduced])
    for rank in range(world):
        with torch.cuda.device(cuda_device):
            _lib.check(lib.st3r_grad_reduce_scatter(world, rank, gp, rp, ctypes.c_int64(L), _lib.stream_ptr()),
                       "st3r_grad_reduce_scatter")
    for t in reduced:
        assert torch.equal(t, gsum)


# ------------------------------------------------------------------------------------------ full train step
def test_train_steps_vs_oracle(cuda_device):
    """Three iterations of gs.py:143-161 (render 3 views, loss, backward, Adam): loss, gradients and updated
    parameters vs the CPU oracle; PSNR of the final renders within 0.1 dB."""
    from starst3r_b200 import gs
    sp, viewmats, Ks, W, H = small_scene(n=400, C=3, W=64, H=48, seed=7)
    g = torch.Generator().manual_seed(1)
    truth = torch.rand(3, H, W, 3, generator=g)
    ref_p = {k: v.clone() for k, v in sp.items()}
    ref_s = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in sp.items()}
    dev_p = {k: v.clone().to(cuda_device).contiguous() for k, v in sp.items()}
    dev_s = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in dev_p.items()}
    cams = gs.make_cams(viewmats.to(cuda_device), Ks.to(cuda_device))
    td = truth.to(cuda_device).contiguous()
    for step in range(1, 4):
        loss_ref, grads_ref, render_ref, _ = go.train_step(ref_p, ref_s, truth, viewmats, Ks, W, H, step)
        for p in ref_p.values():
            p.requires_grad_(False)
        loss, fr = gs.train_step(dev_p, dev_s, td, cams, W, H, step)
        assert abs(loss.item() - loss_ref) < 1e-5 * abs(loss_ref), (step, loss.item(), loss_ref)
        for k, kk in (("means", "means"), ("quats", "quats"), ("scales", "scales"), ("opacities", "opacities")):
            want = grads_ref[kk]
            got = fr.grads[k].cpu()
            assert (got - want).abs().max().item() < 5e-3 * (want.abs().max().item() + 1e-12), (step, k)
        want = grads_ref["shN"][:, :4]
        assert (fr.grads["sh"].cpu() - want).abs().max().item() < 5e-3 * (want.abs().max().item() + 1e-12)
        assert abs(psnr(fr.render.cpu(), truth) - psnr(render_ref, truth)) < 0.1
    # Adam's first steps are sign-like (|update| ~ lr), so compare parameters with an lr-scaled tolerance
    for k in ref_p:
        diff = (dev_p[k].cpu() - ref_p[k].detach()).abs().max().item()
        assert diff < 2.5e-3, (k, diff)
        frac_close = ((dev_p[k].cpu() - ref_p[k].detach()).abs() < 1e-5).float().mean().item()
        assert frac_close > 0.97, (k, frac_close)


@pytest.mark.parametrize("n,scale_mult,C,W,H", [(500, 8.0, 3, 80, 48), (3000, 60.0, 2, 64, 64), (20000, 40.0, 1, 40, 40),
                                                 (50, 1.0, 2, 33, 17), (2500, 6.0, 1, 96, 64), (6000, 8.0, 1, 160, 112)])
def test_fused_binning_equals_radix_chain(cuda_device, n, scale_mult, C, W, H):
    """st3r_gs_bin_tiles (counting sort by tile + in-tile sort) == st3r_gs_isect + st3r_radix_sort_pairs +
    st3r_gs_offsets, bit for bit (isect_ids, flatten_ids, isect_offsets), for all three per-tile sorts (32-bit surrogate
    keys + repair passes = the default, the 64-bit register-resident network, the shared-memory / in-place network), including tiles whose list is longer than
    the shared-memory sort (4096 pairs; the 20000-Gaussian case has ~20000 per tile)."""
    from starst3r_b200 import _lib, gs
    lib = _lib.load()
    sp, viewmats, Ks, W, H = small_scene(n=n, C=C, W=W, H=H, seed=n, scale_mult=scale_mult)
    d = to(cuda_device, sp)
    out = {}
    for mode, variant in (("radix", 2), ("fused", 2), ("fused", 1), ("fused", 0)):
        gs.BINNING = mode
        _lib.check(lib.st3r_gs_bin_set_variant(variant), "st3r_gs_bin_set_variant")
        try:
            with torch.no_grad():
                r, a, info = gs.rasterization(d["means"], d["quats"], d["scales"], d["opacities"], d["shN"],
                                              viewmats.to(cuda_device), Ks.to(cuda_device), W, H)
        finally:
            gs.BINNING = "fused"
            _lib.check(lib.st3r_gs_bin_set_variant(2), "st3r_gs_bin_set_variant")
        out[mode, variant] = (r, a, info)
    ia = out["radix", 2][2]
    per_tile = torch.diff(ia["isect_offsets"].flatten())
    if n == 20000:
        assert per_tile.max().item() > 4096          # exercises the global-memory sort path
    if n == 2500:                                    # 2, 4 and 8 elements per thread of the register-resident sort
        for lo, hi in ((257, 512), (513, 1024), (1025, 2048)):
            assert ((per_tile >= lo) & (per_tile <= hi)).any(), (lo, hi)
    for key in (("fused", 2), ("fused", 1), ("fused", 0)):
        ib = out[key][2]
        assert ia["isect_ids"].numel() == ib["isect_ids"].numel() and ia["isect_ids"].numel() > 0
        for k in ("isect_ids", "flatten_ids", "isect_offsets"):
            assert torch.equal(ia[k], ib[k]), (k, key)
        assert torch.equal(out["radix", 2][0], out[key][0]) and torch.equal(out["radix", 2][1], out[key][1])


@pytest.mark.parametrize("levels", [0, 64, 3])
def test_tile_sort_variants_with_tied_depths(cuda_device, levels):
    """st3r_gs_bin_tiles on synthetic (radius, centre, depth) records, the three per-tile sorts bit for bit: continuous
    depths, depths on 64 levels (runs of equal leading bits AND exact ties, which the entry index breaks) and on 3 levels
    (the repair passes of the surrogate-key sort give up and the shared-memory network finishes)."""
    from starst3r_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(5 + levels)
    N, C, W, H = 6000, 2, 128, 96
    tw, th = W // 16, H // 16
    radii = torch.randint(1, 40, (C * N,), generator=g, dtype=torch.int32)
    radii[torch.rand(C * N, generator=g) < 0.1] = 0
    geomA = torch.zeros(C * N, 4)
    geomA[:, 0] = torch.rand(C * N, generator=g) * W
    geomA[:, 1] = torch.rand(C * N, generator=g) * H
    depth = 0.5 + 9.5 * torch.rand(C * N, generator=g)
    if levels:
        depth = 0.5 + torch.floor(depth * levels / 10.0) * (10.0 / levels)
    geomA[:, 3] = depth
    radii, geomA = radii.to(cuda_device), geomA.to(cuda_device)
    res = {}
    try:
        for variant in (2, 1, 0):
            _lib.check(lib.st3r_gs_bin_set_variant(variant), "st3r_gs_bin_set_variant")
            offsets = torch.zeros(C * tw * th, dtype=torch.int32, device=cuda_device)
            n_dev = torch.zeros(1, dtype=torch.int32, device=cuda_device)
            ws = torch.empty(lib.st3r_gs_bin_ws_bytes(C, W, H, 16, 0), dtype=torch.uint8, device=cuda_device)
            with torch.cuda.device(cuda_device):
                _lib.check(lib.st3r_gs_bin_tiles(_lib.ptr(radii), _lib.ptr(geomA), N, C, W, H, 16, _lib.ptr(offsets), _lib.ptr(n_dev),
                                                 None, None, 0, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "count")
                n = int(n_dev.item())
                keys = torch.zeros(n, dtype=torch.int64, device=cuda_device)
                vals = torch.zeros(n, dtype=torch.int32, device=cuda_device)
                ws = torch.empty(lib.st3r_gs_bin_ws_bytes(C, W, H, 16, n), dtype=torch.uint8, device=cuda_device)
                _lib.check(lib.st3r_gs_bin_tiles(_lib.ptr(radii), _lib.ptr(geomA), N, C, W, H, 16, _lib.ptr(offsets), _lib.ptr(n_dev),
                                                 _lib.ptr(keys), _lib.ptr(vals), n, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "bin")
            res[variant] = (offsets.cpu(), keys.cpu(), vals.cpu())
    finally:
        _lib.check(lib.st3r_gs_bin_set_variant(2), "st3r_gs_bin_set_variant")
    per_tile = torch.diff(res[0][0])
    assert ((per_tile > 256) & (per_tile <= 2048)).any()
    assert (torch.diff(res[0][1]) >= 0).all()                   # sorted by (camera, tile, depth)
    for variant in (2, 1):
        for a, b in zip(res[variant], res[0]):
            assert torch.equal(a, b), variant


def test_golden_vectors_from_gsplat_when_present(cuda_device):
    """The CUDA path against tests/golden/raster_*.npz (gsplat 1.4 outputs written by oracle/gen_golden_raster.py on a
    machine that has gsplat): skipped, with that reason, while RASTER parity is unpinned."""
    import glob
    import os
    import numpy as np
    from starst3r_b200 import gs
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_*.npz")))
    if not files:
        pytest.skip("RASTER parity unpinned: no tests/golden/raster_*.npz (oracle/gen_golden_raster.py needs gsplat 1.4)")
    for path in files:
        z = np.load(path)
        t = {k: torch.from_numpy(z[k]).to(cuda_device) for k in ("means", "quats", "scales", "opacities", "colors", "viewmats", "Ks")}
        with torch.no_grad():
            render, alpha, info = gs.rasterization(t["means"], t["quats"], t["scales"], t["opacities"], t["colors"],
                                                   t["viewmats"], t["Ks"], int(z["width"]), int(z["height"]))
        assert np.array_equal(info["isect_offsets"].cpu().numpy(), z["isect_offsets"])
        assert np.array_equal(info["flatten_ids"].cpu().numpy(), z["flatten_ids"])
        assert np.abs(render.cpu().numpy() - z["render"]).max() < 2e-5
        assert np.abs(alpha.cpu().numpy() - z["alpha"]).max() < 2e-5


@pytest.mark.parametrize("N,C,W,H", [(200_000, 8, 512, 512), (1_000_000, 4, 1024, 768), (3_000_000, 8, 1920, 1072)])
def test_full_size_properties(cuda_device, N, C, W, H):
    """BASELINE.json configs[1] (8 views 512x512, 200 k Gaussians), a configs[2]-sized frame (1 M Gaussians, 1024x768)
    and one rank's share of configs[3] (8 of the 64 views at 1920x1072, 3 M Gaussians): the oracle is too slow here, so
    the pipeline is checked through size-independent properties."""
    import starst3r_b200 as st
    from starst3r_b200 import gs
    viewmats, Ks = synth.look_at_cameras(C, W, H, device=cuda_device)
    sp = synth.random_splats(N, seed=7, scale_mode="init", device=cuda_device)
    args = (sp["means"], sp["quats"], sp["scales"], sp["opacities"], sp["shN"], viewmats, Ks, W, H)
    with torch.no_grad():
        r1, a1, info = gs.rasterization(*args)
        r2, a2, info2 = gs.rasterization(*args)
    n = info["isect_ids"].numel()
    assert n > N                                                   # a real workload
    # determinism / idempotence: the whole forward is bit-reproducible
    assert torch.equal(r1, r2) and torch.equal(a1, a2) and torch.equal(info["isect_ids"], info2["isect_ids"])
    # sortedness, a checksum of checksums on the bin indices
    ids = info["isect_ids"]
    assert bool((ids[1:] >= ids[:-1]).all())
    assert int(info["tiles_per_gauss"].sum().item()) == n
    off = info["isect_offsets"].flatten().long()
    assert bool((off[1:] >= off[:-1]).all()) and off[0].item() == 0 and off[-1].item() <= n
    tile_bits = (info["tile_width"] * info["tile_height"]).bit_length()
    cell_of_key = ((ids >> (32 + tile_bits)) * (info["tile_width"] * info["tile_height"]) + ((ids >> 32) & ((1 << tile_bits) - 1)))
    counts = torch.bincount(cell_of_key, minlength=off.numel())
    assert torch.equal(counts, torch.diff(torch.cat([off, torch.tensor([n], device=off.device)])))
    assert int(info["flatten_ids"].min()) >= 0 and int(info["flatten_ids"].max()) < info["gaussian_ids"].numel()
    # every (Gaussian, view) appears once per touched tile
    assert torch.equal(torch.bincount(info["flatten_ids"].long(), minlength=info["gaussian_ids"].numel()),
                       info["tiles_per_gauss"].long())
    # physical ranges
    assert torch.isfinite(r1).all() and float(a1.min()) >= 0.0 and float(a1.max()) <= 1.0 + 1e-6
    # the generic radix chain gives the same arrays at this size too
    gs.BINNING = "radix"
    try:
        with torch.no_grad():
            r3, a3, info3 = gs.rasterization(*args)
    finally:
        gs.BINNING = "fused"
    assert torch.equal(info3["isect_ids"], ids) and torch.equal(info3["flatten_ids"], info["flatten_ids"])
    assert torch.equal(info3["isect_offsets"], info["isect_offsets"]) and torch.equal(r3, r1)
    # permuting the Gaussians only changes the order of exact depth ties
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    with torch.no_grad():
        rp, ap, _ = gs.rasterization(sp["means"][perm], sp["quats"][perm], sp["scales"][perm], sp["opacities"][perm],
                                     sp["shN"][perm].contiguous(), viewmats, Ks, W, H)
    # (alpha compositing does not commute, so a pixel covered by two Gaussians of bit-equal depth may change: rare)
    assert float(((rp - r1).abs() > 1e-5).float().mean()) < 1e-3 and float(((ap - a1).abs() > 1e-5).float().mean()) < 1e-3
    assert float((rp - r1).abs().mean()) < 1e-6
    if N > 200_000:
        return
    # backward at full size.  (1) It is a linear map of the upstream gradient: g(w1 + w2) = g(w1) + g(w2).
    g = torch.Generator().manual_seed(2)
    w1 = torch.rand(C, H, W, 3, generator=g).to(cuda_device)
    w2 = torch.rand(C, H, W, 3, generator=g).to(cuda_device)

    def grads(wts):
        leaves = [sp[k].clone().requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "shN")]
        r, _, _ = gs.rasterization(*leaves, viewmats, Ks, W, H)
        (r * wts).sum().backward()
        return [x.grad for x in leaves]
    ga, gb, gab = grads(w1), grads(w2), grads(w1 + w2)
    ref = float(gab[0].abs().max())
    for x, y, z, name in zip(ga, gb, gab, ("means", "quats", "scales", "opacities", "shN")):
        # (isotropic splats: the quaternion gradient is pure cancellation noise, so it gets an absolute floor)
        scale = max(float(z.abs().max()), 1e-2 * ref)
        assert float((x + y - z).abs().max()) <= 2e-4 * scale, name
    assert float(gab[4][:, 4:].abs().max()) == 0.0                 # SH coefficients beyond degree 1 get no gradient
    # (2) directional derivative along a random direction of the means (loose: the alpha >= 1/255 cut, the radius
    # rounding and the tile assignment make the rendered loss piecewise smooth, which finite differences see)
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).to(cuda_device)
    analytic = float((ga[0] * d).sum())
    eps = 2e-4
    with torch.no_grad():
        lp = float((gs.rasterization(sp["means"] + eps * d, *args[1:])[0].double() * w1).sum())
        lm = float((gs.rasterization(sp["means"] - eps * d, *args[1:])[0].double() * w1).sum())
    numeric = (lp - lm) / (2 * eps)
    assert abs(analytic - numeric) <= 0.25 * max(abs(numeric), abs(analytic)) + 1.0, (analytic, numeric)


def test_train_plan_matches_unplanned(cuda_device):
    """TrainPlan (persistent buffers, capacity-sized intersection lists, no host sync) runs the same kernels on the
    same data as the exact-size path: identical bin indices and render, parameters equal up to the order of the
    fp32 gradient atomics; it re-sizes itself when the intersection count grows, and fails loudly on overflow."""
    import starst3r_b200 as st
    from starst3r_b200 import gs
    sp, viewmats, Ks, W, H = small_scene(n=800, C=3, seed=5)
    truth = torch.rand(3, H, W, 3, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    cams = gs.make_cams(viewmats.to(cuda_device), Ks.to(cuda_device))

    def fresh():
        p = {k: v.clone().to(cuda_device).contiguous() for k, v in sp.items()}
        return p, {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in p.items()}
    pa, sa = fresh()
    pb, sb = fresh()
    plan = gs.TrainPlan(800, 3, W, H, cuda_device)
    for i in range(6):
        la, fa = gs.train_step(pa, sa, truth, cams, W, H, i + 1)
        lb, fb = gs.train_step(pb, sb, truth, cams, W, H, i + 1, plan=plan)
        n = fa.n_isect
        assert fb.n_isect == n and plan.cap >= n
        if i == 0:      # same inputs: the integer pipeline and the render are bit-identical
            assert torch.equal(fa.keys[:n], fb.keys[:n]) and torch.equal(fa.vals[:n], fb.vals[:n])
            assert torch.equal(fa.offsets, fb.offsets) and torch.equal(fa.render, fb.render)
        assert abs(la.item() - lb.item()) <= 1e-5 * abs(la.item())
    for k in pa:
        assert torch.allclose(pa[k], pb[k], rtol=1e-4, atol=1e-5), k
    plan.poll(wait_all=True)
    # growth: inflate the Gaussians so the intersection count jumps past 80 % of the capacity
    cap0 = plan.cap
    pb["scales"] *= 1.6
    for i in range(3):
        try:
            gs.train_step(pb, sb, truth, cams, W, H, 7 + i, plan=plan)
            plan.poll(wait_all=True)
        except RuntimeError as e:     # a jump beyond the headroom is reported, never silently truncated
            assert "exceed the buffer capacity" in str(e)
    assert plan.cap > cap0
    lb, fb = gs.train_step(pb, sb, truth, cams, W, H, 10, plan=plan)
    plan.poll(wait_all=True)
    assert fb.n_isect <= plan.cap and plan.last_n_isect == fb.n_isect


def test_train_graph_replay_matches_eager(cuda_device, monkeypatch):
    """Steady-state iterations replayed as a CUDA graph (TrainPlan.graph_step) against the same iterations launched one
    by one: same losses and parameters up to the order of the fp32 gradient atomics; the graph survives a second set of
    truth images (second capture), a step-number jump, and a re-sized plan."""
    from starst3r_b200 import _lib, gs
    lib = _lib.load()
    sp, viewmats, Ks, W, H = small_scene(n=800, C=3, seed=8)
    g = torch.Generator().manual_seed(1)
    truths = [torch.rand(3, H, W, 3, generator=g).to(cuda_device) for _ in range(2)]
    cams = gs.make_cams(viewmats.to(cuda_device), Ks.to(cuda_device))

    def run(graph):
        monkeypatch.setattr(gs, "TRAIN_GRAPH", graph)
        p = {k: v.clone().to(cuda_device).contiguous() for k, v in sp.items()}
        st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in p.items()}
        plan = gs.TrainPlan(800, 3, W, H, cuda_device)
        losses, n0 = [], lib.st3r_launch_count()
        steps = list(range(1, 21)) + [40, 41, 42]
        for i, step in enumerate(steps):
            loss, fr = gs.train_step(p, st, truths[i & 1], cams, W, H, step, plan=plan, lr=1e-4)
            losses.append(loss)
        plan.poll(wait_all=True)
        assert fr.n_isect == plan.last_n_isect
        return p, st, [x.item() for x in losses], plan, lib.st3r_launch_count() - n0
    pe, se, le, plan_e, launches_e = run(False)
    pg, sg, lg, plan_g, launches_g = run(True)
    assert plan_e.graph_replays == 0 and plan_g.graph_replays == 23 - gs.TrainPlan.SYNC_FRAMES and len(plan_g._graphs) == 2
    assert int(plan_g.steps_done.item()) == 42
    assert launches_g >= launches_e                 # replays are counted (+ the counter kernel of the device-step Adam)
    for a, b in zip(le, lg):
        assert abs(a - b) <= 1e-5 * abs(a)
    for k in pe:
        assert torch.allclose(pe[k], pg[k], rtol=1e-4, atol=1e-5), k
        assert torch.allclose(se[k][1], sg[k][1], rtol=1e-3, atol=1e-12), k


def test_train_plan_watches_the_intersection_count(cuda_device):
    """The plan's first frames, and the frames after any growth of the count above GROWTH_WATCH, read the count back
    before binning (nothing truncated can reach the optimiser); quiet stretches run without host synchronisation."""
    from starst3r_b200 import gs
    sp, viewmats, Ks, W, H = small_scene(n=800, C=3, seed=6)
    truth = torch.rand(3, H, W, 3, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    cams = gs.make_cams(viewmats.to(cuda_device), Ks.to(cuda_device))
    p = {k: v.clone().to(cuda_device).contiguous() for k, v in sp.items()}
    st = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in p.items()}
    plan = gs.TrainPlan(800, 3, W, H, cuda_device)
    assert plan.sync_mode()
    for i in range(gs.TrainPlan.SYNC_FRAMES + 2):
        gs.train_step(p, st, truth, cams, W, H, i + 1, plan=plan, lr=1e-5)
    plan.poll(wait_all=True)
    assert not plan.sync_mode()                       # quiet: asynchronous counts from here on
    n0 = plan.last_n_isect
    p["scales"] *= 1.25                               # ~1.5x the intersections: seen one frame late, inside the headroom
    gs.train_step(p, st, truth, cams, W, H, 20, plan=plan, lr=1e-5)
    plan.poll(wait_all=True)
    assert plan.last_n_isect > (1 + gs.TrainPlan.GROWTH_WATCH) * n0 and plan.sync_mode()
    p["scales"] *= 2.0                                # far beyond the headroom, but the frame is watched: buffers grow first
    _, fr = gs.train_step(p, st, truth, cams, W, H, 21, plan=plan, lr=1e-5)
    plan.poll(wait_all=True)
    assert fr.n_isect == plan.last_n_isect <= plan.cap and fr.n_isect > 1.5 * n0


def test_scene_api_run_3dgs_optim(cuda_device):
    """Scene.init_3dgs / run_3dgs_optim / render_3dgs_original with the reference's call pattern (main.py:77-88):
    the loss decreases and the API objects have the reference's shape."""
    import starst3r_b200 as st
    W, H, C = 64, 48, 3
    viewmats, Ks = synth.look_at_cameras(C, W, H)
    sp = synth.random_splats(3000, seed=3, scale_mode="rand")
    sp["scales"] = sp["scales"] * 6
    d = to(cuda_device, sp)
    with torch.no_grad():
        target, _, _ = st.gs.rasterization(d["means"], d["quats"], d["scales"], d["opacities"], d["shN"],
                                           viewmats.to(cuda_device), Ks.to(cuda_device), W, H)
    scene = st.Scene(device=cuda_device)
    scene.imgs = [t.clamp(0, 1).cpu().numpy() for t in target]
    scene.c2w = torch.linalg.inv(viewmats).to(cuda_device)
    scene.intrinsics = Ks.to(cuda_device)
    g = torch.Generator().manual_seed(0)
    scene.dense_pts = [(sp["means"] + 0.01 * torch.randn(3000, 3, generator=g)).to(cuda_device)]
    scene.dense_cols = [torch.rand(3000, 3, generator=g)]
    scene.init_3dgs(init_scale=2e-2)
    assert set(scene.gaussians) == {"means", "scales", "quats", "opacities", "sh0", "shN"}
    assert all(isinstance(v, torch.nn.Parameter) for v in scene.gaussians.values())
    assert scene.gaussians["shN"].shape == (3000, 24, 3) and scene.gaussians["quats"][0].tolist() == [1, 0, 0, 0]
    losses = scene.run_3dgs_optim(30)
    assert len(losses) == 30 and all(isinstance(x, float) for x in losses)
    assert losses[-1] < losses[0]
    losses2 = scene.run_3dgs_optim(5, enable_pruning=True)
    assert losses2[-1] <= losses[-1] * 1.05
    assert scene.optimizers["means"].state[scene.gaussians["means"]]["step"].item() == 35
    assert scene.gaussians["sh0"].grad is None
    r, a, info = scene.render_3dgs_original(W, H)
    assert r.shape == (C, H, W, 3) and a.shape == (C, H, W, 1) and "isect_offsets" in info
    # fly-through between two training cameras: end points reproduce the per-camera renders
    path = st.gs.render_3dgs_path(scene, scene.c2w[0], scene.c2w[1], 5, scene.intrinsics[0], W, H, chunk=2)
    assert path.shape == (5, H, W, 3) and torch.isfinite(path).all()
    assert torch.allclose(path[0], r[0], atol=2e-3) and torch.allclose(path[-1], r[1], atol=2e-3)
