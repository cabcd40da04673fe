"""CPU: (1) known-answer tests that pin the RASTER oracle (parity unpinned w.r.t. gsplat, see oracle/gs_oracle.py);
(2) the product's per-Gaussian math header, compiled for the host, against the oracle: projection outputs
bit-exact, analytic backward vs autograd."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import gs_oracle as go
from starst3r_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = tmp_path_factory.mktemp("host") / "libgs_math_host.so"
    src = os.path.join(ROOT, "tests", "host", "gs_math_host.cpp")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


def cams_array(viewmats, Ks):
    C = viewmats.shape[0]
    pos = torch.linalg.inv(viewmats)[:, :3, 3]
    cams = torch.cat([viewmats[:, :3, :3].reshape(C, 9), viewmats[:, :3, 3], Ks[:, 0, 0:1], Ks[:, 1, 1:2], Ks[:, 0, 2:3],
                      Ks[:, 1, 2:3], pos], dim=1).contiguous().float()
    assert cams.shape[1] == 19
    return cams


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def scene(n=400, C=3, W=64, H=48, seed=0, scale_mode="rand"):
    sp = synth.random_splats(n, seed=seed, scale_mode=scale_mode)
    sp["scales"] = sp["scales"] * 8      # a few pixels wide at this resolution
    viewmats, Ks = synth.look_at_cameras(C, W, H)
    return sp, viewmats, Ks


def test_oracle_single_gaussian_closed_form():
    """Isotropic Gaussian on the optical axis: mean2d = principal point, cov2d = (f s / z)^2 I + 0.3 I."""
    W = H = 64
    f, z, s = 80.0, 4.0, 0.05
    viewmats = torch.eye(4)[None]
    Ks = torch.tensor([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]]])
    means = torch.tensor([[0.0, 0.0, z]])
    quats = torch.tensor([[1.0, 0, 0, 0]])
    scales = torch.full((1, 3), s)
    radii, m2, depth, conic = go.project(means, quats, scales, viewmats, Ks, W, H)
    var = (f * s / z) ** 2 + 0.3
    assert torch.allclose(m2[0, 0], torch.tensor([W / 2, H / 2]))
    assert depth[0, 0].item() == z
    assert torch.allclose(conic[0, 0], torch.tensor([1 / var, 0.0, 1 / var]), rtol=1e-6)
    assert radii[0, 0].item() == math.ceil(3 * math.sqrt(var))
    # colour / alpha at the pixel whose centre is (32.5, 32.5)
    shN = torch.zeros(1, 24, 3)
    shN[0, 0] = torch.tensor([1.0, 0.5, -2.0])
    op = torch.tensor([0.8])
    render, alpha, info = go.rasterization(means, quats, scales, op, shN, viewmats, Ks, W, H)
    a = 0.8 * math.exp(-0.5 * (0.5 ** 2 + 0.5 ** 2) / var)
    col = np.maximum(0.2820947917738781 * np.array([1.0, 0.5, -2.0]) + 0.5, 0)
    assert np.allclose(render[0, 32, 32].numpy(), a * col, rtol=1e-5)
    assert np.isclose(alpha[0, 32, 32, 0].item(), a, rtol=1e-5)
    # straddles the 4 tiles around (32, 32)
    assert info["tiles_per_gauss"].tolist() == [4]
    assert info["isect_offsets"].shape == (1, 4, 4)
    tiles = (info["isect_ids"] >> 32).tolist()
    assert tiles == [1 * 4 + 1, 1 * 4 + 2, 2 * 4 + 1, 2 * 4 + 2]
    assert (info["isect_ids"] & 0xffffffff).tolist() == [np.float32(z).view(np.int32)] * 4


def test_oracle_two_gaussians_depth_order_and_termination():
    W = H = 32
    f = 40.0
    viewmats = torch.eye(4)[None]
    Ks = torch.tensor([[[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]]])
    means = torch.tensor([[0.0, 0.0, 5.0], [0.0, 0.0, 2.0]])     # second one is nearer
    quats = torch.tensor([[1.0, 0, 0, 0]] * 2)
    scales = torch.full((2, 3), 0.3)
    shN = torch.zeros(2, 24, 3)
    shN[0, 0] = (1.0 - 0.5) / go.SH_C0      # far: rgb = 1
    shN[1, 0] = (0.25 - 0.5) / go.SH_C0     # near: rgb = 0.25
    op = torch.tensor([0.9, 0.5])
    render, alpha, info = go.rasterization(means, quats, scales, op, shN, viewmats, Ks, W, H)
    # sorted front to back inside each tile: flatten id 1 (z=2) precedes 0 (z=5)
    assert info["flatten_ids"][:2].tolist() == [1, 0]
    var_n, var_f = (f * 0.3 / 2) ** 2 + 0.3, (f * 0.3 / 5) ** 2 + 0.3
    d2 = 0.5
    an, af = 0.5 * math.exp(-0.5 * d2 / var_n), 0.9 * math.exp(-0.5 * d2 / var_f)
    exp_rgb = 0.25 * an + 1.0 * af * (1 - an)
    assert np.allclose(render[0, 16, 16].numpy(), exp_rgb, rtol=1e-5)
    assert np.isclose(alpha[0, 16, 16, 0].item(), 1 - (1 - an) * (1 - af), rtol=1e-5)
    # opacity > 0.999 clamps alpha at 0.999; a stack of them exhausts the pixel (T <= 1e-4) after 1 blend
    means = torch.tensor([[0.0, 0.0, 2.0 + 0.1 * i] for i in range(4)])
    quats = torch.tensor([[1.0, 0, 0, 0]] * 4)
    scales = torch.full((4, 3), 0.5)
    shN = torch.zeros(4, 24, 3)
    render, alpha, info = go.rasterization(means, quats, scales, torch.full((4,), 5.0), shN, viewmats, Ks, W, H)
    # T after first = 1e-3, second would give 1e-6 <= 1e-4 -> second is NOT blended
    assert np.isclose(alpha[0, 16, 16, 0].item(), 0.999, rtol=1e-6)
    assert info["last_ids"][0, 16, 16] == info["isect_offsets"][0, 1, 1]


def test_oracle_ssim_properties():
    g = torch.Generator().manual_seed(0)
    a = torch.rand(1, 3, 40, 36, generator=g)
    assert abs(go.ssim(a, a).item() - 1.0) < 1e-6
    b = torch.rand(1, 3, 40, 36, generator=g)
    assert abs(go.ssim(a, b).item() - go.ssim(b, a).item()) < 1e-7 and go.ssim(a, b).item() < 0.2
    assert abs(go.gaussian_window().sum().item() - 1) < 1e-6


def test_host_projection_bit_exact_vs_oracle(hostlib):
    sp, viewmats, Ks = scene()
    N, C, W, H = sp["means"].shape[0], viewmats.shape[0], 64, 48
    radii, m2, depth, conic = go.project(sp["means"], sp["quats"], sp["scales"], viewmats, Ks, W, H)
    rgb = go.sh_colors(sp["means"], torch.linalg.inv(viewmats)[:, :3, 3], sp["shN"])
    cams = cams_array(viewmats, Ks)
    r = torch.zeros(C * N, dtype=torch.int32)
    geom = torch.zeros(C * N, 6)
    col = torch.zeros(C * N, 3)
    shN = sp["shN"].contiguous()
    hostlib.host_project(P(sp["means"]), P(sp["quats"]), P(sp["scales"]), P(shN), 72, P(cams), N, C,
                         ctypes.c_float(W), ctypes.c_float(H), P(r), P(geom), P(col))
    assert (radii > 0).sum() > 50
    assert torch.equal(r.reshape(C, N), radii)
    vis = (radii > 0).reshape(-1)
    assert torch.equal(geom[vis, 0:2], m2.reshape(-1, 2)[vis])          # bit-exact
    assert torch.equal(geom[vis, 2], depth.reshape(-1)[vis])
    assert torch.equal(geom[vis, 3:6], conic.reshape(-1, 3)[vis])
    assert torch.allclose(col[vis], rgb.reshape(-1, 3)[vis], atol=1e-6)


def test_host_projection_backward_vs_autograd(hostlib):
    sp, viewmats, Ks = scene(n=300, seed=3)
    N, C, W, H = sp["means"].shape[0], viewmats.shape[0], 64, 48
    means, quats, scales, shN = [sp[k].clone().double().requires_grad_(True) for k in ("means", "quats", "scales", "shN")]
    radii, m2, depth, conic = go.project(means, quats, scales, viewmats.double(), Ks.double(), W, H)
    rgb = go.sh_colors(means, torch.linalg.inv(viewmats.double())[:, :3, 3], shN)
    g = torch.Generator().manual_seed(1)
    vis = (radii > 0)
    v_m2 = torch.randn(C, N, 2, generator=g) * vis[..., None]
    v_conic = torch.randn(C, N, 3, generator=g) * vis[..., None]
    v_rgb = torch.randn(C, N, 3, generator=g) * vis[..., None]
    loss = (m2 * v_m2).sum() + (conic * v_conic).sum() + (rgb * v_rgb).sum()
    loss.backward()
    cams = cams_array(viewmats, Ks)
    radii_f = go.project(sp["means"], sp["quats"], sp["scales"], viewmats, Ks, W, H)[0]
    assert torch.equal(radii_f, radii)
    vm, vq, vs, vsh = torch.zeros(N, 3), torch.zeros(N, 4), torch.zeros(N, 3), torch.zeros(N, 12)
    shN32 = sp["shN"].contiguous()
    r32 = radii_f.reshape(-1).contiguous()
    hostlib.host_project_bwd(P(sp["means"]), P(sp["quats"]), P(sp["scales"]), P(shN32), 72, P(cams), N, C,
                             ctypes.c_float(W), ctypes.c_float(H), P(r32), P(v_m2.contiguous()),
                             P(v_conic.contiguous()), P(v_rgb.contiguous()), P(vm), P(vq), P(vs), P(vsh))

    def close(a, b, name):
        b = b.float()
        err = (a - b).abs().max().item()
        scale = b.abs().max().item() + 1e-12
        assert err / scale < 2e-3, (name, err, scale)
        # elementwise on well-conditioned entries
        big = b.abs() > 1e-3 * scale
        assert ((a - b).abs()[big] / b.abs()[big]).median().item() < 1e-4, name
    close(vm, means.grad, "means")
    close(vq, quats.grad, "quats")
    close(vs, scales.grad, "scales")
    close(vsh, shN.grad[:, :4].reshape(N, 12), "sh")
    assert shN.grad[:, 4:].abs().max().item() == 0.0      # coefficients 4.. never get a gradient
