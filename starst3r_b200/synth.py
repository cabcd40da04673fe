"""Deterministic synthetic inputs for tests and bench.py: ring-of-cameras scenes, random splats,
descriptor maps.  Host-side helper only (no hot-path code)."""
import math

import torch


def look_at_cameras(n_views, width, height, radius=3.0, elev=0.35, focal_mult=1.2, device="cpu"):
    """N pinhole cameras on a ring looking at the origin.  Returns (viewmats [C,4,4] world->cam, Ks [C,3,3])."""
    viewmats, Ks = [], []
    f = focal_mult * max(width, height)
    for i in range(n_views):
        th = 2 * math.pi * i / n_views
        eye = torch.tensor([radius * math.cos(th), elev * radius * math.sin(2 * th + 0.3), radius * math.sin(th)])
        fwd = -eye / eye.norm()
        up = torch.tensor([0.0, 1.0, 0.0])
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        down = torch.linalg.cross(fwd, right)
        R = torch.stack([right, down, fwd])          # rows: camera x, y, z axes in world coords
        t = -R @ eye
        V = torch.eye(4)
        V[:3, :3] = R
        V[:3, 3] = t
        viewmats.append(V)
        Ks.append(torch.tensor([[f, 0, width / 2], [0, f, height / 2], [0, 0, 1.0]]))
    return torch.stack(viewmats).to(device), torch.stack(Ks).to(device)


def random_splats(n, seed=0, scale_mode="init", extent=1.0, device="cpu"):
    """Splat parameters in the reference's parametrisation (raw scales / opacities, wxyz quats, shN [N,24,3])."""
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(n, 3, generator=g) * 2 - 1) * extent
    if scale_mode == "init":
        scales = torch.full((n, 3), 3e-3)
    else:
        scales = torch.exp(torch.randn(n, 3, generator=g) * 0.5 - 4.0)
    quats = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1)
    opac = torch.rand(n, generator=g) * 0.9 + 0.1
    shN = torch.randn(n, 24, 3, generator=g) * 0.3
    return {k: v.to(device) for k, v in dict(means=means, scales=scales, quats=quats, opacities=opac, shN=shN).items()}


def descriptor_pair(H, W, d=24, noise=0.3, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    A = torch.nn.functional.normalize(torch.randn(H, W, d, generator=g), dim=-1)
    B = torch.nn.functional.normalize(A + noise * torch.randn(H, W, d, generator=g), dim=-1)
    return A.to(device), B.to(device)
