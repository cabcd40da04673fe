"""Deterministic synthetic inputs for tests and bench.py: ring-of-cameras scenes, random splats,
descriptor maps.  Host-side helper only (no hot-path code)."""
import math

import torch


def look_at_cameras(n_views, width, height, radius=3.0, elev=0.35, focal_mult=1.2, device="cpu", arc_deg=360.0):
    """N pinhole cameras on a ring looking at the origin.  Returns (viewmats [C,4,4] world->cam, Ks [C,3,3])."""
    viewmats, Ks = [], []
    f = focal_mult * max(width, height)
    for i in range(n_views):
        th = math.radians(arc_deg) * i / n_views
        eye = torch.tensor([radius * math.cos(th), -elev * radius * (1.0 + 0.5 * math.sin(2 * th + 0.3)), radius * math.sin(th)])  # y points down: cameras sit above the scene
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


# ------------------------------------------------------------------------------------------------------------
# Scene-consistent stand-in for the MASt3R network: the network itself is outside the hot
# path and its checkpoint is not available offline, so tests / bench.py feed the MATCH + ALIGN path with
# point maps, confidences and descriptors ray-cast from a known scene (sphere + ground plane inside a large
# background sphere) seen by known cameras.  Confidences exceed matching_conf_thr=5 on the objects, so the
# optimiser takes the matching-loss branch (reconstruct.py:283-290) like it does with the real checkpoint.
# ------------------------------------------------------------------------------------------------------------
class SyntheticMast3r:
    """`symmetric_inference(img1, img2)` with the output contract of sparse_ga.py:571-592:
    (res11, res21, res22, res12), each {pts3d [1,H,W,3], conf [1,H,W], desc [1,H,W,24], desc_conf [1,H,W]}."""

    def __init__(self, n_views, width, height, seed=0, desc_noise=0.05, pts_noise=0.0, device="cpu", low_conf=False,
                 arc_deg=360.0):
        self.W, self.H, self.device = width, height, torch.device(device)
        self.viewmats, self.Ks = look_at_cameras(n_views, width, height, device="cpu", arc_deg=arc_deg)
        self.c2w = torch.linalg.inv(self.viewmats)
        g = torch.Generator().manual_seed(seed)
        self.freq = torch.randn(24, 3, generator=g) * 2.5
        self.phase = torch.rand(24, generator=g) * 6.283
        self.desc_noise, self.pts_noise, self.seed, self.low_conf = desc_noise, pts_noise, seed, low_conf
        self._world = [self._raycast(i) for i in range(n_views)]      # (world points [H,W,3], object mask [H,W])

    @staticmethod
    def _hit(o, d):
        """First intersection of the rays o + t d (d: [..., 3]) with the scene: returns (t, is_object)."""
        shape = d.shape[:-1]
        t_best = torch.full(shape, float("inf"))
        obj = torch.zeros(shape, dtype=torch.bool)

        def sphere(center, radius, inside=False):
            oc = o - center
            a = (d * d).sum(-1)
            b = 2 * (d * oc).sum(-1)
            c = (oc * oc).sum() - radius * radius
            disc = b * b - 4 * a * c
            sq = torch.sqrt(disc.clamp_min(0))
            t = (-b + sq) / (2 * a) if inside else (-b - sq) / (2 * a)
            return torch.where((disc > 0) & (t > 1e-3), t, torch.full_like(t, float("inf")))
        t_s = sphere(torch.zeros(3), 0.8)
        t_p = (0.8 - o[1]) / d[..., 1]                               # ground plane y = 0.8 (y points down)
        hit = o + t_p[..., None] * d
        t_p = torch.where((t_p > 1e-3) & (hit[..., 0].abs() < 2.5) & (hit[..., 2].abs() < 2.5), t_p,
                          torch.full_like(t_p, float("inf")))
        t_b = sphere(torch.zeros(3), 8.0, inside=True)
        for t, is_obj in ((t_s, True), (t_p, True), (t_b, False)):
            closer = t < t_best
            t_best = torch.where(closer, t, t_best)
            obj = torch.where(closer, torch.full_like(obj, is_obj), obj)
        return t_best, obj

    def _raycast(self, i):
        H, W = self.H, self.W
        K, c2w = self.Ks[i], self.c2w[i]
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                                indexing="ij")
        d = torch.stack([(xs - K[0, 2]) / K[0, 0], (ys - K[1, 2]) / K[1, 1], torch.ones_like(xs)], -1)
        d = d @ c2w[:3, :3].T
        o = c2w[:3, 3]
        t, obj = self._hit(o, d)
        return o + t[..., None] * d, obj

    def _desc(self, Xw, salt):
        g = torch.Generator().manual_seed(self.seed * 7919 + salt)
        f = torch.cos(Xw @ self.freq.T + self.phase)
        f = f + self.desc_noise * torch.randn(f.shape, generator=g)
        return torch.nn.functional.normalize(f, dim=-1)

    def _visible(self, src, other):
        """Is the surface point seen by each pixel of `src` also seen by camera `other` (inside its frame and the
        first thing its ray hits)?"""
        Xw, _ = self._world[src]
        V, K = self.viewmats[other], self.Ks[other]
        X = Xw @ V[:3, :3].T + V[:3, 3]
        z = X[..., 2].clamp_min(1e-6)
        u = K[0, 0] * X[..., 0] / z + K[0, 2]
        v = K[1, 1] * X[..., 1] / z + K[1, 2]
        inside = (X[..., 2] > 0) & (u >= 0) & (u <= self.W - 1) & (v >= 0) & (v <= self.H - 1)
        o = self.c2w[other][:3, 3]
        t, _ = self._hit(o, Xw - o)                  # target sits at t = 1 along this ray
        return inside & (t > 0.98)

    def _res(self, src, frame, salt, other):
        """Image `src`'s pixels: their 3-D points expressed in camera `frame`'s coordinates; confidences are high
        only where the surface is an object AND co-visible with image `other` (1 + 9 * visibility)."""
        Xw, obj = self._world[src]
        V = self.viewmats[frame]
        X = Xw @ V[:3, :3].T + V[:3, 3]
        if self.pts_noise:
            g = torch.Generator().manual_seed(self.seed * 104729 + salt)
            X = X * (1 + self.pts_noise * torch.randn(X.shape[:2] + (1,), generator=g))
        good = obj & self._visible(src, other)
        conf = torch.where(good, torch.tensor(6.0 + 1.5 * (src % 4)), torch.tensor(1.5))   # per-view confidence level
        if self.low_conf:
            conf = conf.clamp(max=3.0)           # forces the loss_dust3r fallback (max conf <= 5)
        dev = self.device
        return {"pts3d": X[None].to(dev), "conf": conf[None].to(dev), "desc": self._desc(Xw, salt)[None].to(dev),
                "desc_conf": conf[None].clone().to(dev)}

    def symmetric_inference(self, img1, img2, device=None):
        i, j = int(img1["idx"]), int(img2["idx"])
        n = len(self._world)
        return (self._res(i, i, (i * n + j) * 4, j), self._res(j, i, (i * n + j) * 4 + 1, i),
                self._res(j, j, (i * n + j) * 4 + 2, i), self._res(i, j, (i * n + j) * 4 + 3, j))

    def symmetric_inference_batch(self, imgs1, imgs2, device=None):
        """The batched entry point reconstruct.symmetric_inference_batch looks for; the synthetic
        predictions are procedural, so a batch is just the pairs in turn."""
        return [self.symmetric_inference(a, b) for a, b in zip(imgs1, imgs2)]

    def images(self):
        """[-1, 1] normalised (3, H, W) images (shading from the descriptor field; only colours downstream)."""
        out = []
        for Xw, obj in self._world:
            rgb = 0.5 + 0.5 * torch.cos(Xw * 3.0)
            out.append((rgb.permute(2, 0, 1) * 2 - 1).contiguous())
        return out
