"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — CPU restatement of the ALIGN optimiser.

Follows starster/reconstruct.py:116-457 (sparse_scene_optimizer_slam: make_K_cam_depth :209-261, loss_dust3r
:311-323, loss_3d :325-353, loss_2d :355-369, optimize_loop :371-406) and mast3r/cloud_opt/sparse_ga.py:464-501,
977-981 (make_pts3d, proj3d, reproj2d), cloud_opt/utils/losses.py:19-28, schedules.py:15-17, in plain PyTorch
with autograd.  Parity pinned: tests/test_oracle_golden.py::test_align_* compare it with fixtures produced by the
unmodified reference (oracle/gen_golden_align.py -> tests/golden/align_*.pt).
Inputs are the reference's condense_data structures in "plain" form (tuples instead of PairOfSlices, ("slice", a, b)
instead of slice objects)."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _sl(s):
    if isinstance(s, slice):
        return s
    return slice(s[1], s[2])


def unitquat_to_rotmat_xyzw(q):
    """roma.unitquat_to_rotmat."""
    x, y, z, w = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(q.shape[:-1] + (3, 3))


def gamma_loss(gamma):
    """cloud_opt/utils/losses.py:19-28."""
    if gamma == 1:
        return lambda x, y: torch.linalg.norm(x - y, dim=-1)
    offset = (1 / gamma) ** (1 / (gamma - 1))
    return lambda x, y: (torch.linalg.norm(x - y, dim=-1) + offset) ** gamma - offset ** gamma


def cosine_schedule(alpha, lr_base, lr_end=0):
    return lr_end + (lr_base - lr_end) * (1 + np.cos(alpha * np.pi)) / 2


class Problem:
    """Everything the optimiser closes over (reconstruct.py:141-309)."""

    def __init__(self, inp, matching_conf_thr=5.0):
        self.imgs = list(inp["imgs"])
        N = len(self.imgs)
        self.imsizes = inp["imsizes"].float()
        self.base_focals = inp["base_focals"].clone().float()
        self.pps0 = inp["pps"].float() / self.imsizes                                   # :170
        core = [c.clone().float() for c in inp["core_depth"]]
        self.median = torch.stack([c.median() for c in core])                           # :176
        self.core = [(c / m).ravel() for c, m in zip(core, self.median)]                # :177
        diags = self.imsizes.norm(dim=1)
        self.min_focals, self.max_focals = 0.25 * diags, 10 * diags                     # :203-205
        self.root, self.edges = inp["mst"][0], [tuple(e) for e in inp["mst"][1]]
        self.anchors = {int(k): (v[0].float(), v[1].long(), v[2].float()) for k, v in inp["anchors"].items()}
        _, _, slices = inp["corres"]
        ok = {(s[0], s[4]): bool(s[8].max() > matching_conf_thr) for s in slices}       # :283-286
        self.loss3d_slices = [s for s in slices if ok[s[0], s[4]]]
        self.dust3r_slices = [s for s in slices if not ok[s[0], s[4]]]
        self.cleaned2d = []                                                             # :291-309
        for img1, pix1, confs, _, sls in inp["corres2d"]:
            cur, px, cf, keep = 0, [], [], []
            for img2, sl2 in sls:
                sl2 = _sl(sl2)
                n = sl2.stop - sl2.start
                if ok[img1, img2]:
                    px.append(pix1[cur:cur + n])
                    cf.append(confs[cur:cur + n])
                    keep.append((img2, sl2))
                cur += n
            if px:
                self.cleaned2d.append((img1, torch.cat(px).float(), torch.cat(cf).float(), keep))
        self.preds_21 = inp["preds_21"]
        self.N = N

    def init_params(self):
        N = self.N
        return dict(pps=self.pps0.clone(), log_focals=self.base_focals.log().clone(),
                    quats=torch.tensor([[0.0, 0, 0, 1]]).repeat(N, 1), trans=torch.zeros(N, 3),
                    log_sizes=torch.zeros(N))

    def make_K_cam_depth(self, p):                                                       # :209-261
        N = self.N
        focals = p["log_focals"].exp().clip(min=self.min_focals, max=self.max_focals)
        K = torch.eye(3).repeat(N, 1, 1)
        K[:, 0, 0] = K[:, 1, 1] = focals
        K[:, 0:2, 2] = p["pps"] * self.imsizes
        sizes = p["log_sizes"].exp()
        g = 1 / sizes.min()
        z_cam = sizes * self.median * focals / self.base_focals
        rel = torch.eye(4).repeat(N, 1, 1)
        rel[:, :3, :3] = unitquat_to_rotmat_xyzw(F.normalize(p["quats"], dim=1))
        rel[:, :3, 3] = p["trans"]
        T = [None] * N
        T[self.root] = rel[self.root]
        for i, j in self.edges:
            T[j] = T[i] @ rel[j]
        T = torch.stack(T)
        off = z_cam[:, None] * torch.cat((self.imsizes / focals[:, None] * (0.5 - p["pps"]), torch.ones(N, 1)), -1)
        new_t = g * (T[:, :3, 3:4] - T[:, :3, :3] @ off[:, :, None])
        cam2w = torch.cat((torch.cat((T[:, :3, :3], new_t), 2), torch.tensor([0.0, 0, 0, 1]).expand(N, 1, 4)), 1)
        depth = [g * (z_cam[i] + (self.core[i] - 1) * (self.median[i] * sizes[i])) for i in range(N)]
        return K, cam2w, depth

    def make_pts3d(self, K, cam2w, depth):                                               # sparse_ga.py:475-501
        out = []
        for i in range(self.N):
            pix, idx, off = self.anchors[i]
            f = K[i, 0, 0]
            o = 1 + (off - 1) * (self.base_focals[i] / f)
            z = depth[i][idx] * o
            pc = z[:, None] * torch.stack([(pix[:, 0] - K[i, 0, 2]) / f, (pix[:, 1] - K[i, 1, 2]) / f,
                                           torch.ones_like(z)], -1)
            out.append(pc @ cam2w[i, :3, :3].T + cam2w[i, :3, 3])
        return out

    def loss_3d(self, K, cam2w, pts3d, pix_loss):                                        # :325-353
        if not self.loss3d_slices:
            return torch.zeros(())
        a = torch.cat([pts3d[s[0]][_sl(s[1])] for s in self.loss3d_slices])
        b = torch.cat([pts3d[s[4]][_sl(s[5])] for s in self.loss3d_slices])
        c = torch.cat([s[8] for s in self.loss3d_slices]).float()
        return c @ pix_loss(a, b) / c.sum()

    def loss_2d(self, K, cam2w, pts3d, pix_loss):                                        # :355-369
        w2c = torch.linalg.inv(cam2w)
        proj = K @ w2c[:, :3]
        loss, npix = 0.0, 0.0
        for img1, pix1, confs, keep in self.cleaned2d:
            P = torch.cat([pts3d[img2][sl2] for img2, sl2 in keep])
            r = P @ proj[img1][:3, :3].T + proj[img1][:3, 3]                            # sparse_ga.py:977-981
            uv = (r[:, 0:2] / r[:, 2:3].clip(min=1e-3)).clip(min=-1000, max=2000)
            loss = loss + confs @ pix_loss(pix1, uv)
            npix = npix + confs.sum()
        return loss / npix if npix != 0 else torch.zeros(())

    def loss_dust3r(self, cam2w, pts3d, pix_loss):                                       # :311-323
        loss, cf = 0.0, 0.0
        for s in self.dust3r_slices:
            tgt, tc = self.preds_21[self.imgs[s[4]]][self.imgs[s[0]]]
            tgt = tgt.float() @ cam2w[s[4], :3, :3].T + cam2w[s[4], :3, 3]
            cf = cf + tc.sum()
            loss = loss + tc.float() @ pix_loss(pts3d[s[0]], tgt)
        return loss / cf if cf != 0 else torch.zeros(())

    def total_loss(self, p, mode, gamma, dust3r_w=0.01):
        K, cam2w, depth = self.make_K_cam_depth(p)
        pts3d = self.make_pts3d(K, cam2w, depth)
        main = (self.loss_3d if mode == 0 else self.loss_2d)(K, cam2w, pts3d, gamma_loss(gamma))
        return main + dust3r_w * self.loss_dust3r(cam2w, pts3d, gamma_loss(1.1)), (K, cam2w, depth, pts3d)

    def optimize(self, p, mode, lr_base, niter, gamma, train):                           # :371-406
        params = [p[k].requires_grad_(k in train) for k in ("pps", "log_focals", "quats", "trans", "log_sizes")]
        opt = torch.optim.Adam([q for q in params if q.requires_grad], lr=1, weight_decay=0, betas=(0.9, 0.9))
        losses, res = [], None
        for it in range(niter or 1):
            loss, state = self.total_loss(p, mode, gamma)
            res = state
            if niter == 0:
                break
            for gr in opt.param_groups:
                gr["lr"] = float(cosine_schedule(it / niter, lr_base, 0))
            opt.zero_grad()
            loss.backward()
            opt.step()
            with torch.no_grad():
                p["quats"] /= p["quats"].norm(dim=1, keepdim=True)
            losses.append(float(loss))
        K, cam2w, depth, pts3d = res
        return dict(intrinsics=K.detach(), cam2w=cam2w.detach(), depthmaps=[d.detach() for d in depth],
                    pts3d=[x.detach() for x in pts3d]), losses


def run(inp, lr1=0.07, niter1=500, lr2=0.014, niter2=200, opt_pp=True, thr=5.0):
    """sparse_scene_optimizer(...) on fixture inputs; returns (res_coarse, res_fine, params, losses)."""
    pb = Problem(inp, thr)
    p = pb.init_params()
    res_c, l1 = pb.optimize(p, 0, lr1, niter1, 1.1, ("quats", "trans", "log_sizes"))
    res_f, l2 = None, []
    if niter2:
        train = ("quats", "trans", "log_sizes", "log_focals") + (("pps",) if opt_pp else ())
        res_f, l2 = pb.optimize(p, 1, lr2, niter2, 0.4, train)
    return res_c, res_f, p, (l1, l2)


# ----------------------------------------------------------------------------- canonical view / dense points
def canonical_view(ptmaps11, confs11, subsample=8):
    """sparse_ga.py:817-855, mode='avg-angle'."""
    c = confs11.unsqueeze(-1) - 0.999
    canon = (c * ptmaps11).sum(0) / c.sum(0)
    depth = ptmaps11[..., 2].unsqueeze(1)
    S = slice(subsample // 2, None, subsample)
    center = depth[:, :, S, S].clip(min=torch.finfo(depth.dtype).eps)
    sd = F.pixel_unshuffle(depth, subsample)
    sc = F.pixel_unshuffle(c[:, None, :, :, 0], subsample)
    xy = ptmaps11[..., 0:2].permute(0, 3, 1, 2)
    sxy = F.pixel_unshuffle(xy, subsample)
    B, _, H, W = sxy.shape
    rad = (sxy.view(B, 2, -1, H, W) - xy[:, :, None, S, S]).norm(dim=1).clip(min=1e-8)
    ang = torch.arctan((sd - center) / rad)
    avg = (sc * ang).sum(0) / sc.sum(0)
    sdepth = rad.mean(0) * torch.tan(avg)
    canon2 = F.pixel_shuffle((1 + sdepth / canon[S, S, 2]).unsqueeze(0), subsample).squeeze()
    conf = (c.square().sum(0) / c.sum(0)).squeeze()
    return canon, canon2, conf


def estimate_focal_weiszfeld(canon, min_focal=0.5, max_focal=3.5):
    """dust3r/post_process.py:36-58 for one [H,W,3] point map with pp at the image centre."""
    H, W = canon.shape[:2]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    pix = torch.stack([xs - W / 2, ys - H / 2], -1).reshape(-1, 2)
    p = canon.reshape(-1, 3)
    xyz = (p[:, :2] / p[:, 2:3]).nan_to_num(posinf=0, neginf=0)
    dxp, dxx = (xyz * pix).sum(-1), xyz.square().sum(-1)
    f = dxp.mean() / dxx.mean()
    for _ in range(10):
        w = (pix - f * xyz).norm(dim=-1).clip(min=1e-8).reciprocal()
        f = (w * dxp).mean() / (w * dxx).mean()
    fb = max(H, W) / (2 * math.tan(math.radians(60) / 2))
    return f.clip(min=min_focal * fb, max=max_focal * fb)


def clean_pointcloud(confs, K, w2c, depthmaps, pts3d, tol=0.001):
    """dust3r/cloud_opt/base_opt.py:369-405."""
    res = [c.clone() for c in confs]
    pts = [p.view(*c.shape, 3) for p, c in zip(pts3d, confs)]
    dm = [d.view(*c.shape) for d, c in zip(depthmaps, confs)]
    for i, P in enumerate(pts):
        for j in range(len(pts)):
            if i == j:
                continue
            proj = P @ w2c[j, :3, :3].T + w2c[j, :3, 3]
            z = proj[..., 2]
            uvh = proj @ K[j].T
            uv = (uvh[..., :2] / uvh[..., 2:3]).round().long()
            u, v = uv[..., 0], uv[..., 1]
            H, W = confs[j].shape
            m = (z > 0) & (0 <= u) & (u < W) & (0 <= v) & (v < H)
            vj, uj = v[m], u[m]
            bad = (z[m] < (1 - tol) * dm[j][vj, uj]) & (res[i][m] < res[j][vj, uj])
            mm = m.clone()
            mm[m] = bad
            res[i][mm] = res[i][mm].clip(max=0)
    return res
