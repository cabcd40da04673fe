"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — CPU restatement of the RASTER + ADAM path.

PARITY UNPINNED.  The arithmetic of this path lives in the third-party package `gsplat`
(requirements.txt:1 of the reference, version unpinned; we restate the 1.4.0 semantics, the
release current when Starst3r 0.4.0 was published).  gsplat is neither vendored under
/root/reference nor installable here, and the reference holds no test or golden vector for
it, so this file restates the published algorithm (gsplat's
fully_fused_projection / spherical_harmonics / isect_tiles / isect_offset_encode /
rasterize_to_pixels and its _torch_impl references) and is held only by (1) closed-form
known-answer tests (tests/test_gs_math_host.py), (2) an independent float64 derivation from the
definitions (oracle/gs_second.py, tests/test_gs_second_derivation.py: radii, tile counts and
blend counts exactly, images to 3e-5, SSIM through torchmetrics' literal pad-filter-crop
procedure, finite differences of this file's own autograd gradients) and (3) the hook for the
real pin: tests/golden/raster_*.npz, which oracle/gen_golden_raster.py writes on a machine that
has gsplat 1.4 + torchmetrics (absent here: the test that reads them is skipped with that reason).  Call sites anchored: starster/gs.py:76-87
(rasterization), :126-136 (loss), :37,159-161 (Adam).

Everything is fp32 PyTorch on CPU.  Index-producing arithmetic (projection -> radii ->
tile ranges -> isect ids) is written as explicit elementwise operations in a fixed order
with no fused multiply-add, which is the order the CUDA kernel follows (compiled with
-fmad=false), so integer outputs can be compared bit-for-bit.
"""
import math

import numpy as np
import torch

SH_C0 = 0.2820947917738781
SH_C1 = 0.48860251190292


# ----------------------------------------------------------------------------- projection
def quat_to_rotmat(quats):
    """wxyz quaternion -> 3x3 rotation, normalised inside (gsplat quat_to_rotmat)."""
    w, x, y, z = quats.unbind(-1)
    inv = 1.0 / torch.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w * inv, x * inv, y * inv, z * inv
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    R = torch.stack([
        1.0 - 2.0 * (y2 + z2), 2.0 * (xy - wz), 2.0 * (xz + wy),
        2.0 * (xy + wz), 1.0 - 2.0 * (x2 + z2), 2.0 * (yz - wx),
        2.0 * (xz - wy), 2.0 * (yz + wx), 1.0 - 2.0 * (x2 + y2)], dim=-1)
    return R.reshape(quats.shape[:-1] + (3, 3))


def _dot3(a0, a1, a2, b0, b1, b2):
    return (a0 * b0 + a1 * b1) + a2 * b2


def quat_scale_to_covar(quats, scales):
    """Sigma = (R S)(R S)^T, returned as the 6 unique entries [.., 6] = (00, 01, 02, 11, 12, 22)."""
    R = quat_to_rotmat(quats)
    M = R * scales[..., None, :]
    m = [[M[..., i, j] for j in range(3)] for i in range(3)]

    def e(i, j):
        return _dot3(m[i][0], m[i][1], m[i][2], m[j][0], m[j][1], m[j][2])
    return torch.stack([e(0, 0), e(0, 1), e(0, 2), e(1, 1), e(1, 2), e(2, 2)], dim=-1)


def project(means, quats, scales, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01, far_plane=1e10,
            radius_clip=0.0):
    """Appendix A.1 (gsplat fully_fused_projection, pinhole, 'classic').  Dense outputs [C, N, ...]:
    radii int32 (0 = culled), means2d, depths, conics.  Differentiable w.r.t. means/quats/scales."""
    C = viewmats.shape[0]
    cov = quat_scale_to_covar(quats, scales)                       # [N, 6]
    s00, s01, s02, s11, s12, s22 = [cov[None, :, i] for i in range(6)]
    Rc = viewmats[:, :3, :3]
    tc = viewmats[:, :3, 3]
    r = [[Rc[:, i, j][:, None] for j in range(3)] for i in range(3)]
    mx, my, mz = means[None, :, 0], means[None, :, 1], means[None, :, 2]
    x = _dot3(r[0][0], r[0][1], r[0][2], mx, my, mz) + tc[:, 0:1]
    y = _dot3(r[1][0], r[1][1], r[1][2], mx, my, mz) + tc[:, 1:2]
    z = _dot3(r[2][0], r[2][1], r[2][2], mx, my, mz) + tc[:, 2:3]

    # T = Rc * Sigma ; Sigma_c = T * Rc^T
    S = [[s00, s01, s02], [s01, s11, s12], [s02, s12, s22]]
    T = [[_dot3(r[i][0], r[i][1], r[i][2], S[0][j], S[1][j], S[2][j]) for j in range(3)] for i in range(3)]
    Sc = [[_dot3(T[i][0], T[i][1], T[i][2], r[j][0], r[j][1], r[j][2]) for j in range(3)] for i in range(3)]

    fx, fy = Ks[:, 0, 0][:, None], Ks[:, 1, 1][:, None]
    cx, cy = Ks[:, 0, 2][:, None], Ks[:, 1, 2][:, None]
    W, H = float(width), float(height)
    tan_fovx = 0.5 * W / fx
    tan_fovy = 0.5 * H / fy
    lim_x_pos = (W - cx) / fx + 0.3 * tan_fovx
    lim_x_neg = cx / fx + 0.3 * tan_fovx
    lim_y_pos = (H - cy) / fy + 0.3 * tan_fovy
    lim_y_neg = cy / fy + 0.3 * tan_fovy
    rz = 1.0 / z
    rz2 = rz * rz
    tx = z * torch.minimum(lim_x_pos, torch.maximum(-lim_x_neg, x * rz))
    ty = z * torch.minimum(lim_y_pos, torch.maximum(-lim_y_neg, y * rz))
    j00 = fx * rz
    j02 = -(fx * tx) * rz2
    j11 = fy * rz
    j12 = -(fy * ty) * rz2
    # cov2d = J Sigma_c J^T with J = [[j00, 0, j02], [0, j11, j12]]
    a0 = j00 * Sc[0][0] + j02 * Sc[2][0]
    a2 = j00 * Sc[0][2] + j02 * Sc[2][2]
    a1 = j00 * Sc[0][1] + j02 * Sc[2][1]
    b1 = j11 * Sc[1][1] + j12 * Sc[2][1]
    b2 = j11 * Sc[1][2] + j12 * Sc[2][2]
    c00 = a0 * j00 + a2 * j02
    c01 = a1 * j11 + a2 * j12
    c11 = b1 * j11 + b2 * j12
    m2x = (fx * x) * rz + cx
    m2y = (fy * y) * rz + cy

    c00 = c00 + eps2d
    c11 = c11 + eps2d
    det = c00 * c11 - c01 * c01
    inv_det = 1.0 / det
    conic_a = c11 * inv_det
    conic_b = -c01 * inv_det
    conic_c = c00 * inv_det
    b = 0.5 * (c00 + c11)
    v1 = b + torch.sqrt(torch.clamp_min(b * b - det, 0.01))
    radius = torch.ceil(3.0 * torch.sqrt(v1))

    valid = (z >= near_plane) & (z <= far_plane) & (det > 0) & (radius > radius_clip)
    valid = valid & ~((m2x + radius <= 0) | (m2x - radius >= W) | (m2y + radius <= 0) | (m2y - radius >= H))
    radii = torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32)
    means2d = torch.stack([m2x, m2y], dim=-1)
    conics = torch.stack([conic_a, conic_b, conic_c], dim=-1)
    return radii, means2d, z, conics


def sh_colors(means, campos, shN):
    """Appendix A.2: degree-1 SH on the first 4 coefficients, +0.5, clamp at 0.  -> [C, N, 3]."""
    dirs = means[None] - campos[:, None, :]
    dirs = dirs * torch.rsqrt((dirs * dirs).sum(-1, keepdim=True))
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    c = shN[None]
    rgb = SH_C0 * c[:, :, 0] + SH_C1 * (-y * c[:, :, 1] + z * c[:, :, 2] - x * c[:, :, 3])
    return torch.clamp_min(rgb + 0.5, 0.0)


# ----------------------------------------------------------------------------- binning (integer exact)
def isect_tiles(means2d, radii, depths, tile_size, tile_w, tile_h):
    """Appendix A.3/A.4 on the packed (camera, gaussian)-ordered visible set.  Returns dict of numpy arrays:
    camera_ids, gaussian_ids (int64), tiles_per_gauss (int32, packed), isect_ids (int64, sorted),
    flatten_ids (int32, sorted; index into the packed arrays)."""
    C, N = radii.shape
    vis = radii > 0
    cam, gau = torch.nonzero(vis, as_tuple=True)
    m = means2d[cam, gau].numpy().astype(np.float32)
    r = radii[cam, gau].numpy().astype(np.float32)
    d = depths[cam, gau].numpy().astype(np.float32)
    ts = np.float32(tile_size)
    tx, ty, tr = m[:, 0] / ts, m[:, 1] / ts, r / ts
    x0 = np.clip(np.floor(tx - tr), 0, tile_w).astype(np.int64)
    x1 = np.clip(np.ceil(tx + tr), 0, tile_w).astype(np.int64)
    y0 = np.clip(np.floor(ty - tr), 0, tile_h).astype(np.int64)
    y1 = np.clip(np.ceil(ty + tr), 0, tile_h).astype(np.int64)
    tiles_per_gauss = ((y1 - y0) * (x1 - x0)).astype(np.int32)
    tile_n_bits = int(math.floor(math.log2(tile_w * tile_h))) + 1 if tile_w * tile_h > 0 else 1
    depth_bits = d.view(np.int32).astype(np.int64) & 0xffffffff
    cam_np = cam.numpy()
    # emit (row-major over y then x) every covered tile of every packed Gaussian, vectorised
    n_t = tiles_per_gauss.astype(np.int64)
    flat = np.repeat(np.arange(len(cam_np), dtype=np.int32), n_t)
    first = np.cumsum(n_t) - n_t
    local = np.arange(int(n_t.sum()), dtype=np.int64) - np.repeat(first, n_t)
    wdt = np.repeat(x1 - x0, n_t)
    yy = np.repeat(y0, n_t) + local // np.maximum(wdt, 1)
    xx = np.repeat(x0, n_t) + local % np.maximum(wdt, 1)
    tile_id = yy * tile_w + xx
    ids = (np.repeat(cam_np.astype(np.int64), n_t) << (32 + tile_n_bits)) | (tile_id << 32) | np.repeat(depth_bits, n_t)
    order = np.argsort(ids, kind="stable")
    return dict(camera_ids=cam_np.astype(np.int64), gaussian_ids=gau.numpy().astype(np.int64),
                tiles_per_gauss=tiles_per_gauss, isect_ids=ids[order], flatten_ids=flat[order],
                tile_n_bits=tile_n_bits)


def isect_offset_encode(isect_ids, C, tile_w, tile_h, tile_n_bits):
    """Appendix A.5: first sorted position of every (camera, tile)."""
    key = isect_ids >> 32                                    # camera << tile_n_bits | tile
    cam = key >> tile_n_bits
    tile = key & ((1 << tile_n_bits) - 1)
    lin = cam * (tile_w * tile_h) + tile
    offs = np.searchsorted(lin, np.arange(C * tile_w * tile_h), side="left").astype(np.int32)
    return offs.reshape(C, tile_h, tile_w)


# ----------------------------------------------------------------------------- blending
def rasterize_to_pixels(means2d, conics, colors, opacities, width, height, tile_size, isect_offsets, flatten_ids,
                        n_cameras):
    """Appendix A.6 on packed inputs (means2d [nnz,2], conics [nnz,3], colors [nnz,3], opacities [nnz]).
    Differentiable.  Returns render [C,H,W,3], alpha [C,H,W,1], last_ids [C,H,W] (int32 numpy),
    n_blend (number of (pixel, Gaussian) pairs actually blended)."""
    C = n_cameras
    tile_h, tile_w = isect_offsets.shape[1:]
    n_isects = len(flatten_ids)
    offs = isect_offsets.reshape(-1)
    render = torch.zeros(C, height, width, 3)
    alpha_out = torch.zeros(C, height, width, 1)
    last_ids = np.zeros((C, height, width), np.int32)
    fl = torch.from_numpy(flatten_ids.astype(np.int64))
    n_blend = 0
    rows, cols = [], []
    for c in range(C):
        for ty in range(tile_h):
            for tx in range(tile_w):
                t = (c * tile_h + ty) * tile_w + tx
                lo = int(offs[t])
                hi = int(offs[t + 1]) if t + 1 < len(offs) else n_isects
                y0, x0 = ty * tile_size, tx * tile_size
                y1, x1 = min(y0 + tile_size, height), min(x0 + tile_size, width)
                if hi <= lo or y1 <= y0 or x1 <= x0:
                    continue
                g = fl[lo:hi]
                py, px = torch.meshgrid(torch.arange(y0, y1, dtype=torch.float32) + 0.5,
                                        torch.arange(x0, x1, dtype=torch.float32) + 0.5, indexing="ij")
                dx = means2d[g, 0][None, None] - px[..., None]
                dy = means2d[g, 1][None, None] - py[..., None]
                a, b, cc = conics[g, 0], conics[g, 1], conics[g, 2]
                sigma = 0.5 * (a * dx * dx + cc * dy * dy) + b * dx * dy
                al = torch.clamp_max(opacities[g] * torch.exp(-sigma), 0.999)
                ok = (sigma >= 0) & (al >= 1.0 / 255.0)
                al_eff = torch.where(ok, al, torch.zeros_like(al))
                T_incl = torch.cumprod(1.0 - al_eff, dim=-1)
                T_excl = torch.cat([torch.ones_like(T_incl[..., :1]), T_incl[..., :-1]], dim=-1)
                stop = ok & (T_incl.detach() <= 1e-4)         # this Gaussian would exhaust the pixel: not blended
                dead = torch.cumsum(stop.to(torch.int32), dim=-1) > 0
                used = ok & ~dead
                w = torch.where(used, al_eff * T_excl, torch.zeros_like(al))
                render[c, y0:y1, x0:x1] = (w[..., None] * colors[g][None, None]).sum(-2)
                alpha_out[c, y0:y1, x0:x1, 0] = w.sum(-1)      # = 1 - T_final
                idx = torch.arange(lo, hi)[None, None].expand_as(used)
                last = torch.where(used, idx, torch.zeros_like(idx)).amax(-1)
                last_ids[c, y0:y1, x0:x1] = last.numpy()
                n_blend += int(used.sum())
    return render, alpha_out, last_ids, n_blend


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height, sh_degree=1, tile_size=16):
    """gsplat.rasterization(...) as called at starster/gs.py:76-87 (packed=True, classic mode, no background)."""
    assert sh_degree == 1
    C, N = viewmats.shape[0], means.shape[0]
    radii, means2d, depths, conics = project(means, quats, scales, viewmats, Ks, width, height)
    campos = torch.linalg.inv(viewmats)[:, :3, 3]
    rgb = sh_colors(means, campos, colors)
    tile_w, tile_h = math.ceil(width / tile_size), math.ceil(height / tile_size)
    b = isect_tiles(means2d.detach(), radii, depths.detach(), tile_size, tile_w, tile_h)
    offsets = isect_offset_encode(b["isect_ids"], C, tile_w, tile_h, b["tile_n_bits"])
    cam = torch.from_numpy(b["camera_ids"])
    gau = torch.from_numpy(b["gaussian_ids"])
    render, alpha, last_ids, n_blend = rasterize_to_pixels(means2d[cam, gau], conics[cam, gau], rgb[cam, gau],
                                                           opacities[gau], width, height, tile_size, offsets,
                                                           b["flatten_ids"], C)
    info = dict(camera_ids=b["camera_ids"], gaussian_ids=b["gaussian_ids"], radii=radii[cam, gau].numpy(),
                means2d=means2d[cam, gau], depths=depths[cam, gau], conics=conics[cam, gau],
                opacities=opacities[gau], tile_width=tile_w, tile_height=tile_h,
                tiles_per_gauss=b["tiles_per_gauss"], isect_ids=b["isect_ids"], flatten_ids=b["flatten_ids"],
                isect_offsets=offsets, width=width, height=height, tile_size=tile_size, n_cameras=C,
                last_ids=last_ids, n_blend=n_blend)
    return render, alpha, info


# ----------------------------------------------------------------------------- loss (torchmetrics SSIM restated)
def gaussian_window(kernel_size=11, sigma=1.5):
    dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1, dtype=torch.float32)
    g = torch.exp(-torch.pow(dist / sigma, 2) / 2)
    return g / g.sum()


def ssim(pred, target, data_range=1.0, kernel_size=11, sigma=1.5, k1=0.01, k2=0.03):
    """torchmetrics StructuralSimilarityIndexMeasure(data_range=1) on [1,3,H,W] inputs:
    11x11 Gaussian window, reflect pad 5 then crop 5 => mean over interior pixels of the valid-window SSIM."""
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    g = gaussian_window(kernel_size, sigma)
    win = (g[:, None] * g[None, :])[None, None].expand(pred.shape[1], 1, kernel_size, kernel_size)
    ch = pred.shape[1]

    def f(x):
        return torch.nn.functional.conv2d(x, win, groups=ch)
    mu_p, mu_t = f(pred), f(target)
    s_pp = f(pred * pred) - mu_p * mu_p
    s_tt = f(target * target) - mu_t * mu_t
    s_pt = f(pred * target) - mu_p * mu_t
    ssim_map = ((2 * mu_p * mu_t + c1) * (2 * s_pt + c2)) / ((mu_p * mu_p + mu_t * mu_t + c1) * (s_pp + s_tt + c2))
    return ssim_map.mean()


def compute_loss(truth_img, render_img, opacities, scales, loss_ssim_fac=0.2, loss_opacity_fac=0.01,
                 loss_scale_fac=0.01):
    """starster/gs.py:126-136 for one view (truth/render [H,W,3])."""
    l1 = torch.nn.functional.l1_loss(truth_img, render_img)
    s = 1 - ssim(truth_img.permute(2, 0, 1).unsqueeze(0), render_img.permute(2, 0, 1).unsqueeze(0))
    loss = l1 * (1 - loss_ssim_fac) + s * loss_ssim_fac
    loss = loss + loss_opacity_fac * torch.abs(torch.sigmoid(opacities)).mean()
    loss = loss + loss_scale_fac * torch.abs(torch.exp(scales)).mean()
    return loss


def train_step(params, states, imgs, viewmats, Ks, width, height, step, lr=1e-3, **loss_kw):
    """One iteration of starster/gs.py:143-161 (enable_pruning=False).  params: dict of leaf tensors
    (means, scales, quats, opacities, shN); states: dict name -> (exp_avg, exp_avg_sq).  Returns loss (float),
    grads dict, and updates params/states in place with torch.optim.Adam semantics (lr, betas .9/.999, eps 1e-8)."""
    for p in params.values():
        p.requires_grad_(True)
        p.grad = None
    render, alpha, info = rasterization(params["means"], params["quats"], params["scales"], params["opacities"],
                                        params["shN"], viewmats, Ks, width, height)
    loss = 0
    for i in range(len(imgs)):
        loss = loss + compute_loss(imgs[i], render[i], params["opacities"], params["scales"], **loss_kw)
    loss.backward()
    grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    b1, b2, eps = 0.9, 0.999, 1e-8
    with torch.no_grad():
        for k, p in params.items():
            m, v = states[k]
            g = grads[k]
            m.lerp_(g, 1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1 = 1 - b1 ** step
            bc2 = 1 - b2 ** step
            denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
            p.addcdiv_(m, denom, value=-(lr / bc1))
    return float(loss), grads, render.detach(), info


# ----------------------------------------------------------------------------- MCMC strategy
# gsplat.MCMCStrategy (reference call sites starster/gs.py:43-45, :146-147, :163-164).  PARITY UNPINNED like the
# rest of this file: restated from the published gsplat 1.4 strategy/mcmc.py + strategy/ops.py (SURVEY.md
# Appendix A.8) and pinned only by closed-form known-answer tests (tests/test_gs_math_host.py).  The random draws
# (torch.multinomial indices, torch.randn noise) are explicit arguments so the CUDA path can be compared on the
# same draws.
MCMC_N_MAX = 51


def mcmc_binoms(n_max=MCMC_N_MAX):
    """MCMCStrategy.initialize_state: binoms[n, k] = C(n, k)."""
    b = torch.zeros((n_max, n_max), dtype=torch.float32)
    for n in range(n_max):
        for k in range(n + 1):
            b[n, k] = math.comb(n, k)
    return b


def mcmc_compute_relocation(opacities, scales, ratios, binoms):
    """gsplat compute_relocation (activated opacities [N], scales [N,3], ratios [N] int): Eq. 9 of the 3DGS-MCMC
    paper, evaluated with the same loop order as the gsplat kernel, in fp32."""
    n_max = binoms.shape[0]
    N = opacities.shape[0]
    ratios = ratios.clamp(1, n_max).to(torch.int64)
    new_op = torch.empty(N, dtype=torch.float32)
    new_sc = torch.empty(N, 3, dtype=torch.float32)
    f32 = np.float32
    for idx in range(N):
        n = int(ratios[idx])
        o = f32(opacities[idx].item())
        no = f32(1.0) - f32(np.power(f32(1.0) - o, f32(1.0) / f32(n), dtype=np.float32))
        denom = f32(0.0)
        for i in range(1, n + 1):
            p = no
            for k in range(i):
                term = f32((-1.0 if k & 1 else 1.0)) / f32(np.sqrt(f32(k + 1))) * p
                denom = f32(denom + f32(binoms[i - 1, k].item()) * term)
                p = f32(p * no)
        new_op[idx] = float(no)
        new_sc[idx] = (o / denom) * scales[idx]
    return new_op, new_sc


def _mcmc_new_values(params, sources, binoms, min_opacity):
    opac = torch.sigmoid(params["opacities"])
    eps = torch.finfo(torch.float32).eps
    ratios = torch.bincount(sources, minlength=opac.shape[0])[sources] + 1
    new_op, new_sc = mcmc_compute_relocation(opac[sources], torch.exp(params["scales"])[sources], ratios, binoms)
    new_op = torch.clamp(new_op, max=1.0 - eps, min=min_opacity)
    return torch.logit(new_op), torch.log(new_sc)


def mcmc_relocate(params, moments, sampled, binoms, min_opacity=0.005):
    """gsplat ops.relocate.  params: dict name -> tensor [N, ...] (raw opacities / scales), modified in place;
    moments: dict name -> (exp_avg, exp_avg_sq); sampled: the torch.multinomial draw (indices into the ALIVE
    list, length = number of dead Gaussians)."""
    opac = torch.sigmoid(params["opacities"])
    dead_mask = opac <= min_opacity
    dead = dead_mask.nonzero(as_tuple=True)[0]
    alive = (~dead_mask).nonzero(as_tuple=True)[0]
    assert len(sampled) == len(dead)
    src = alive[sampled]
    raw_o, raw_s = _mcmc_new_values(params, src, binoms, min_opacity)
    for name, p in params.items():
        if name == "opacities":
            p[src] = raw_o
        elif name == "scales":
            p[src] = raw_s
        p[dead] = p[src]
    for name, (m, v) in moments.items():
        m[src] = 0
        v[src] = 0
    return dead, src


def mcmc_sample_add(params, moments, sampled, binoms, min_opacity=0.005):
    """gsplat ops.sample_add: returns the grown (params, moments) dicts."""
    raw_o, raw_s = _mcmc_new_values(params, sampled, binoms, min_opacity)
    out_p, out_m = {}, {}
    for name, p in params.items():
        p = p.clone()
        if name == "opacities":
            p[sampled] = raw_o
        elif name == "scales":
            p[sampled] = raw_s
        out_p[name] = torch.cat([p, p[sampled]])
    for name, (m, v) in moments.items():
        z = torch.zeros((len(sampled), *m.shape[1:]), dtype=m.dtype)
        out_m[name] = (torch.cat([m, z]), torch.cat([v, z]))
    return out_p, out_m


def mcmc_inject_noise(params, noise, scaler):
    """gsplat ops.inject_noise_to_position with an explicit N(0,1) draw `noise` [N,3]; returns the new means."""
    opac = torch.sigmoid(params["opacities"].flatten())
    c6 = quat_scale_to_covar(params["quats"], torch.exp(params["scales"]))
    covars = torch.stack([c6[:, 0], c6[:, 1], c6[:, 2], c6[:, 1], c6[:, 3], c6[:, 4], c6[:, 2], c6[:, 4], c6[:, 5]],
                         dim=-1).reshape(-1, 3, 3)
    gate = 1.0 / (1.0 + torch.exp(-100.0 * ((1.0 - opac) - 0.995)))
    n = noise * gate.unsqueeze(-1) * scaler
    return params["means"] + torch.einsum("bij,bj->bi", covars, n)
