"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) - an independent SECOND derivation of the RASTER path, in float64
numpy, written from the mathematical definitions rather than from oracle/gs_oracle.py: matrices through numpy.linalg,
one Python loop per Gaussian and per pixel, no tiles lists, no vectorised masks.  PARITY UNPINNED like gs_oracle.py
(gsplat is not available here): its purpose is to catch a mis-remembered detail in ONE of the two restatements - the
pieces the judge's review lists: radius / tile-range rounding, eps2d, the alpha clamp, the 1/255 cut, the T <= 1e-4
termination, the signs of the degree-1 SH basis, the SSIM window and its valid region.
References: gsplat 1.4 `_torch_impl.py` (_quat_scale_to_covar_preci, _persp_proj, _fully_fused_projection,
_isect_tiles, accumulate / _rasterize_to_pixels) and `cuda/csrc` (rasterize_to_pixels_fwd), SURVEY.md Appendix A;
torchmetrics 1.x `functional/image/ssim.py::_ssim_update`, SURVEY.md Appendix B."""
import math

import numpy as np

C0 = 0.28209479177387814        # Y_0^0
C1 = 0.4886025119029199         # |Y_1^m| prefactor


def rotation_wxyz(q):
    """Rotation matrix of a (w, x, y, z) quaternion through the Rodrigues form R = (w^2 - v.v) I + 2 v v^T + 2 w [v]x."""
    q = np.asarray(q, np.float64)
    q = q / np.linalg.norm(q)
    w, v = q[0], q[1:]
    vx = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return (w * w - v @ v) * np.eye(3) + 2 * np.outer(v, v) + 2 * w * vx


def project_one(mean, quat, scale, viewmat, K, W, H, eps2d=0.3, near=0.01, far=1e10):
    """One Gaussian in one pinhole camera -> None (culled) or dict(mean2d, depth, conic 2x2, radius int, cov2d)."""
    R = rotation_wxyz(quat)
    Sigma = R @ np.diag(np.asarray(scale, np.float64) ** 2) @ R.T
    Rcw, t = viewmat[:3, :3], viewmat[:3, 3]
    p = Rcw @ mean + t
    if not (near <= p[2] <= far):
        return None
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    # the Jacobian is evaluated at a point clamped to a frustum 30 % wider than the image (gsplat _persp_proj)
    tan_x, tan_y = 0.5 * W / fx, 0.5 * H / fy
    x = p[2] * min((W - cx) / fx + 0.3 * tan_x, max(-(cx / fx + 0.3 * tan_x), p[0] / p[2]))
    y = p[2] * min((H - cy) / fy + 0.3 * tan_y, max(-(cy / fy + 0.3 * tan_y), p[1] / p[2]))
    J = np.array([[fx / p[2], 0.0, -fx * x / p[2] ** 2], [0.0, fy / p[2], -fy * y / p[2] ** 2]])
    cov2d = J @ (Rcw @ Sigma @ Rcw.T) @ J.T + eps2d * np.eye(2)
    det = np.linalg.det(cov2d)
    if det <= 0:
        return None
    mean2d = np.array([fx * p[0] / p[2] + cx, fy * p[1] / p[2] + cy])
    half_trace = 0.5 * np.trace(cov2d)
    lam_max = half_trace + math.sqrt(max(0.01, half_trace ** 2 - det))      # larger eigenvalue, floored discriminant
    radius = math.ceil(3.0 * math.sqrt(lam_max))
    if radius <= 0:
        return None
    if mean2d[0] + radius <= 0 or mean2d[0] - radius >= W or mean2d[1] + radius <= 0 or mean2d[1] - radius >= H:
        return None
    return dict(mean2d=mean2d, depth=p[2], conic=np.linalg.inv(cov2d), radius=int(radius), cov2d=cov2d)


def sh_colour(mean, campos, sh):
    """Degree-1 real spherical harmonics in gsplat's coefficient order (1, y, z, x) with the Condon-Shortley signs
    (-y, +z, -x), + 0.5, clamped at zero.  sh: [>=4, 3]."""
    d = np.asarray(mean, np.float64) - campos
    d = d / np.linalg.norm(d)
    rgb = C0 * sh[0] + C1 * (-d[1] * sh[1] + d[2] * sh[2] - d[0] * sh[3])
    return np.maximum(rgb + 0.5, 0.0)


def tile_rect(mean2d, radius, tile, tile_w, tile_h):
    lo_x = min(max(0, math.floor((mean2d[0] - radius) / tile)), tile_w)
    hi_x = min(max(0, math.ceil((mean2d[0] + radius) / tile)), tile_w)
    lo_y = min(max(0, math.floor((mean2d[1] - radius) / tile)), tile_h)
    hi_y = min(max(0, math.ceil((mean2d[1] + radius) / tile)), tile_h)
    return lo_x, hi_x, lo_y, hi_y


def render(means, quats, scales, opacities, sh, viewmats, Ks, W, H, tile=16):
    """-> (image [C,H,W,3], alpha [C,H,W], radii [C,N] (0 = culled), tiles touched [C,N], blended pairs)."""
    means, quats, scales = (np.asarray(a, np.float64) for a in (means, quats, scales))
    opacities, sh = np.asarray(opacities, np.float64), np.asarray(sh, np.float64)
    viewmats, Ks = np.asarray(viewmats, np.float64), np.asarray(Ks, np.float64)
    C, N = len(viewmats), len(means)
    tile_w, tile_h = math.ceil(W / tile), math.ceil(H / tile)
    img, alpha = np.zeros((C, H, W, 3)), np.zeros((C, H, W))
    radii, touched = np.zeros((C, N), np.int64), np.zeros((C, N), np.int64)
    n_blend = 0
    for c in range(C):
        campos = -viewmats[c][:3, :3].T @ viewmats[c][:3, 3]
        vis = []
        for g in range(N):
            pr = project_one(means[g], quats[g], scales[g], viewmats[c], Ks[c], W, H)
            if pr is None:
                continue
            pr["rect"] = tile_rect(pr["mean2d"], pr["radius"], tile, tile_w, tile_h)
            pr["rgb"] = sh_colour(means[g], campos, sh[g])
            pr["opacity"] = opacities[g]
            radii[c, g] = pr["radius"]
            touched[c, g] = (pr["rect"][1] - pr["rect"][0]) * (pr["rect"][3] - pr["rect"][2])
            vis.append((np.float32(pr["depth"]), g, pr))      # the sort key is the fp32 depth, ties by index
        vis.sort(key=lambda e: (e[0], e[1]))
        for py in range(H):
            for px in range(W):
                tx, ty = px // tile, py // tile
                T, acc = 1.0, np.zeros(3)
                for _, g, pr in vis:
                    r = pr["rect"]
                    if not (r[0] <= tx < r[1] and r[2] <= ty < r[3]):
                        continue                               # a Gaussian only reaches the tiles its square touches
                    d = pr["mean2d"] - np.array([px + 0.5, py + 0.5])
                    sigma = 0.5 * d @ pr["conic"] @ d
                    a = min(0.999, pr["opacity"] * math.exp(-sigma))
                    if sigma < 0 or a < 1.0 / 255.0:
                        continue
                    if T * (1 - a) <= 1e-4:
                        break                                  # this Gaussian would exhaust the pixel: it is NOT blended
                    acc += pr["rgb"] * a * T
                    T *= 1 - a
                    n_blend += 1
                img[c, py, px], alpha[c, py, px] = acc, 1 - T
    return img, alpha, radii, touched, n_blend


def ssim_torchmetrics(pred, target, data_range=1.0, kernel=11, sigma=1.5, k1=0.01, k2=0.03):
    """torchmetrics' procedure, literally: reflect-pad both images by (kernel - 1) / 2, 'valid' Gaussian-window
    statistics (back at the input size), crop the pad again, mean.  pred / target: [H, W, 3] float64."""
    pad = (kernel - 1) // 2
    g = np.exp(-(np.arange(kernel) - pad) ** 2 / (2 * sigma ** 2))
    g = g / g.sum()
    win = np.outer(g, g)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    H, W, _ = pred.shape
    total, count = 0.0, 0
    for ch in range(3):
        P = np.pad(pred[..., ch], pad, mode="reflect")
        Tt = np.pad(target[..., ch], pad, mode="reflect")
        for y in range(pad, H - pad):                          # the crop keeps output pixels pad .. H - pad - 1
            for x in range(pad, W - pad):
                p, t = P[y:y + kernel, x:x + kernel], Tt[y:y + kernel, x:x + kernel]
                mp, mt = (win * p).sum(), (win * t).sum()
                spp, stt, spt = (win * p * p).sum() - mp * mp, (win * t * t).sum() - mt * mt, (win * p * t).sum() - mp * mt
                total += ((2 * mp * mt + c1) * (2 * spt + c2)) / ((mp * mp + mt * mt + c1) * (spp + stt + c2))
                count += 1
    return total / count
