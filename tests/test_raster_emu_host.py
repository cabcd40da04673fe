"""CPU: the blend kernels' OWN SOURCE (starst3r_b200/csrc/gs_raster.cu: the fragment-pool pair raster_fwd_pool_kernel /
raster_bwd_pool_kernel that the library launches, and the visit-list pair raster_fwd_kernel / raster_bwd_kernel kept as
cross-check) compiled for the host and executed thread by thread by a small SIMT emulator (tests/host/simt_emu.h:
fibers, rendezvous semantics for __syncthreads / shuffles / votes / redux, deadlock detection), against a direct float64
evaluation of gsplat's rasterize_to_pixels forward / backward.  This runs the code the GPU will
run - indexing, scans, compaction, slot walks, masks, dense batches, barriers - without a GPU; what it cannot show is
timing and memory-model behaviour."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TILE = 16
ALPHA_MIN, ALPHA_MAX, T_MIN = 1.0 / 255.0, 0.999, 1e-4
POOL, VISIT = 0, 1          # st3r_gs_set_raster_variant values


def contribution(px, py, Ak, Bk, colk, T, buf, T_final, v_rgb, v_a):
    """Backward step for one (pixel, Gaussian), back to front: returns None or (alpha T, vis dL/dalpha, new T, new buf)."""
    dx, dy = Ak[0] - px, Ak[1] - py
    sigma = 0.5 * (Bk[0] * dx * dx + Bk[2] * dy * dy) + Bk[1] * dx * dy
    vis = np.exp(-sigma)
    alpha = min(ALPHA_MAX, Ak[2] * vis)
    if sigma < 0 or alpha < ALPHA_MIN:
        return None
    ra = 1.0 / (1.0 - alpha)
    T = T * ra
    fac = alpha * T
    v_alpha = float(((colk * T - buf * ra) * v_rgb).sum() + T_final * ra * v_a)
    w = vis * v_alpha if Ak[2] * vis <= ALPHA_MAX else 0.0
    return fac, w, T, buf + colk * fac


def grad_terms(Ak, Bk, dx, dy, fac, w, v_rgb):
    """d/d(x, y, opacity), d/d(conic a, b, c), d/d(rgb) of one (pixel, Gaussian) contribution."""
    vs = -Ak[2] * w
    return np.array([vs * (Bk[0] * dx + Bk[1] * dy), vs * (Bk[1] * dx + Bk[2] * dy), w, 0.5 * vs * dx * dx,
                     vs * dx * dy, 0.5 * vs * dy * dy, fac * v_rgb[0], fac * v_rgb[1], fac * v_rgb[2]])


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "libraster_emu.so"
    src = os.path.join(ROOT, "tests", "host", "raster_emu_host.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_frame(n, C, W, H, sigma_small, sigma_large, seed, big_frac=0.15):
    """Entries e = c * n + g with random 2-D Gaussians; per (camera, tile) lists sorted by depth."""
    rng = np.random.default_rng(seed)
    E = C * n
    A = np.zeros((E, 4), np.float32)          # x, y, opacity, depth
    B = np.zeros((E, 4), np.float32)          # conic a, b, c
    col = np.zeros((E, 4), np.float32)
    A[:, 0] = rng.uniform(-4, W + 4, E)
    A[:, 1] = rng.uniform(-4, H + 4, E)
    A[:, 2] = rng.uniform(0.05, 0.7, E)
    A[:, 3] = rng.uniform(1, 10, E)
    big = rng.random(E) < big_frac
    s = np.where(big[:, None], sigma_large, sigma_small) * np.exp(0.3 * rng.standard_normal((E, 2)))
    rho = rng.uniform(-0.5, 0.5, E)
    cov = np.stack([s[:, 0] ** 2, rho * s[:, 0] * s[:, 1], s[:, 1] ** 2], 1)
    det = cov[:, 0] * cov[:, 2] - cov[:, 1] ** 2
    B[:, 0], B[:, 1], B[:, 2] = cov[:, 2] / det, -cov[:, 1] / det, cov[:, 0] / det
    col[:, :3] = rng.uniform(0, 1, (E, 3))
    radius = 3.5 * s.max(1) + 1
    tw, th = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    flatten, offsets = [], []
    for c in range(C):
        for ty in range(th):
            for tx in range(tw):
                offsets.append(len(flatten))
                e = np.arange(c * n, (c + 1) * n)
                hit = ((A[e, 0] + radius[e] > tx * TILE) & (A[e, 0] - radius[e] < (tx + 1) * TILE) &
                       (A[e, 1] + radius[e] > ty * TILE) & (A[e, 1] - radius[e] < (ty + 1) * TILE))
                ids = e[hit]
                flatten.extend(ids[np.argsort(A[ids, 3], kind="stable")].tolist())
    return (A, B, col, np.asarray(offsets, np.int32), np.asarray([len(flatten)], np.int32),
            np.asarray(flatten + [0], np.uint32), tw, th)


def reference(A, B, col, offsets, n_isect, flatten, tw, th, C, W, H, v_render, v_alphas):
    """float64, pixel by pixel: forward outputs and the gradients w.r.t. (x, y, opacity), conic and colour."""
    A64, B64, c64 = A.astype(np.float64), B.astype(np.float64), col.astype(np.float64)[:, :3]
    render = np.zeros((C, H, W, 3))
    alphas = np.zeros((C, H, W))
    g = np.zeros((A.shape[0], 9))
    ends = list(offsets[1:]) + [int(n_isect[0])]
    for c in range(C):
        for i in range(H):
            for j in range(W):
                t = c * tw * th + (i // TILE) * tw + (j // TILE)
                ids = flatten[offsets[t]:ends[t]]
                px, py = j + 0.5, i + 0.5
                T, last, rgb = 1.0, -1, np.zeros(3)
                for k, e in enumerate(ids):
                    dx, dy = A64[e, 0] - px, A64[e, 1] - py
                    sigma = 0.5 * (B64[e, 0] * dx * dx + B64[e, 2] * dy * dy) + B64[e, 1] * dx * dy
                    alpha = min(ALPHA_MAX, A64[e, 2] * np.exp(-sigma))
                    if sigma < 0 or alpha < ALPHA_MIN:
                        continue
                    nT = T * (1 - alpha)
                    if nT <= T_MIN:
                        break
                    rgb += c64[e] * alpha * T
                    T, last = nT, k
                render[c, i, j], alphas[c, i, j] = rgb, 1 - T
                Tb, buf = T, np.zeros(3)
                for k in range(last, -1, -1):
                    e = ids[k]
                    r = contribution(px, py, A64[e], B64[e], c64[e], Tb, buf, T, v_render[c, i, j], v_alphas[c, i, j])
                    if r is None:
                        continue
                    fac, w, Tb, buf = r
                    g[e] += grad_terms(A64[e], B64[e], A64[e, 0] - px, A64[e, 1] - py, fac, w, v_render[c, i, j])
    return render, alphas, g


def run_fwd(emu, variant, frame, C, W, H):
    A, B, col, offsets, n_isect, flatten, tw, th = frame
    render = np.zeros((C, H, W, 3), np.float32)
    alphas = np.zeros((C, H, W), np.float32)
    last_ids = np.zeros((C, H, W), np.int32)
    n_blend = np.zeros(1, np.uint64)
    rc = emu.emu_raster_fwd(variant, P(offsets), P(n_isect), P(flatten), P(A), P(B), P(col), C, W, H, P(render), P(alphas),
                            P(last_ids), P(n_blend))
    assert rc == 0, f"deadlock in the forward kernel, variant {variant}"
    return render, alphas, last_ids, int(n_blend[0])


def run_bwd(emu, variant, frame, C, W, H, alphas, last_ids, v_render, v_alphas):
    A, B, col, offsets, n_isect, flatten, tw, th = frame
    vA, vB, vC = np.zeros_like(A), np.zeros_like(B), np.zeros_like(col)
    rc = emu.emu_raster_bwd(variant, P(offsets), P(n_isect), P(flatten), P(A), P(B), P(col), C, W, H, P(alphas),
                            P(last_ids), P(v_render), P(v_alphas), P(vA), P(vB), P(vC))
    assert rc == 0, f"deadlock in the backward kernel, variant {variant}"
    assert np.all(vA[:, 3] == 0) and np.all(vB[:, 3] == 0) and np.all(vC[:, 3] == 0)
    return np.concatenate([vA[:, :3], vB[:, :3], vC[:, :3]], axis=1)


# small splats + 15 % tile-sized ones (pool batches, pool overflow); mostly tile-sized ones (dense batches); > 256 per tile
@pytest.mark.parametrize("n,C,W,H,ss,sl,big,seed", [(420, 2, 40, 24, 0.8, 5.0, 0.15, 0), (700, 1, 24, 24, 0.7, 3.0, 0.15, 1),
                                                    (300, 1, 40, 24, 0.8, 7.0, 0.9, 2)])
def test_blend_kernels_on_the_simt_emulator(emu, n, C, W, H, ss, sl, big, seed):
    frame = make_frame(n, C, W, H, ss, sl, seed, big)
    A, B, col, offsets, n_isect, flatten, tw, th = frame
    per_tile = np.diff(np.r_[offsets, n_isect])
    rng = np.random.default_rng(seed + 10)
    v_render = rng.standard_normal((C, H, W, 3)).astype(np.float32)
    v_alphas = rng.standard_normal((C, H, W)).astype(np.float32)
    counts = (ctypes.c_long * 16)()
    emu.emu_counts(counts, 1)
    render, alphas, last_ids, n_blend = run_fwd(emu, POOL, frame, C, W, H)
    want_render, want_alphas, want_g = reference(A, B, col, offsets, n_isect, flatten, tw, th, C, W, H, v_render, v_alphas)
    assert np.abs(render - want_render).max() < 2e-5 and np.abs(alphas - want_alphas).max() < 2e-5
    assert n_blend > 5 * C * H * W
    # the visit-list forward: same per-pixel arithmetic in the same order => identical outputs, bit for bit
    render2, alphas2, last2, n_blend2 = run_fwd(emu, VISIT, frame, C, W, H)
    assert np.array_equal(render, render2) and np.array_equal(alphas, alphas2) and np.array_equal(last_ids, last2)
    assert n_blend == n_blend2
    scale = np.abs(want_g).max(0) + 1e-12
    got = {}
    for variant in (POOL, VISIT):
        g = run_bwd(emu, variant, frame, C, W, H, alphas, last_ids, v_render, v_alphas)
        got[variant] = g
        err = np.abs(g - want_g).max(0) / scale
        # fp32 kernels (and fp32 forward state: long per-pixel chains under tile-sized splats) vs the float64 evaluation;
        # both kernel pairs show the same deviation to three digits
        assert err.max() < 5e-4, (variant, err)
    assert (np.abs(got[POOL] - got[VISIT]).max(0) / scale).max() < 1e-4
    # path coverage of the fragment-pool kernels in this run (per-lane hits)
    emu.emu_counts(counts, 1)
    if big < 0.5:
        # box pixels tested, contributions, Gaussians deferred because the pool was full
        assert counts[10] > 0 and counts[11] > 0 and counts[12] > 0, list(counts)
    else:
        assert counts[13] > 0 and counts[14] > 0, list(counts)    # dense batches, forward and backward
    if n >= 700:
        assert per_tile.max() > 256                  # more than one batch per tile


def test_blend_kernel_pairs_agree_on_edge_shapes(emu):
    """Randomised sweep over the shapes that stress the bookkeeping: one Gaussian, list lengths around the batch size
    (256 +- 1) and the candidate-count steps (32 / 64 / 128), images smaller than a tile and ragged ones, near-opaque
    splats (early termination, alpha clamp), almost transparent ones (no box at all), all sizes of splats: the two
    kernel pairs agree (forward bit for bit) and none of them deadlocks."""
    rng = np.random.default_rng(123)
    for it in range(32):
        n = int(rng.choice([1, 3, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 400, 900]))
        C, W, H = int(rng.integers(1, 3)), int(rng.choice([5, 16, 17, 31, 33, 48])), int(rng.choice([3, 16, 18, 32, 40]))
        ss, sl = float(rng.choice([0.3, 0.8, 1.5, 3.0])), float(rng.choice([2.0, 5.0, 12.0]))
        big = float(rng.choice([0.0, 0.15, 0.5, 1.0]))
        frame = make_frame(n, C, W, H, ss, sl, int(rng.integers(1 << 30)), big)
        A = frame[0]
        u = rng.random()
        if u < 0.3:
            A[:, 2] = rng.uniform(0.9, 1.0, len(A))
        elif u < 0.5:
            A[:, 2] = rng.uniform(0.001, 0.01, len(A))
        v_render = rng.standard_normal((C, H, W, 3)).astype(np.float32)
        v_alphas = rng.standard_normal((C, H, W)).astype(np.float32)
        render, alphas, last_ids, n_blend = run_fwd(emu, POOL, frame, C, W, H)
        render2, alphas2, last2, n_blend2 = run_fwd(emu, VISIT, frame, C, W, H)
        assert np.array_equal(render, render2) and np.array_equal(alphas, alphas2), (it, n, C, W, H)
        assert np.array_equal(last_ids, last2) and n_blend == n_blend2, (it, n, C, W, H)
        outs = [run_bwd(emu, v, frame, C, W, H, alphas, last_ids, v_render, v_alphas) for v in (POOL, VISIT)]
        scale = np.abs(outs[1]).max(0) + 1e-12
        assert (np.abs(outs[0] - outs[1]).max(0) / scale).max() < 1e-4, (it, n, C, W, H)
