"""CPU emulation of the LOGIC of raster_bwd_queue_kernel (csrc/gs_raster.cu, variant 1 of the blend backward): the
split of a (pixel, Gaussian) contribution into the sequential per-pixel part (alpha T, vis dL/dalpha) and the nine
gradient terms rebuilt from a queued record, the ballot-ranked appends, queue overflow drains, the segmented suffix
sum over equal Gaussian slots with padding lanes, and the dense-visit bypass - lane by lane, as the kernel does it -
against a direct per-pixel float64 evaluation of gsplat's rasterize_to_pixels backward.
The CUDA code itself is checked on the GPU by tests/test_experimental_gpu.py."""
import numpy as np

ALPHA_MIN, ALPHA_MAX, T_MIN = 1.0 / 255.0, 0.999, 1e-4
QW, DENSE_MIN, BLOCK = 160, 16, 256


def forward(px, py, A, B, col):
    """Per-pixel front-to-back blend: returns T_final, last contributing index (sorted position)."""
    T, last = 1.0, 0
    for k in range(len(A)):
        dx, dy = A[k, 0] - px, A[k, 1] - py
        sigma = 0.5 * (B[k, 0] * dx * dx + B[k, 2] * dy * dy) + B[k, 1] * dx * dy
        alpha = min(ALPHA_MAX, A[k, 2] * np.exp(-sigma))
        if sigma < 0 or alpha < ALPHA_MIN:
            continue
        nT = T * (1 - alpha)
        if nT <= T_MIN:
            break
        T, last = nT, k
    return T, last


def contribution(px, py, Ak, Bk, colk, T, buf, T_final, v_rgb, v_a):
    """Sequential part for one (pixel, Gaussian): returns None or (fac, w, new T, new buf)."""
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


def contribution_scalar(px, py, Ak, Bk, colk, T, behind_v, T_final, v_rgb, v_a):
    """The same step the way raster_bwd_queue_kernel does it: the colour behind the Gaussian is carried as ONE scalar,
    its dot product with the pixel's upstream colour gradient."""
    dx, dy = Ak[0] - px, Ak[1] - py
    sigma = 0.5 * (Bk[0] * dx * dx + Bk[2] * dy * dy) + Bk[1] * dx * dy
    vis = np.exp(-sigma)
    alpha = min(ALPHA_MAX, Ak[2] * vis)
    if sigma < 0 or alpha < ALPHA_MIN:
        return None
    ra = 1.0 / (1.0 - alpha)
    T = T * ra
    fac = alpha * T
    cv = float((colk * v_rgb).sum())
    v_alpha = T * cv + ra * (T_final * v_a - behind_v)
    w = vis * v_alpha if Ak[2] * vis <= ALPHA_MAX else 0.0
    return fac, w, T, behind_v + fac * cv


def grad_terms(Ak, Bk, dx, dy, fac, w, v_rgb):
    vs = -Ak[2] * w
    return np.array([vs * (Bk[0] * dx + Bk[1] * dy), vs * (Bk[1] * dx + Bk[2] * dy), w, 0.5 * vs * dx * dx,
                     vs * dx * dy, 0.5 * vs * dy * dy, fac * v_rgb[0], fac * v_rgb[1], fac * v_rgb[2]])


def drain(queue, A, B, acc, v_rgb_lanes, px0, py0, stats):
    """32 records at a time: terms per lane, segmented suffix sum with the kernel's `same` test, head adds."""
    n = len(queue)
    for base in range(0, n, 32):
        t = np.empty(32, np.int64)
        g = np.zeros((32, 9))
        have = np.zeros(32, bool)
        for lane in range(32):
            i = base + lane
            if i < n:
                key, fac, w = queue[i]
                tt, pl = key >> 5, key & 31
                t[lane], have[lane] = tt, True
                g[lane] = grad_terms(A[tt], B[tt], A[tt, 0] - (px0 + (pl & 15)), A[tt, 1] - (py0 + (pl >> 4)), fac, w,
                                     v_rgb_lanes[pl])
            else:
                t[lane] = (0xFFFFFFE0 + lane) >> 5
        off = 1
        while off < DENSE_MIN:          # a segment is one visit's records: fewer than DENSE_MIN
            new = g.copy()
            for lane in range(32):
                if lane + off < 32 and t[lane + off] == t[lane]:
                    new[lane] = g[lane] + g[lane + off]
            g, off = new, off * 2
        for lane in range(32):
            if have[lane] and (lane == 0 or t[lane - 1] != t[lane]):
                acc[t[lane]] += g[lane]
                stats["heads"] += 1


def emulate_tile(A, B, col, T_final, last, v_rgb, v_a, stats):
    """Returns [n, 9] gradient sums the way the kernel accumulates them (tile at the origin, 8 warps x 2 rows)."""
    n = len(A)
    out = np.zeros((n, 9))
    nb = (n + BLOCK - 1) // BLOCK
    state = {}                       # per pixel: T, buf
    for b in range(nb):
        batch_end = n - 1 - BLOCK * b
        slots = [batch_end - t for t in range(min(BLOCK, batch_end + 1))]      # slot t holds sorted position batch_end - t
        As, Bs, cs = A[slots], B[slots], col[slots]
        acc = np.zeros((len(slots), 9))
        for wrp in range(8):
            px0, py0 = 0.5, 2 * wrp + 0.5
            lanes = [(px0 + (l & 15), py0 + (l >> 4), 2 * wrp + (l >> 4), l & 15) for l in range(32)]
            queue = []
            for t in range(len(slots)):
                recs = {}
                for l, (px, py, i, j) in enumerate(lanes):
                    if batch_end - t > last[i, j]:
                        continue
                    T, buf = state.get((i, j), (T_final[i, j], 0.0))
                    r = contribution_scalar(px, py, As[t], Bs[t], cs[t], T, buf, T_final[i, j], v_rgb[i, j], v_a[i, j])
                    if r is None:
                        continue
                    fac, w, T, buf = r
                    state[(i, j)] = (T, buf)
                    recs[l] = (fac, w)
                if not recs:
                    continue
                if len(recs) >= DENSE_MIN:
                    stats["dense"] += 1
                    for l, (fac, w) in recs.items():
                        px, py = lanes[l][0], lanes[l][1]
                        acc[t] += grad_terms(As[t], Bs[t], As[t, 0] - px, As[t, 1] - py, fac, w, v_rgb[lanes[l][2], lanes[l][3]])
                else:
                    if len(queue) + len(recs) > QW:
                        stats["overflow_drains"] += 1
                        drain(queue, As, Bs, acc, [v_rgb[i, j] for (_, _, i, j) in lanes], px0, py0, stats)
                        queue = []
                    for l in sorted(recs):          # pos = qlen + popc(ballot & lanes_below)
                        queue.append(((t << 5) | l, recs[l][0], recs[l][1]))
            drain(queue, As, Bs, acc, [v_rgb[i, j] for (_, _, i, j) in lanes], px0, py0, stats)
        out[slots] += acc
    return out


def direct(A, B, col, T_final, last, v_rgb, v_a):
    n = len(A)
    out = np.zeros((n, 9))
    for i in range(16):
        for j in range(16):
            px, py = j + 0.5, i + 0.5
            T, buf = T_final[i, j], np.zeros(3)
            for k in range(last[i, j], -1, -1):
                r = contribution(px, py, A[k], B[k], col[k], T, buf, T_final[i, j], v_rgb[i, j], v_a[i, j])
                if r is None:
                    continue
                fac, w, T, buf = r
                out[k] += grad_terms(A[k], B[k], A[k, 0] - px, A[k, 1] - py, fac, w, v_rgb[i, j])
    return out


def make_tile(n, sigma_px, seed):
    rng = np.random.default_rng(seed)
    A = np.stack([rng.uniform(-2, 18, n), rng.uniform(-2, 18, n), rng.uniform(0.05, 0.6, n)], 1)
    s = sigma_px * np.exp(0.3 * rng.standard_normal((n, 2)))
    rho = rng.uniform(-0.5, 0.5, n)
    cov = np.stack([s[:, 0] ** 2, rho * s[:, 0] * s[:, 1], s[:, 1] ** 2], 1)
    det = cov[:, 0] * cov[:, 2] - cov[:, 1] ** 2
    B = np.stack([cov[:, 2] / det, -cov[:, 1] / det, cov[:, 0] / det], 1)          # conic
    col = rng.uniform(0, 1, (n, 3))
    T_final, last = np.zeros((16, 16)), np.zeros((16, 16), np.int64)
    for i in range(16):
        for j in range(16):
            T_final[i, j], last[i, j] = forward(j + 0.5, i + 0.5, A, B, col)
    return A, B, col, T_final, last, rng.standard_normal((16, 16, 3)), rng.standard_normal((16, 16))


def test_queue_logic_matches_direct_backward():
    seen = {"heads": 0, "dense": 0, "overflow_drains": 0}
    for n, sigma_px, seed in [(300, 0.8, 0), (700, 0.6, 1), (120, 5.0, 2), (400, 2.0, 3)]:
        A, B, col, T_final, last, v_rgb, v_a = make_tile(n, sigma_px, seed)
        want = direct(A, B, col, T_final, last, v_rgb, v_a)
        got = emulate_tile(A, B, col, T_final, last, v_rgb, v_a, seen)
        assert np.abs(want).max() > 1e-3
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12), (n, sigma_px)
    # every code path of the kernel was exercised: segment heads, the dense bypass and a mid-batch overflow drain
    assert seen["heads"] > 100 and seen["dense"] > 10 and seen["overflow_drains"] > 0, seen
