"""CPU checks of the arithmetic the split-precision tcgen05 matcher (csrc/nn_tc.cu, kSplit) relies on: the tf32
head / tail split is exact, and the three products it accumulates reproduce the fp32 dot product to the bound its
candidate band (DELTA_COEF_SPLIT) assumes.  numpy restatement of split_tf32; no GPU."""
import numpy as np

MASK = np.uint32(0xFFFFE000)      # sign, exponent, 10 leading mantissa bits: what kind::tf32 reads of an fp32 word


def split_tf32(x):
    x = np.asarray(x, np.float32)
    hi = (x.view(np.uint32) & MASK).view(np.float32)
    rem = (x - hi).astype(np.float32)
    lo = (rem.view(np.uint32) & MASK).view(np.float32)
    return hi, lo, rem


def test_split_is_exact_and_tails_are_small():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200_000), rng.standard_normal(1000) * 1e-20,
                        rng.standard_normal(1000) * 1e20, [0.0, -0.0, 1.0, -1.0]]).astype(np.float32)
    hi, lo, rem = split_tf32(x)
    assert np.array_equal(hi.astype(np.float64) + rem.astype(np.float64), x.astype(np.float64))   # x - hi is exact
    assert np.all(np.abs(rem) <= 2.0 ** -10 * np.abs(x))
    assert np.all(np.abs(rem - lo) <= 2.0 ** -10 * np.abs(rem))
    assert np.all(np.abs(x.astype(np.float64) - hi - lo) <= 2.0 ** -20 * np.abs(x))
    # heads and tails are tf32 values: converting them again (truncation or rounding) changes nothing
    for v in (hi, lo):
        assert np.array_equal((v.view(np.uint32) & MASK).view(np.float32), v)


def test_three_products_meet_the_band_the_kernel_assumes():
    rng = np.random.default_rng(1)
    d = 24
    worst_split = worst_plain = 0.0
    for scale in (1.0, 7.3, 1e-3):
        Q = (rng.standard_normal((512, d)) * scale).astype(np.float32)
        D = (rng.standard_normal((4096, d)) * scale).astype(np.float32)
        D[:512] = Q + 1e-3 * scale * rng.standard_normal((512, d)).astype(np.float32)        # near-duplicates too
        qh, ql, _ = split_tf32(Q)
        dh, dl, _ = split_tf32(D)
        exact = Q.astype(np.float64) @ D.astype(np.float64).T
        f = lambda a, b: a.astype(np.float64) @ b.astype(np.float64).T      # noqa: E731  (products of tf32 pairs are exact)
        approx = f(qh, dh) + f(qh, dl) + f(ql, dh)
        plain = f(qh, dh)
        norm = np.linalg.norm(Q.astype(np.float64), axis=1)[:, None] * np.linalg.norm(D.astype(np.float64), axis=1).max()
        worst_split = max(worst_split, float((np.abs(approx - exact) / norm).max()))
        worst_plain = max(worst_plain, float((np.abs(plain - exact) / norm).max()))
    assert worst_split <= 3 * 2.0 ** -20                 # dropped lo.lo term + the two third-level remainders
    assert worst_plain <= 2.0 ** -9                      # the plain TF32 bound the default band (DELTA_COEF) uses
    # DELTA_COEF_SPLIT = 1e-4 = 2 x (this + fp32 accumulation of <= 80 terms, <= ~2e-5 relative), with 2x margin
    assert 2 * (worst_split + 80 * 2.0 ** -22) < 1.0e-4
