"""GPU parity of the MATCH path: CUDA kernels (through the C ABI) vs the CPU oracle and the reference's golden
vectors.  Integer outputs must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import match_oracle as mo

pytestmark = pytest.mark.gpu

IMPLS = ["simt", "tcgen05"]


def unit(x):
    return (x / np.linalg.norm(x, axis=-1, keepdims=True)).astype(np.float32)


def _impl_ok(impl):
    if impl == "tcgen05":
        from starst3r_b200 import _lib
        # built with the tcgen05 kernel?
        return True
    return True


@pytest.mark.parametrize("impl", IMPLS)
def test_nn_argmax_golden(golden, cuda_device, impl):
    from starst3r_b200 import match
    g = golden("match_nn.npz")
    Q, DB = torch.from_numpy(g["Q"]).to(cuda_device), torch.from_numpy(g["DB"]).to(cuda_device)
    assert np.array_equal(match.nn_argmax(Q, DB, impl=impl).cpu().numpy(), g["nnA"])
    assert np.array_equal(match.nn_argmax(DB, Q, impl=impl).cpu().numpy(), g["nnB"])


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("M,N", [(1, 1), (2, 5), (7, 129), (128, 128), (129, 4097), (300, 20000), (1000, 66000)])
def test_nn_argmax_vs_oracle(cuda_device, impl, M, N):
    from starst3r_b200 import match
    rng = np.random.default_rng(M * 7919 + N)
    Q = unit(rng.standard_normal((M, 24)))
    DB = unit(rng.standard_normal((N, 24)))
    if N > 64:
        DB[N // 2] = DB[3]            # tie on purpose
        Q[0] = DB[3]
    ref_idx, ref_best = mo.nn_argmax_dot_c(Q, DB)
    idx, best = match.nn_argmax(torch.from_numpy(Q).to(cuda_device), torch.from_numpy(DB).to(cuda_device),
                                impl=impl, return_score=True)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(best.cpu().numpy(), ref_best)      # scores are bit-identical (fp32 FMA chain)


@pytest.mark.parametrize("impl", IMPLS)
def test_nn_argmax_unnormalised_and_negative(cuda_device, impl):
    """Non-unit descriptors (large dynamic range) and all-negative scores."""
    from starst3r_b200 import match
    rng = np.random.default_rng(11)
    Q = (rng.standard_normal((257, 24)) * np.exp(rng.standard_normal((257, 1)) * 2)).astype(np.float32)
    DB = (rng.standard_normal((5000, 24)) * np.exp(rng.standard_normal((5000, 1)) * 2)).astype(np.float32)
    Q[:10] = -np.abs(Q[:10])
    DB = np.abs(DB)
    ref_idx, _ = mo.nn_argmax_dot_c(Q, DB)
    idx = match.nn_argmax(torch.from_numpy(Q).to(cuda_device), torch.from_numpy(DB).to(cuda_device), impl=impl)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)


def test_nn_generic_dim(cuda_device):
    from starst3r_b200 import match
    rng = np.random.default_rng(5)
    for d in (3, 16, 32, 40):
        Q = rng.standard_normal((50, d)).astype(np.float32)
        DB = rng.standard_normal((777, d)).astype(np.float32)
        ref_idx, _ = mo.nn_argmax_dot_c(Q, DB)
        idx = match.nn_argmax(torch.from_numpy(Q).to(cuda_device), torch.from_numpy(DB).to(cuda_device), impl="simt")
        assert np.array_equal(idx.cpu().numpy(), ref_idx)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fast_reciprocal_nns_golden(golden, cuda_device, impl, tag):
    from starst3r_b200 import match
    g = golden("match_recip.npz")
    A, B = torch.from_numpy(g[f"A_{tag}"]), torch.from_numpy(g[f"B_{tag}"])
    i1, i2 = match.fast_reciprocal_NNs(A, B, subsample_or_initxy1=8, ret_xy=False, device=cuda_device, dist="dot",
                                       block_size=2 ** 13, impl=impl)
    assert i1.dtype == np.int32
    assert np.array_equal(i1, g[f"idx1_{tag}"]) and np.array_equal(i2, g[f"idx2_{tag}"])
    x1, x2 = match.fast_reciprocal_NNs(A, B, subsample_or_initxy1=8, ret_xy=True, device=cuda_device, dist="dot",
                                       impl=impl)
    assert np.array_equal(x1, g[f"xy1_{tag}"]) and np.array_equal(x2, g[f"xy2_{tag}"])


def test_fast_reciprocal_nns_general_form_equals_fused(golden, cuda_device):
    """Host-loop form (explicit seeds / ret_basin) and the fused device chain agree."""
    from starst3r_b200 import match
    g = golden("match_recip.npz")
    A, B = torch.from_numpy(g["A_a"]), torch.from_numpy(g["B_a"])
    H1, W1 = A.shape[:2]
    i1, i2, basin = match.fast_reciprocal_NNs(A, B, 8, ret_xy=False, ret_basin=True, device=cuda_device, dist="dot")
    assert basin.shape == (H1 * W1 + 1,)
    y1, x1 = np.mgrid[4:H1:8, 4:W1:8].reshape(2, -1)
    j1, j2 = match.fast_reciprocal_NNs(A, B, (x1, y1), ret_xy=False, device=cuda_device, dist="dot")
    # max_iter = 1 for explicit seeds (fast_nn.py:128): converged subset of the 10-iteration result
    full = set(zip(g["idx1_a"].tolist(), g["idx2_a"].tolist()))
    assert set(zip(j1.tolist(), j2.tolist())) <= full


def test_merge_corres_golden(golden, cuda_device):
    from starst3r_b200 import match
    g = golden("match_merge.npz")
    o1, o2, idx = match.merge_corres(g["idx1"], g["idx2"], ret_xy=False, ret_index=True, device=cuda_device)
    assert np.array_equal(o1, g["out1"]) and np.array_equal(o2, g["out2"]) and np.array_equal(idx, g["index"])
    o1, o2 = match.merge_corres(np.zeros(0, np.int32), np.zeros(0, np.int32), ret_xy=False, device=cuda_device)
    assert len(o1) == 0


@pytest.mark.parametrize("n", [1, 5, 1000, 1025, 4096, 5000, 16384, 16385, 40000, 4095, 16383])
def test_merge_corres_both_sort_variants_vs_numpy(cuda_device, n):
    """st3r_merge_corres with the one-CTA register sort (lists up to 16384 keys; variant 2, the default: 32-bit surrogate
    words + repair passes, variant 1: the 64-bit words) and with the radix chain (st3r_recip_set_variant(0), also the
    path of longer lists) against the oracle's np.unique restatement of fast_nn.py:87-106: unique pairs in (idx1, idx2)
    order and the index of each pair's FIRST occurrence - many duplicates, so the stability of the order is observable.
    n = 4095: idx1 from five values (the repair passes give up, the 64-bit network finishes); n = 16383: idx1 from a pool
    of 4000 with unrelated idx2 (runs of ~4 to repair)."""
    from oracle import match_oracle as mo
    from starst3r_b200 import _lib, match
    lib = _lib.load()
    rng = np.random.default_rng(n)
    hw = 512 * 512
    pool1 = rng.integers(0, hw, size=max(n // 3, 1), dtype=np.int64)
    pool2 = rng.integers(0, hw, size=max(n // 3, 1), dtype=np.int64)
    pick = rng.integers(0, len(pool1), size=n)
    idx1 = pool1[pick].astype(np.int32)
    idx2 = np.where(rng.random(n) < 0.7, pool2[pick], rng.integers(0, hw, size=n)).astype(np.int32)
    if n == 4095:
        idx1 = pool1[:5][rng.integers(0, 5, size=n)].astype(np.int32)
    if n == 16383:
        idx1 = pool1[:4000][rng.integers(0, 4000, size=n)].astype(np.int32)
        idx2 = rng.integers(0, hw, size=n).astype(np.int32)
    want = mo.merge_corres(idx1, idx2, ret_xy=False, ret_index=True)
    try:
        for variant in (2, 1, 0):
            _lib.check(lib.st3r_recip_set_variant(variant), "st3r_recip_set_variant")
            got = match.merge_corres(idx1, idx2, (512, 512), (512, 512), ret_xy=False, ret_index=True, device=cuda_device)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), (variant, n)
    finally:
        _lib.check(lib.st3r_recip_set_variant(2), "st3r_recip_set_variant")


@pytest.mark.parametrize("impl", IMPLS)
def test_extract_correspondences_golden(golden, cuda_device, impl):
    from starst3r_b200 import match
    g = golden("match_extract.npz")
    T = torch.from_numpy
    xy1, xy2, conf = match.extract_correspondences([T(g["f11"]), T(g["f21"]), T(g["f22"]), T(g["f12"])],
                                                   [T(g["q11"]), T(g["q21"]), T(g["q22"]), T(g["q12"])],
                                                   subsample=8, device=cuda_device, impl=impl)
    assert xy1.dtype == torch.int64 and conf.dtype == torch.float32 and xy1.is_cuda
    assert np.array_equal(xy1.cpu().numpy(), g["xy1"]) and np.array_equal(xy2.cpu().numpy(), g["xy2"])
    assert np.array_equal(conf.cpu().numpy(), g["conf"])
    from starst3r_b200 import _lib
    lib = _lib.load()
    for variant in (0, 1):      # the radix chain / the 64-bit one-CTA sort instead of the surrogate-word sort: identical
        try:
            _lib.check(lib.st3r_recip_set_variant(variant), "st3r_recip_set_variant")
            r = match.extract_correspondences([T(g["f11"]), T(g["f21"]), T(g["f22"]), T(g["f12"])],
                                              [T(g["q11"]), T(g["q21"]), T(g["q22"]), T(g["q12"])],
                                              subsample=8, device=cuda_device, impl=impl)
        finally:
            _lib.check(lib.st3r_recip_set_variant(2), "st3r_recip_set_variant")
        assert torch.equal(r[0], xy1) and torch.equal(r[1], xy2) and torch.equal(r[2], conf)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("H,W,noise", [(128, 128, 0.3), (256, 256, 0.3), (96, 160, 1.0)])
def test_fast_reciprocal_nns_vs_oracle(cuda_device, impl, H, W, noise):
    from starst3r_b200 import match
    rng = np.random.default_rng(H + W)
    A = unit(rng.standard_normal((H, W, 24)))
    B = unit(A + noise * rng.standard_normal((H, W, 24)).astype(np.float32))
    r1, r2 = mo.fast_reciprocal_nns(A, B, 8)
    i1, i2 = match.fast_reciprocal_NNs(torch.from_numpy(A), torch.from_numpy(B), 8, ret_xy=False,
                                       device=cuda_device, dist="dot", impl=impl)
    assert np.array_equal(i1, r1) and np.array_equal(i2, r2)


def test_bruteforce_reciprocal_nns_and_cdist_matcher(golden, cuda_device):
    """SURVEY 8 rows a1 / a2 through the reference-named entry points (fast_nn.py:16-84): bruteforce_reciprocal_nns
    returns BOTH arg-max directions as int64 numpy arrays, cdistMatcher.query returns (None, nnA) and (None, []) for
    an empty query - against the reference's golden vectors (single block and 8192-blocked call are the same integers)."""
    from starst3r_b200 import match
    g = golden("match_nn.npz")
    Q, DB = torch.from_numpy(g["Q"]), torch.from_numpy(g["DB"])
    for impl in IMPLS:
        nnA, nnB = match.bruteforce_reciprocal_nns(Q, DB, device=cuda_device, dist="dot", block_size=2 ** 13, impl=impl)
        assert isinstance(nnA, np.ndarray) and nnA.dtype == np.int64 and nnB.dtype == np.int64
        assert np.array_equal(nnA, g["nnA"]) and np.array_equal(nnB, g["nnB"])
        assert np.array_equal(nnA, g["nnA_blk"]) and np.array_equal(nnB, g["nnB_blk"])
        nnA, nnB = match.bruteforce_reciprocal_nns(g["Q"], g["DB"], device=cuda_device, dist="dot", impl=impl)   # numpy in
        assert np.array_equal(nnA, g["nnA"]) and np.array_equal(nnB, g["nnB"])
        m = match.cdistMatcher(DB, device=cuda_device)
        dis, nn = m.query(Q, dist="dot", block_size=2 ** 13, impl=impl)
        assert dis is None and nn.dtype == np.int64 and np.array_equal(nn, g["nnA"])
        dis, nn = m.query(Q[:0], dist="dot", block_size=2 ** 13, impl=impl)           # fast_nn.py:79-80
        assert dis is None and nn == []
    with pytest.raises(ValueError):
        match.bruteforce_reciprocal_nns(Q, DB, device=cuda_device, dist="cosine")      # fast_nn.py:36-37
    with pytest.raises(NotImplementedError):                                           # CPU-only reference branch
        match.bruteforce_reciprocal_nns(Q, DB, device=cuda_device, dist="l2")


def test_edge_cases(cuda_device):
    from starst3r_b200 import match
    P = torch.ones(3, 3, 24)
    i1, i2 = match.fast_reciprocal_NNs(P, P, 8, ret_xy=False, device=cuda_device, dist="dot")
    assert len(i1) == 0 and len(i2) == 0
    P = torch.ones(16, 16, 24)
    i1, i2 = match.fast_reciprocal_NNs(P, P, 8, ret_xy=False, device=cuda_device, dist="dot")
    assert list(i1) == [0] and list(i2) == [0]


def _smooth_maps(H, W, n, seed):
    """Descriptor fields like real MASt3R maps (and the synthetic scene): smooth functions of the pixel position plus a
    little noise, so ~100 DB columns per query row lie inside the TF32 error band of the best score."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    X = torch.stack([xx / W, yy / H, (xx + yy) / (W + H)], -1)
    freq = torch.randn(24, 3, generator=g) * 2.5
    return [torch.nn.functional.normalize(torch.cos(X @ freq.T + 0.01 * k) + 0.003 * torch.randn(H, W, 24, generator=g),
                                          dim=-1) for k in range(n)]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("kind", ["random", "smooth"])
def test_full_size_vs_oracle(cuda_device, impl, kind):
    """BASELINE configs[1] sizes against the C oracle itself (not the other kernel): one M = 4096 x N = 262 144 arg-max
    (where the TF32 band of the tcgen05 kernel holds the most columns) with bit-identical indices AND scores, and one
    512 x 512 fast_reciprocal_NNs, on random descriptors and on smooth fields (the rare path's stress case)."""
    from starst3r_b200 import match
    if kind == "random":
        g = torch.Generator().manual_seed(5)
        A = torch.nn.functional.normalize(torch.randn(512, 512, 24, generator=g), dim=-1)
        B = torch.nn.functional.normalize(A + 0.3 * torch.randn(512, 512, 24, generator=g), dim=-1)
    else:
        A, B = _smooth_maps(512, 512, 2, seed=7)
    An, Bn = A.numpy(), B.numpy()
    ys, xs = np.mgrid[4:512:8, 4:512:8].reshape(2, -1)
    Q = An.reshape(-1, 24)[np.sort(xs + 512 * ys)]                       # the 4096 seed rows of the first NN call
    ref_idx, ref_best = mo.nn_argmax_dot_c(Q, Bn.reshape(-1, 24))
    idx, best = match.nn_argmax(torch.from_numpy(Q).to(cuda_device), B.reshape(-1, 24).to(cuda_device), impl=impl,
                                return_score=True)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(best.cpu().numpy(), ref_best)
    r1, r2 = mo.fast_reciprocal_nns(An, Bn, 8)
    i1, i2 = match.fast_reciprocal_NNs(A, B, 8, ret_xy=False, device=cuda_device, dist="dot", impl=impl)
    assert len(r1) > (1000 if kind == "random" else 50)
    assert np.array_equal(i1, r1) and np.array_equal(i2, r2)


@pytest.mark.parametrize("impl", IMPLS)
def test_full_size_properties(cuda_device, impl):
    """512x512 (BASELINE configs[1] map size): size-independent properties instead of a CPU oracle sweep:
    every returned pair is a mutual nearest neighbour (reciprocity), output sorted & unique, and matching a
    map against itself returns exactly the seed grid."""
    from starst3r_b200 import match
    g = torch.Generator().manual_seed(0)
    A = torch.nn.functional.normalize(torch.randn(512, 512, 24, generator=g), dim=-1)
    B = torch.nn.functional.normalize(A + 0.3 * torch.randn(512, 512, 24, generator=g), dim=-1)
    i1, i2 = match.fast_reciprocal_NNs(A, B, 8, ret_xy=False, device=cuda_device, dist="dot", impl=impl)
    assert len(i1) > 1000
    key = i1.astype(np.int64) << 32 | i2.astype(np.int64)
    assert np.all(np.diff(key) > 0)
    Ad, Bd = A.reshape(-1, 24).to(cuda_device), B.reshape(-1, 24).to(cuda_device)
    t1 = torch.from_numpy(i1.astype(np.int64)).to(cuda_device)
    t2 = torch.from_numpy(i2.astype(np.int64)).to(cuda_device)
    assert torch.equal(match.nn_argmax(Ad[t1], Bd, impl="simt").long(), t2)
    assert torch.equal(match.nn_argmax(Bd[t2], Ad, impl="simt").long(), t1)
    s1, s2 = match.fast_reciprocal_NNs(A, A, 8, ret_xy=False, device=cuda_device, dist="dot", impl=impl)
    y, x = np.mgrid[4:512:8, 4:512:8].reshape(2, -1)
    assert np.array_equal(s1, np.sort(x + 512 * y)) and np.array_equal(s1, s2)


@pytest.fixture
def nn_split(monkeypatch):
    from starst3r_b200 import match
    monkeypatch.setattr(match, "NN_SPLIT", True)
    yield
    from starst3r_b200 import _lib
    _lib.load().st3r_nn_tc_set_split(0)


def test_split_precision_variant_vs_oracle(golden, cuda_device, nn_split):
    """The split-precision tcgen05 kernel (what "auto" selects on smooth descriptor fields) pinned against the oracle
    and the reference's golden vectors on its own: every shape of the arg-max sweep, ties, unnormalised inputs,
    extract_correspondences, and the full-size checks."""
    for M, N in [(1, 1), (2, 5), (7, 129), (128, 128), (129, 4097), (300, 20000), (1000, 66000)]:
        test_nn_argmax_vs_oracle(cuda_device, "tcgen05", M, N)
    test_nn_argmax_golden(golden, cuda_device, "tcgen05")
    test_nn_argmax_unnormalised_and_negative(cuda_device, "tcgen05")
    test_extract_correspondences_golden(golden, cuda_device, "tcgen05")
    test_full_size_vs_oracle(cuda_device, "tcgen05", "smooth")
    test_full_size_properties(cuda_device, "tcgen05")


def test_matcher_variants_identical_and_auto_selected(cuda_device):
    """The tcgen05 matcher's data-dependent switches - warp-cooperative rare path (st3r_nn_tc_set_cooperative) and split
    precision (st3r_nn_tc_set_split) - return bit-identical correspondences in every combination, on random AND on
    smooth descriptor fields (many near-tie columns per row), and "auto" follows the kernels' statistics (exact list
    resolutions per query row): smooth fields switch the split on, random ones switch it off again."""
    import ctypes
    from starst3r_b200 import _lib, match, synth
    lib = _lib.load()
    H = W = 256
    smooth = [m.to(cuda_device) for m in _smooth_maps(H, W, 4, seed=0)]
    A, B = synth.descriptor_pair(H, W, seed=3, device=cuda_device)
    cases = {"random": [A, B, B, A], "smooth": smooth}
    q = [torch.ones(H, W, device=cuda_device) for _ in range(4)]
    out, ratio = {}, {}

    def stats():
        st = (ctypes.c_ulonglong * 2)()
        _lib.check(lib.st3r_nn_tc_stats(st, 1), "stats")
        return int(st[1]) / max(int(st[0]), 1)
    try:
        for coop in (False, True):
            for split in (False, True):
                match.NN_COOPERATIVE, match.NN_SPLIT = coop, split
                for name, feats in cases.items():
                    stats()
                    out[name, coop, split] = [t.cpu() for t in match.extract_correspondences(feats, q, 8, device=cuda_device)]
                    ratio[name, coop, split] = stats()
        for name in cases:
            for key, val in out.items():
                if key[0] == name:
                    for a, b in zip(out[name, False, False], val):
                        assert torch.equal(a, b), key
            assert out[name, False, False][0].shape[0] > 50
        # the regimes the switches separate
        assert ratio["random", False, False] < 0.1 < match.SPLIT_ON_RATIO < ratio["smooth", False, False], ratio
        assert ratio["smooth", False, True] < 0.1 * ratio["smooth", False, False], ratio
        assert ratio["random", False, False] < match.SPLIT_OFF_RATIO, ratio
        match.NN_COOPERATIVE = match.NN_SPLIT = "auto"
        match._variant.update(on=False, split=False, probe_in=0)
        stats()
        match.extract_correspondences(cases["smooth"], q, 8, device=cuda_device)
        assert match._variant["split"] is True                       # plain kernel resolved > 0.5 lists per row
        for _ in range(3):                                           # ... and stays on while the data stays smooth
            match.extract_correspondences(cases["smooth"], q, 8, device=cuda_device)
            assert match._variant["split"] is True and match._variant["ran_split"] is True
        match._variant["probe_in"] = 1                               # the next call is the periodic plain-kernel probe
        match.extract_correspondences(cases["random"], q, 8, device=cuda_device)
        assert match._variant["ran_split"] is False and match._variant["split"] is False
        match.extract_correspondences(cases["random"], q, 8, device=cuda_device)
        assert match._variant["split"] is False and match._variant["on"] is False
    finally:
        match.NN_COOPERATIVE = match.NN_SPLIT = "auto"
        match._variant.update(on=False, split=False, probe_in=0)
