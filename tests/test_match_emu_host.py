"""CPU: the MATCH path of the C ABI - st3r_nn_argmax, st3r_merge_corres, st3r_recip_nn, st3r_extract_corres with their
complete launch sequences (exact SIMT nearest neighbour, device-resident reciprocal search, radix sort, unique) -
compiled for the host by tests/host/build_emu_lib.py (kernel launches rewritten onto the SIMT emulator) and compared
BIT FOR BIT with the golden vectors produced by the unmodified reference (tests/golden/match_*.npz,
oracle/gen_golden.py).  The same entry points, called by the same C signatures, are what tests/test_match_gpu.py
checks on the B200; the tcgen05 kernel is the one piece that only exists there."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SIMT = 1      # ST3R_NN_SIMT


@pytest.fixture
def lib(emu_lib):
    return emu_lib


def P(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def ok(lib, rc):
    assert rc == 0, lib.st3r_last_error()
    assert lib.st3r_emu_launch_failed() == 0, "emulator deadlock / unsupported launch"


def ws_of(nbytes):
    return np.zeros(max(int(nbytes), 256) // 8 + 8, np.uint64)        # 8-byte aligned workspace


def nn_argmax(lib, Q, DB):
    Q, DB = np.ascontiguousarray(Q, np.float32), np.ascontiguousarray(DB, np.float32)
    M, d = Q.shape
    idx, best = np.zeros(M, np.int32), np.zeros(M, np.float32)
    ws = ws_of(lib.st3r_nn_argmax_ws_bytes(M, DB.shape[0], d))
    ok(lib, lib.st3r_nn_argmax(P(Q), M, P(DB), DB.shape[0], d, P(idx), P(best), P(ws), ws.nbytes, SIMT, None))
    return idx, best


def test_nn_argmax_golden(lib):
    g = np.load(os.path.join(GOLD, "match_nn.npz"))
    idx, best = nn_argmax(lib, g["Q"], g["DB"])
    assert np.array_equal(idx, g["nnA"])
    assert np.array_equal(nn_argmax(lib, g["DB"], g["Q"])[0], g["nnB"])
    # the score is the sequential fp32 FMA chain of the winning row
    want = np.zeros(len(idx), np.float32)
    for i, j in enumerate(idx):
        s = np.float32(0)
        for k in range(24):
            s = np.float32(np.float64(g["Q"][i, k]) * np.float64(g["DB"][j, k]) + np.float64(s))
        want[i] = s
    assert np.allclose(best, want, rtol=1e-6)


def test_bruteforce_reciprocal_nns_and_cdist_matcher(emu_backend, golden):
    """SURVEY 8 row a2 on the CPU: match.bruteforce_reciprocal_nns / match.cdistMatcher.query (fast_nn.py:16-84) through
    the product's Python glue and the emulated library - both arg-max directions, int64 numpy results, (None, []) for an
    empty query - against the reference's golden vectors (the GPU run of the same check: tests/test_match_gpu.py)."""
    import torch
    from starst3r_b200 import match
    g = golden("match_nn.npz")
    Q, DB = torch.from_numpy(g["Q"]), torch.from_numpy(g["DB"])
    nnA, nnB = match.bruteforce_reciprocal_nns(Q, DB, device="cpu", dist="dot", block_size=2 ** 13, impl="simt")
    assert nnA.dtype == np.int64 and np.array_equal(nnA, g["nnA"]) and np.array_equal(nnB, g["nnB"])
    assert np.array_equal(nnA, g["nnA_blk"]) and np.array_equal(nnB, g["nnB_blk"])
    m = match.cdistMatcher(DB, device="cpu")
    dis, nn = m.query(Q, dist="dot", block_size=2 ** 13, impl="simt")
    assert dis is None and np.array_equal(nn, g["nnA"])
    assert m.query(Q[:0], dist="dot", impl="simt") == (None, [])
    with pytest.raises(ValueError):
        match.bruteforce_reciprocal_nns(Q, DB, device="cpu", dist="cosine")


def test_merge_corres_golden(lib):
    g = np.load(os.path.join(GOLD, "match_merge.npz"))
    n = len(g["idx1"])
    hw = int(max(g["idx1"].max(), g["idx2"].max())) + 1
    o1, o2, oi, n_out = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(1, np.int32)
    ws = ws_of(lib.st3r_merge_corres_ws_bytes(n))
    ok(lib, lib.st3r_merge_corres(P(np.ascontiguousarray(g["idx1"])), P(np.ascontiguousarray(g["idx2"])), n, hw, hw, P(o1),
                                  P(o2), P(oi), P(n_out), P(ws), ws.nbytes, None))
    k = int(n_out[0])
    assert k == len(g["out1"])
    assert np.array_equal(o1[:k], g["out1"]) and np.array_equal(o2[:k], g["out2"]) and np.array_equal(oi[:k], g["index"])


@pytest.mark.parametrize("n", [1, 700, 1500, 3000, 6000, 12000, 16384, 17000, 2999, 5999])
def test_merge_corres_one_cta_sort_every_instance(lib, n):
    """The one-CTA register sort + unique (small_sort_unique<1 / 2 / 4 / 8 / 16>, 1024 emulated threads; variant 2 = through
    32-bit surrogate words + repair passes, variant 1 = on the 64-bit words) and, beyond 16384 keys and with
    st3r_recip_set_variant(0), the radix chain: unique pairs and first-occurrence indices == np.unique.  n = 2999: every
    idx1 from a pool of five values (runs far longer than the repair passes: the 64-bit network finishes); n = 5999: idx1
    from a pool of 1500 (runs of ~4 in arbitrary idx2 order)."""
    sys.path.insert(0, ROOT)
    from oracle import match_oracle as mo
    rng = np.random.default_rng(n)
    hw = 300 * 200
    pool1, pool2 = rng.integers(0, hw, size=max(n // 3, 1)), rng.integers(0, hw, size=max(n // 3, 1))
    pick = rng.integers(0, len(pool1), size=n)
    idx1 = pool1[pick].astype(np.int32)
    idx2 = np.where(rng.random(n) < 0.7, pool2[pick], rng.integers(0, hw, size=n)).astype(np.int32)
    if n == 2999:
        idx1 = pool1[:5][rng.integers(0, 5, size=n)].astype(np.int32)
    if n == 5999:
        idx1 = pool1[:1500][rng.integers(0, 1500, size=n)].astype(np.int32)
        idx2 = rng.integers(0, hw, size=n).astype(np.int32)
    want = mo.merge_corres(idx1, idx2, ret_xy=False, ret_index=True)
    for variant in ((2, 1, 0) if n in (700, 6000) else (2, 1) if n in (16384, 2999) else (2,)):
        lib.st3r_recip_set_variant(variant)
        o1, o2, oi, n_out = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(1, np.int32)
        ws = ws_of(lib.st3r_merge_corres_ws_bytes(n))
        try:
            ok(lib, lib.st3r_merge_corres(P(idx1), P(idx2), n, hw, hw, P(o1), P(o2), P(oi), P(n_out), P(ws), ws.nbytes, None))
        finally:
            lib.st3r_recip_set_variant(2)
        k = int(n_out[0])
        assert k == len(want[0]), (variant, k, len(want[0]))
        assert np.array_equal(o1[:k], want[0]) and np.array_equal(o2[:k], want[1]) and np.array_equal(oi[:k], want[2])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fast_reciprocal_nns_golden(lib, tag):
    g = np.load(os.path.join(GOLD, "match_recip.npz"))
    A, B = np.ascontiguousarray(g[f"A_{tag}"]), np.ascontiguousarray(g[f"B_{tag}"])
    H1, W1, d = A.shape
    H2, W2, _ = B.shape
    nseed = lib.st3r_recip_seed_count(H1, W1, 8)
    o1, o2, n_out = np.zeros(max(nseed, 1), np.int32), np.zeros(max(nseed, 1), np.int32), np.zeros(1, np.int32)
    ws = ws_of(lib.st3r_recip_nn_ws_bytes(nseed, nseed, 10))
    ok(lib, lib.st3r_recip_nn(P(A), H1, W1, P(B), H2, W2, d, 8, None, 0, 10, P(o1), P(o2), P(n_out), P(ws), ws.nbytes, SIMT,
                              None))
    k = int(n_out[0])
    assert np.array_equal(o1[:k], g[f"idx1_{tag}"]) and np.array_equal(o2[:k], g[f"idx2_{tag}"])


def test_extract_correspondences_golden(lib):
    g = np.load(os.path.join(GOLD, "match_extract.npz"))
    f = [np.ascontiguousarray(g[k]) for k in ("f11", "f21", "f22", "f12")]
    q = [np.ascontiguousarray(g[k]) for k in ("q11", "q21", "q22", "q12")]
    H1, W1, d = f[0].shape
    H2, W2, _ = f[1].shape
    cap = lib.st3r_extract_corres_cap(H1, W1, H2, W2, 8)
    xy1, xy2 = np.zeros((cap, 2), np.int64), np.zeros((cap, 2), np.int64)
    conf, n_out = np.zeros(cap, np.float32), np.zeros(1, np.int32)
    ws = ws_of(lib.st3r_extract_corres_ws_bytes(H1, W1, H2, W2, 8, 10))
    ok(lib, lib.st3r_extract_corres(P(f[0]), P(f[1]), P(f[2]), P(f[3]), P(q[0]), P(q[1]), P(q[2]), P(q[3]), H1, W1, H2, W2, d, 8,
                                    10, P(xy1), P(xy2), P(conf), P(n_out), P(ws), ws.nbytes, SIMT, None))
    k = int(n_out[0])
    assert k == len(g["conf"])
    assert np.array_equal(xy1[:k], g["xy1"]) and np.array_equal(xy2[:k], g["xy2"]) and np.array_equal(conf[:k], g["conf"])


# ---- the tcgen05 matcher on the software model of TMA / mbarriers / tensor memory / tcgen05.mma --------------------
TCGEN05 = 2


@pytest.fixture(params=[(0, 0), (1, 0), (0, 1), (1, 1)], ids=["plain", "cooperative", "split", "cooperative+split"])
def tc_variant(request, lib):
    coop, split = request.param
    lib.st3r_nn_tc_set_cooperative(coop)
    lib.st3r_nn_tc_set_split(split)
    yield request.param
    lib.st3r_nn_tc_set_cooperative(0)
    lib.st3r_nn_tc_set_split(0)


def nn_argmax_impl(lib, Q, DB, impl):
    Q, DB = np.ascontiguousarray(Q, np.float32), np.ascontiguousarray(DB, np.float32)
    M, d = Q.shape
    idx, best = np.zeros(M, np.int32), np.zeros(M, np.float32)
    ws = ws_of(lib.st3r_nn_argmax_ws_bytes(M, DB.shape[0], d))
    ok(lib, lib.st3r_nn_argmax(P(Q), M, P(DB), DB.shape[0], d, P(idx), P(best), P(ws), ws.nbytes, impl, None))
    return idx, best


def test_tcgen05_kernel_source_on_the_software_model(lib, tc_variant):
    """nn_tc_kernel's own source - TMA producer, MMA issuer, eight epilogue warps, the mbarrier protocol, descriptor and
    tensor-memory addressing, candidate lists and exact re-scores - against the exact SIMT kernel: identical indices and
    scores, for every variant (per-thread / warp-cooperative rare path, plain / split precision), on random descriptors,
    on a smooth field (hundreds of near ties inside the TF32 band: lists overflow and are resolved on the spot), with
    exact duplicates (ties -> lowest index) and with ragged sizes (partial tiles, rows beyond the last query tile)."""
    g = np.load(os.path.join(GOLD, "match_nn.npz"))
    idx, _ = nn_argmax_impl(lib, g["Q"], g["DB"], TCGEN05)
    assert np.array_equal(idx, g["nnA"])
    rng = np.random.default_rng(0)
    unit = lambda x: x / np.linalg.norm(x, axis=-1, keepdims=True)      # noqa: E731
    cases = {}
    cases["random"] = (unit(rng.standard_normal((300, 24))), unit(rng.standard_normal((2500, 24))))
    t = np.linspace(0, 1, 3000)[:, None]
    field = unit(np.cos(t * rng.standard_normal((1, 24)) * 2.0 + rng.standard_normal((1, 24))))
    cases["smooth"] = (field[::11][:270] + 1e-4 * rng.standard_normal((270, 24)), field)
    dup = unit(rng.standard_normal((700, 24)))
    cases["duplicates"] = (dup[5:140], np.concatenate([dup, dup[::-1], dup]))
    cases["unnormalised"] = (3.0 * rng.standard_normal((129, 24)), 0.2 * rng.standard_normal((1000, 24)) - 0.1)
    for name, (Q, DB) in cases.items():
        Q, DB = Q.astype(np.float32), DB.astype(np.float32)
        i_tc, b_tc = nn_argmax_impl(lib, Q, DB, TCGEN05)
        i_ex, b_ex = nn_argmax_impl(lib, Q, DB, SIMT)
        assert np.array_equal(i_tc, i_ex), name
        assert np.array_equal(b_tc, b_ex), name
    st = (ctypes.c_ulonglong * 2)()
    ok(lib, lib.st3r_nn_tc_stats(st, 1))
    if not tc_variant[1]:
        assert st[1] > 0          # the smooth field forced exact list resolutions (the rare path ran)


def test_extract_correspondences_golden_tcgen05(lib, tc_variant):
    g = np.load(os.path.join(GOLD, "match_extract.npz"))
    f = [np.ascontiguousarray(g[k]) for k in ("f11", "f21", "f22", "f12")]
    q = [np.ascontiguousarray(g[k]) for k in ("q11", "q21", "q22", "q12")]
    H1, W1, d = f[0].shape
    H2, W2, _ = f[1].shape
    cap = lib.st3r_extract_corres_cap(H1, W1, H2, W2, 8)
    xy1, xy2 = np.zeros((cap, 2), np.int64), np.zeros((cap, 2), np.int64)
    conf, n_out = np.zeros(cap, np.float32), np.zeros(1, np.int32)
    ws = ws_of(lib.st3r_extract_corres_ws_bytes(H1, W1, H2, W2, 8, 10))
    ok(lib, lib.st3r_extract_corres(P(f[0]), P(f[1]), P(f[2]), P(f[3]), P(q[0]), P(q[1]), P(q[2]), P(q[3]), H1, W1, H2, W2, d, 8,
                                    10, P(xy1), P(xy2), P(conf), P(n_out), P(ws), ws.nbytes, TCGEN05, None))
    k = int(n_out[0])
    assert k == len(g["conf"])
    assert np.array_equal(xy1[:k], g["xy1"]) and np.array_equal(xy2[:k], g["xy2"]) and np.array_equal(conf[:k], g["conf"])


@pytest.mark.parametrize("define", [("NN_TC_ONE_ISSUER",), ("NN_TC_ONE_ISSUER", "NN_TC_DECOUPLE")], ids=["one-issuer", "one-issuer-decoupled"])
def test_tcgen05_decoupled_barrier_build(tmp_path, define):
    """The default build has one UMMA-issuing warp per query tile; -DNN_TC_ONE_ISSUER (one MMA thread, shared accumulator
    barriers: the first structure, kept for the instrumented experiments) and -DNN_TC_ONE_ISSUER -DNN_TC_DECOUPLE (that
    thread with per-tile barriers) must complete without a deadlock on the software model and return the exact results."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "host"))
    import build_emu_lib
    from starst3r_b200 import _lib
    path, _ = build_emu_lib.build(str(tmp_path), defines=define)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _lib.parse_header().items():
        getattr(lib, name).restype, getattr(lib, name).argtypes = restype, argtypes
    rng = np.random.default_rng(2)
    Q = rng.standard_normal((300, 24)).astype(np.float32)
    DB = rng.standard_normal((2500, 24)).astype(np.float32)
    for split in (0, 1):
        lib.st3r_nn_tc_set_split(split)
        i_tc, b_tc = nn_argmax_impl(lib, Q, DB, TCGEN05)
        i_ex, b_ex = nn_argmax_impl(lib, Q, DB, SIMT)
        assert np.array_equal(i_tc, i_ex) and np.array_equal(b_tc, b_ex)


def test_tcgen05_variants_on_edge_sizes(lib):
    """Sizes around the tile boundaries (128 DB rows, 256 query rows per CTA), single rows, and inputs that stress the
    candidate band (smooth fields, exact duplicates, badly scaled operands): every variant of the tcgen05 kernel returns
    the exact kernel's indices and scores."""
    rng = np.random.default_rng(7)
    for it in range(14):
        M = int(rng.choice([1, 2, 31, 33, 127, 128, 129, 255, 256, 257, 513]))
        N = int(rng.choice([1, 2, 127, 128, 129, 256, 1000, 1024, 1025, 4097]))
        kind = ["rand", "smooth", "dup", "scaled"][it % 4]
        if kind == "rand":
            Q, DB = rng.standard_normal((M, 24)), rng.standard_normal((N, 24))
        elif kind == "smooth":
            t = np.linspace(0, 1, N)[:, None]
            DB = np.cos(t * rng.standard_normal((1, 24)) * 2 + rng.standard_normal((1, 24)))
            Q = DB[rng.integers(0, N, M)] + 1e-4 * rng.standard_normal((M, 24))
        elif kind == "dup":
            base = rng.standard_normal((max(N // 3, 1), 24))
            DB = np.concatenate([base, base, base, base])[:N]
            Q = DB[rng.integers(0, N, M)]
        else:
            Q, DB = 50 * rng.standard_normal((M, 24)), 1e-3 * rng.standard_normal((N, 24)) + 0.5
        Q, DB = Q.astype(np.float32), DB.astype(np.float32)
        i_ex, b_ex = nn_argmax_impl(lib, Q, DB, SIMT)
        try:
            for coop in (0, 1):
                for split in (0, 1):
                    lib.st3r_nn_tc_set_cooperative(coop)
                    lib.st3r_nn_tc_set_split(split)
                    i_tc, b_tc = nn_argmax_impl(lib, Q, DB, TCGEN05)
                    assert np.array_equal(i_tc, i_ex) and np.array_equal(b_tc, b_ex), (M, N, kind, coop, split)
        finally:
            lib.st3r_nn_tc_set_cooperative(0)
            lib.st3r_nn_tc_set_split(0)
