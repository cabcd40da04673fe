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
