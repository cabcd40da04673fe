"""Generates tests/golden/*.npz by running the UNMODIFIED reference (mounted read-only at
/root/reference) on seeded synthetic inputs, in the build container (CPU).

    python -m oracle.gen_golden            # writes tests/golden/

The fixtures carry inputs AND reference outputs so that they travel to the GPU box,
where /root/reference does not exist.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

from oracle import ref_bootstrap

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def unit(x):
    return (x / np.linalg.norm(x, axis=-1, keepdims=True)).astype(np.float32)


def desc_pair(rng, H1, W1, H2, W2, d=24, noise=0.3):
    """Correlated descriptor maps: map 2 is a shifted / cropped noisy copy of map 1."""
    H, W = max(H1, H2) + 4, max(W1, W2) + 4
    base = rng.standard_normal((H, W, d)).astype(np.float32)
    A = unit(base[:H1, :W1])
    B = unit(base[2:2 + H2, 3:3 + W2] + noise * rng.standard_normal((H2, W2, d)).astype(np.float32))
    return A, B


def gen_match(fast_nn, sparse_ga):
    kw = dict(device="cpu", dist="dot", block_size=2 ** 13)
    rng = np.random.default_rng(20241220)

    # 1. brute-force NN, single block, blocked, ties
    Q = unit(rng.standard_normal((37, 24)))
    DB = unit(rng.standard_normal((1000, 24)))
    DB[700] = DB[13]; DB[701] = DB[13]            # exact duplicates -> ties resolved to the lowest index
    Q[5] = DB[13]
    nnA, nnB = fast_nn.bruteforce_reciprocal_nns(torch.from_numpy(Q), torch.from_numpy(DB), **kw)
    nnA_blk, nnB_blk = fast_nn.bruteforce_reciprocal_nns(torch.from_numpy(Q), torch.from_numpy(DB), device="cpu",
                                                         dist="dot", block_size=2 ** 4)
    np.savez_compressed(os.path.join(OUT, "match_nn.npz"), Q=Q, DB=DB, nnA=nnA, nnB=nnB, nnA_blk=nnA_blk,
                        nnB_blk=nnB_blk)

    # 2. fast_reciprocal_NNs, ragged shapes, three noise levels
    out = {}
    for tag, (H1, W1, H2, W2, noise) in {"a": (48, 64, 40, 56, 0.1), "b": (50, 37, 50, 37, 0.3),
                                          "c": (32, 32, 64, 48, 1.0)}.items():
        A, B = desc_pair(rng, H1, W1, H2, W2, noise=noise)
        i1, i2 = fast_nn.fast_reciprocal_NNs(torch.from_numpy(A), torch.from_numpy(B), subsample_or_initxy1=8,
                                             ret_xy=False, **kw)
        x1, x2 = fast_nn.fast_reciprocal_NNs(torch.from_numpy(A), torch.from_numpy(B), subsample_or_initxy1=8,
                                             ret_xy=True, **kw)
        out.update({f"A_{tag}": A, f"B_{tag}": B, f"idx1_{tag}": i1, f"idx2_{tag}": i2,
                    f"xy1_{tag}": np.ascontiguousarray(x1), f"xy2_{tag}": np.ascontiguousarray(x2)})
    np.savez_compressed(os.path.join(OUT, "match_recip.npz"), **out)

    # 3. merge_corres with duplicates
    i1 = rng.integers(0, 50, 400).astype(np.int32)
    i2 = rng.integers(0, 40, 400).astype(np.int32)
    m1, m2, mi = fast_nn.merge_corres(i1, i2, ret_xy=False, ret_index=True)
    np.savez_compressed(os.path.join(OUT, "match_merge.npz"), idx1=i1, idx2=i2, out1=m1, out2=m2, index=mi)

    # 4. extract_correspondences for one image pair (sparse_ga.py:595-630)
    H1, W1, H2, W2 = 48, 64, 40, 56
    f11, f21 = desc_pair(rng, H1, W1, H2, W2, noise=0.2)
    f12 = unit(f11 + 0.15 * rng.standard_normal(f11.shape).astype(np.float32))
    f22 = unit(f21 + 0.15 * rng.standard_normal(f21.shape).astype(np.float32))
    q11, q12 = [1 + 9 * rng.random((H1, W1)).astype(np.float32) for _ in range(2)]
    q21, q22 = [1 + 9 * rng.random((H2, W2)).astype(np.float32) for _ in range(2)]
    T = torch.from_numpy
    xy1, xy2, conf = sparse_ga.extract_correspondences([T(f11), T(f21), T(f22), T(f12)],
                                                       [T(q11), T(q21), T(q22), T(q12)], subsample=8, device="cpu")
    np.savez_compressed(os.path.join(OUT, "match_extract.npz"), f11=f11, f21=f21, f22=f22, f12=f12, q11=q11, q21=q21,
                        q22=q22, q12=q12, xy1=xy1.numpy(), xy2=xy2.numpy(), conf=conf.numpy())


def main():
    os.makedirs(OUT, exist_ok=True)
    fast_nn, sparse_ga = ref_bootstrap.bootstrap()
    which = sys.argv[1:] or ["match"]
    if "match" in which:
        gen_match(fast_nn, sparse_ga)
    if "align" in which:
        from oracle import gen_golden_align
        gen_golden_align.generate(sparse_ga, OUT)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
