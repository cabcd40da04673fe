"""MATCH microbenchmark of BASELINE.json configs[4] (CUDA events, one GPU): 512 x 512 x 24 descriptor maps, the tcgen05
kernel against the SIMT kernel and against what the reference itself executes on a GPU - fast_nn.py:30-67: cuBLAS GEMM of
8192 x 8192 blocks + torch.max, fp32 and TF32 - plus the measured dense TF32 / bf16 peaks of this GPU (cuBLAS 8192^3),
which are the roofline denominators.  Development / evidence script; bench.py is the judged entry point.

  python scripts/bench_match.py > gpurun_out/<tag>_match_micro.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from starst3r_b200 import match, synth  # noqa: E402


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cublas_peak(dtype, tf32):
    """Dense GEMM rate of cuBLAS at 8192^3 (best of 10), TFLOP/s."""
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    best = 1e9
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = False
    return 2.0 * n ** 3 / best / 1e9


def reference_nn(A, B, block=2 ** 13):
    """fast_nn.py:39-67 on the GPU (the reference's own algorithm, row arg-max only like cdistMatcher.query uses it)."""
    if len(A) * len(B) <= block ** 2:
        return torch.max(A @ B.T, dim=1)[1]
    best = torch.full((len(A),), float("-inf"), device=A.device)
    nn = torch.full((len(A),), -1, dtype=torch.int64, device=A.device)
    for i0 in range(0, len(A), block):
        for j0 in range(0, len(B), block):
            sim, arg = torch.max(A[i0:i0 + block] @ B[j0:j0 + block].T, dim=1)
            upd = sim > best[i0:i0 + block]
            best[i0:i0 + block][upd] = sim[upd]
            nn[i0:i0 + block][upd] = arg[upd] + j0
    return nn


def main():
    dev = torch.device("cuda:0")
    out = {"peaks_tflops": {"cublas_fp32": cublas_peak(torch.float32, False), "cublas_tf32": cublas_peak(torch.float32, True),
                            "cublas_bf16": cublas_peak(torch.bfloat16, False)}}
    A, B = synth.descriptor_pair(512, 512, seed=0, device=dev)
    DB = B.reshape(-1, 24).contiguous()
    nn = {}
    for M in (4096, 32768, 262144):
        Q = A.reshape(-1, 24)[:M].contiguous()
        flop = 2.0 * M * DB.shape[0] * 24
        exact = match.nn_argmax(Q, DB, impl="simt")
        row = {}
        for impl in ("simt", "tcgen05"):
            ms = timeit(lambda: match.nn_argmax(Q, DB, impl=impl), iters=5 if M > 4096 else 20)
            row[impl] = {"ms": ms, "tflops": flop / ms / 1e9,
                         "identical_to_exact": bool(torch.equal(match.nn_argmax(Q, DB, impl=impl), exact))}
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ms = timeit(lambda: reference_nn(Q, DB), warm=1, iters=3 if M > 4096 else 10)
            got = reference_nn(Q, DB)
            row["cublas_tf32+max" if tf32 else "cublas_fp32+max"] = {
                "ms": ms, "tflops": flop / ms / 1e9, "rows_differing_from_exact": int((got != exact.long()).sum())}
        torch.backends.cuda.matmul.allow_tf32 = False
        nn[f"M{M}_N262144"] = row
    out["nn_argmax"] = nn
    # all-pairs over N maps (configs[4]: N = 128; a sample of ordered pairs, seeded reciprocal matching = 4 searches each)
    n_maps, n_pairs = 16, 32
    # "views of one scene": every map is the same field plus its own noise, so any two of them match like A and B above
    g = torch.Generator().manual_seed(100)
    maps = [torch.nn.functional.normalize(A.cpu() + 0.3 * torch.randn(512, 512, 24, generator=g), dim=-1).to(dev)
            for _ in range(n_maps)]
    q = torch.ones(512, 512, device=dev) * 2
    g = torch.Generator().manual_seed(0)
    pairs = [(int(i), int(j)) for i, j in torch.randint(0, n_maps, (n_pairs, 2), generator=g).tolist() if i != j]

    def run_pairs(impl):
        for i, j in pairs:
            match.extract_correspondences_device([maps[i], maps[j], maps[j], maps[i]], [q, q, q, q], 8, impl=impl)
    for impl in ("tcgen05", "simt"):
        ms = timeit(lambda: run_pairs(impl), warm=1, iters=2) / len(pairs)
        out[f"all_pairs_{impl}"] = {"ms_per_pair": ms, "pairs_per_s": 1000.0 / ms, "pairs_timed": len(pairs),
                                    "all_8128_pairs_of_128_maps_s": 8128 * ms / 1000.0}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
