"""Quick MATCH microbenchmarks (CUDA events) - development aid; bench.py is the judged entry point."""
import json
import sys
import time

import torch

from starst3r_b200 import match


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


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    H = W = 512
    A = torch.nn.functional.normalize(torch.randn(H, W, 24, generator=g), dim=-1).to(dev)
    B = torch.nn.functional.normalize(A.cpu() + 0.3 * torch.randn(H, W, 24, generator=g), dim=-1).to(dev)
    out = {}
    for impl in sys.argv[1:] or ["simt", "tcgen05"]:
        try:
            for M in (4096, 32768, 262144):
                Q = A.reshape(-1, 24)[:M].contiguous()
                DB = B.reshape(-1, 24)
                ms = timeit(lambda: match.nn_argmax(Q, DB, impl=impl), iters=5 if M > 4096 else 20)
                out[f"nn_{impl}_M{M}_ms"] = ms
                out[f"nn_{impl}_M{M}_tflops"] = 2.0 * M * DB.shape[0] * 24 / ms / 1e9
            q = torch.ones(H, W, device=dev) * 2
            ms = timeit(lambda: match.extract_correspondences_device([A, B, B, A], [q, q, q, q], 8, impl=impl), iters=5)
            out[f"extract_{impl}_ms"] = ms
            out[f"extract_{impl}_pairs_per_s"] = 1000.0 / ms
        except Exception as e:  # keep going with the other impl
            out[f"{impl}_error"] = repr(e)[:300]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
