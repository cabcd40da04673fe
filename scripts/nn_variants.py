"""A/B of matcher builds: `ST3R_B200_LIB=<lib> python scripts/nn_variants.py` prints one JSON line with the tcgen05 kernel's
time at M = 4096 / 32768 / 262144 rows against the 512 x 512 map (random-like descriptors), whether it equals the exact
SIMT kernel, and the pair time of extract_correspondences on random-like and on smooth descriptor maps."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from starst3r_b200 import match, synth  # noqa: E402
from scripts.bench_match import timeit  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    out = {"lib": os.path.basename(os.environ.get("ST3R_B200_LIB", "default"))}
    A, B = synth.descriptor_pair(512, 512, seed=0, device=dev)
    DB = B.reshape(-1, 24).contiguous()
    for M in (4096, 32768, 262144):
        Q = A.reshape(-1, 24)[:M].contiguous()
        flop = 2.0 * M * DB.shape[0] * 24
        ms = timeit(lambda: match.nn_argmax(Q, DB, impl="tcgen05"), iters=5 if M > 4096 else 20)
        out[f"M{M}"] = {"ms": round(ms, 4), "tflops": round(flop / ms / 1e9, 1)}
        if M <= 32768:
            out[f"M{M}"]["identical"] = bool(torch.equal(match.nn_argmax(Q, DB, impl="tcgen05"), match.nn_argmax(Q, DB, impl="simt")))
    q = torch.ones(512, 512, device=dev) * 2
    match.NN_SPLIT = False
    match.NN_COOPERATIVE = False

    def pair(X, Y):
        return match.extract_correspondences_device([X, Y, Y, X], [q, q, q, q], 8, impl="tcgen05")
    ms = timeit(lambda: pair(A, B), warm=3, iters=20)
    out["pair_random"] = {"ms": round(ms, 4), "pairs_per_s": round(1000 / ms, 1)}
    # smooth fields (like real MASt3R maps): the synthetic scene's descriptor maps
    net = synth.SyntheticMast3r(2, 512, 512, seed=0, device="cpu", arc_deg=120.0)
    res = net.symmetric_inference({"idx": 1}, {"idx": 0})
    feats = [r["desc"][0].float().to(dev).contiguous() for r in res]
    qonfs = [r["desc_conf"][0].float().to(dev).contiguous() for r in res]
    for split in (False, True):
        match.NN_SPLIT = split
        match.NN_COOPERATIVE = not split
        ms = timeit(lambda: match.extract_correspondences_device(feats, qonfs, 8, impl="tcgen05"), warm=3, iters=10)
        out["pair_smooth_split" if split else "pair_smooth_plain"] = {"ms": round(ms, 4), "pairs_per_s": round(1000 / ms, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
