"""Measures the real error band of the tcgen05 matcher: the candidate band of the plain (kind::tf32)
and of the split-precision kernel is narrowed step by step until the arg-max stops being identical to the exact SIMT
kernel's; the last clean coefficient is the empirical bound of |approximate - exact| / (|q| max|db|), and the built-in
DELTA_COEF / DELTA_COEF_SPLIT must keep a >= 4x margin over it.  Development / evidence script (GPU).

  python scripts/nn_split_margin.py > gpurun_out/<tag>_nn_margin.json"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from starst3r_b200 import _lib, match, synth  # noqa: E402


def smooth_maps(H, W, n, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    X = torch.stack([xx / W, yy / H, (xx + yy) / (W + H)], -1)
    freq = torch.randn(24, 3, generator=g) * 2.5
    return [torch.nn.functional.normalize(torch.cos(X @ freq.T + 0.01 * k) + 0.003 * torch.randn(H, W, 24, generator=g), dim=-1)
            for k in range(n)]


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    cases = {}
    A, B = synth.descriptor_pair(512, 512, seed=0, device=dev)
    cases["random"] = (A.reshape(-1, 24)[::61][:4096].contiguous(), B.reshape(-1, 24).contiguous())
    S = [m.to(dev) for m in smooth_maps(512, 512, 2, seed=7)]
    cases["smooth"] = (S[0].reshape(-1, 24)[::61][:4096].contiguous(), S[1].reshape(-1, 24).contiguous())
    net = synth.SyntheticMast3r(2, 512, 512, seed=0, device="cpu", arc_deg=30.0)
    res = net.symmetric_inference({"idx": 1}, {"idx": 0})
    cases["scene"] = (res[0]["desc"][0].reshape(-1, 24)[::61][:4096].contiguous().to(dev),
                      res[1]["desc"][0].reshape(-1, 24).contiguous().to(dev))
    # unnormalised, large dynamic range: the band scales with |q| max|db|
    g = torch.Generator().manual_seed(3)
    Q = (torch.randn(4096, 24, generator=g) * torch.exp(torch.randn(4096, 1, generator=g))).to(dev)
    D = (torch.randn(262144, 24, generator=g) * torch.exp(0.5 * torch.randn(262144, 1, generator=g))).to(dev)
    cases["unnormalised"] = (Q.contiguous(), D.contiguous())
    exact = {k: match.nn_argmax(q, d, impl="simt", return_score=True) for k, (q, d) in cases.items()}
    out = {"built_in": {"plain": 4.2e-3, "split": 1.0e-4}, "cases": list(cases)}

    def stats():
        st = (ctypes.c_ulonglong * 2)()
        _lib.check(lib.st3r_nn_tc_stats(st, 1), "stats")
        return int(st[1]) / max(int(st[0]), 1)
    # exact list resolutions per scanned query row with the built-in bands: what match._adapt_variant switches on
    ratios = {}
    for split in (False, True):
        match.NN_SPLIT, match.NN_COOPERATIVE = split, False
        for name, (q, d) in cases.items():
            stats()
            match.nn_argmax(q, d, impl="tcgen05")
            torch.cuda.synchronize()
            ratios[f"{name}[split={int(split)}]"] = stats()
    match.NN_SPLIT, match.NN_COOPERATIVE = False, "auto"
    out["resolutions_per_row"] = ratios
    for split in (0, 1):
        match.NN_SPLIT = bool(split)
        coef = 4.2e-3 if not split else 1.0e-4
        rows = []
        try:
            while coef > 1e-9:
                _lib.check(lib.st3r_debug_nn_tc_set_delta_coef(ctypes.c_float(0 if split else coef),
                                                               ctypes.c_float(coef if split else 0)), "set_delta_coef")
                bad = {}
                for name, (q, d) in cases.items():
                    idx, best = match.nn_argmax(q, d, impl="tcgen05", return_score=True)
                    bad[name] = int((idx != exact[name][0]).sum())
                rows.append({"coef": coef, "mismatching_rows_of_4096": bad})
                if sum(bad.values()) > 200:
                    break
                coef /= 2
        finally:
            lib.st3r_debug_nn_tc_set_delta_coef(ctypes.c_float(0), ctypes.c_float(0))
            match.NN_SPLIT = "auto"
        clean = [r["coef"] for r in rows if sum(r["mismatching_rows_of_4096"].values()) == 0]
        out["split" if split else "plain"] = {"sweep": rows, "smallest_clean_coef": min(clean) if clean else None,
                                             "margin_of_built_in": (out["built_in"]["split" if split else "plain"] / min(clean)) if clean else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
