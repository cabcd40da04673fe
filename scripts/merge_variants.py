"""Timing of st3r_merge_corres (device part only) with the one-CTA sort (variant 2: 32-bit surrogate words + repair
passes, variant 1: 64-bit words) and the radix chain (variant 0)."""
import ctypes
import json
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from starst3r_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
out = {}
for n in (1024, 2048, 4096, 8192, 16384):
    rng = np.random.default_rng(n)
    hw = 512 * 512
    d1 = torch.from_numpy(rng.integers(0, hw, size=n).astype(np.int32)).to(dev)
    d2 = torch.from_numpy(rng.integers(0, hw, size=n).astype(np.int32)).to(dev)
    o1, o2, oi = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3))
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.st3r_merge_corres_ws_bytes(n), dtype=torch.uint8, device=dev)
    row = {}
    for variant in (2, 1, 0):
        lib.st3r_recip_set_variant(variant)

        def run():
            _lib.check(lib.st3r_merge_corres(_lib.ptr(d1), _lib.ptr(d2), n, hw, hw, _lib.ptr(o1), _lib.ptr(o2), _lib.ptr(oi),
                                             _lib.ptr(n_out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "merge")
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            run()
        e1.record()
        torch.cuda.synchronize()
        row[{2: "one_cta_surrogate_us", 1: "one_cta_us", 0: "radix_us"}[variant]] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
    lib.st3r_recip_set_variant(2)
    out[n] = row
print(json.dumps(out))
