"""CPU: dense-geometry kernels of starst3r_b200/csrc/align_dense.cu run from their own source on the SIMT emulator
(tests/host/simt_emu.h): the Weiszfeld focal kernel in its single-CTA form and in its thread-block-CLUSTER form (8 CTAs
per image, partial sums exchanged through distributed shared memory, one cluster barrier per IRLS iteration - the
emulator keeps one copy of the shared-memory section per CTA of the cluster), and canonical_view, against the oracle
restatements of dust3r/post_process.py:36-58 and sparse_ga.py:817-855."""
import ctypes
import os
import subprocess

import pytest
import torch

from oracle import align_oracle as ao

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "libdense_emu.so"
    src = os.path.join(ROOT, "tests", "host", "dense_emu_host.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def pointmap(H, W, focal, seed):
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    z = 2.0 + torch.rand(H, W, generator=g) * 3.0
    X = torch.stack([(xs - W / 2) / focal * z, (ys - H / 2) / focal * z, z], -1)
    X = X + 0.02 * torch.randn(H, W, 3, generator=g)
    X[::7, ::5] *= 1.5                      # a few outliers: what the IRLS re-weighting is there for
    return X.contiguous()


@pytest.mark.parametrize("H,W", [(48, 64), (40, 40)])
def test_weiszfeld_single_cta_and_cluster(emu, H, W):
    maps = torch.stack([pointmap(H, W, 70.0, 0), pointmap(H, W, 55.0, 1)]).contiguous()
    want = torch.stack([ao.estimate_focal_weiszfeld(m, 0.5, 3.5) for m in maps])
    for variant in (0, 1):
        out = torch.zeros(2)
        rc = emu.emu_focal_weiszfeld(variant, P(maps), 2, H, W, ctypes.c_float(0.5), ctypes.c_float(3.5), P(out))
        assert rc == 0, f"deadlock in variant {variant}"
        assert torch.allclose(out, want, rtol=2e-4), (variant, out, want)
    assert abs(want[0].item() - 70.0) < 3.0 and abs(want[1].item() - 55.0) < 3.0


def test_canonical_view(emu):
    g = torch.Generator().manual_seed(3)
    Pn, H, W, S = 3, 32, 40, 8
    base = pointmap(H, W, 50.0, 2)
    pts = torch.stack([base + 0.01 * torch.randn(H, W, 3, generator=g) for _ in range(Pn)]).contiguous()
    conf = (1.0 + 5.0 * torch.rand(Pn, H, W, generator=g)).contiguous()
    canon, canon2, cconf = torch.zeros(H, W, 3), torch.zeros(H, W), torch.zeros(H, W)
    assert emu.emu_canonical_view(P(pts), P(conf), Pn, H, W, S, P(canon), P(canon2), P(cconf)) == 0
    w_canon, w_canon2, w_conf = ao.canonical_view(pts, conf, S)
    assert torch.allclose(canon, w_canon, rtol=1e-4, atol=1e-5)
    assert torch.allclose(canon2, w_canon2, rtol=1e-4, atol=1e-5)
    assert torch.allclose(cconf, w_conf, rtol=1e-4, atol=1e-5)
