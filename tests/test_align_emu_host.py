"""CPU: the ALIGN optimiser kernels' OWN SOURCE (starst3r_b200/csrc/align.cu) executed by the SIMT emulator
(tests/host/simt_emu.h) on the reference fixtures: one iteration = camera forward, per-correspondence loss kernel(s),
camera backward + Adam, in the launch sequence of st3r_align_optimize.  Loss and parameter gradients are compared
with autograd through the oracle (the check tests/test_align_gpu.py makes on the GPU), for the default kernels and for
variant 1 (segmented loss kernels, replicated gradient tables, camera records staged in shared memory), including
launch shapes that force ranges to straddle image pairs."""
import ctypes
import os
import subprocess

import pytest
import torch

from oracle import align_oracle as ao

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "libalign_emu.so"
    src = os.path.join(ROOT, "tests", "host", "align_emu_host.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)], check=True)
    lib = ctypes.CDLL(str(out))
    assert lib.emu_align_cam_grads() == 17
    return lib


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def setup(name, seed=5):
    from starst3r_b200 import reconstruct as rc
    f = torch.load(os.path.join(GOLD, name), weights_only=False)
    inp = f["inputs"]
    pb = ao.Problem(inp)
    t, meta = rc.flatten_problem(inp["imgs"], inp["imsizes"], inp["pps"], inp["base_focals"], inp["core_depth"],
                                 inp["anchors"], inp["corres"], inp["corres2d"], inp["preds_21"], inp["mst"], 5.0, "cpu")
    g = torch.Generator().manual_seed(seed)
    p = pb.init_params()
    p["quats"] = torch.nn.functional.normalize(p["quats"] + 0.2 * torch.randn(pb.N, 4, generator=g), dim=1)
    p["trans"] = 0.3 * torch.randn(pb.N, 3, generator=g)
    p["log_sizes"] = 0.2 * torch.randn(pb.N, generator=g)
    p["log_focals"] = p["log_focals"] + 0.1 * torch.randn(pb.N, generator=g)
    return rc, pb, t, meta, p


def iteration(emu, rc, t, meta, p, variant, mode, gamma, lr, seg_blocks=0, seg_per_warp=0):
    N = meta["N"]
    prob = rc.problem_struct(t, meta)
    q = {k: v.clone().float().contiguous() for k, v in p.items()}
    m, v = torch.zeros(N, 11), torch.zeros(N, 11)
    loss, grad = torch.zeros(1), torch.zeros(N, 11)
    r = emu.emu_align_iteration(variant, ctypes.byref(prob), P(q["pps"]), P(q["log_focals"]), P(q["quats"]), P(q["trans"]),
                                P(q["log_sizes"]), P(m), P(v), mode, 31, ctypes.c_float(gamma), ctypes.c_float(1.1),
                                ctypes.c_float(0.01), ctypes.c_float(lr), 1, ctypes.c_double(0.9), ctypes.c_double(0.9),
                                ctypes.c_double(1e-8), P(loss), P(grad), seg_blocks, seg_per_warp)
    assert r == 0, f"emulator returned {r} (-1 deadlock, -2 launch shape, -3 gradient table not cleared)"
    return loss.item(), grad, q


@pytest.mark.parametrize("name,mode", [("align_match3.pt", 0), ("align_match3.pt", 1), ("align_dust3r3.pt", 0)])
def test_align_kernels_on_the_simt_emulator(emu, name, mode):
    rc, pb, t, meta, p = setup(name)
    gamma = 1.1 if mode == 0 else 0.4
    ref = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss, _ = pb.total_loss(ref, mode, gamma)
    loss.backward()
    want = torch.cat([ref["pps"].grad, ref["log_focals"].grad[:, None], ref["quats"].grad, ref["trans"].grad,
                      ref["log_sizes"].grad[:, None]], dim=1)
    n_main = int(t["e3_a1"].numel() if mode == 0 else t["e2_img1"].numel())
    shapes = [(0, 0, 0), (1, 0, 0)]                                    # (variant, seg_blocks, seg_per_warp)
    if n_main:
        shapes.append((1, 3, (n_main + 23) // 24 // 32 * 32 + 32))     # 24 long ranges: several pairs per warp
        shapes.append((1, (n_main + 255) // 256, 32))                  # one row per warp
    for variant, sb, spw in shapes:
        got_loss, got, _ = iteration(emu, rc, t, meta, p, variant, mode, gamma, 0.0, sb, spw)
        assert abs(got_loss - loss.item()) < 1e-4 * max(1.0, abs(loss.item())), (variant, sb, spw)
        assert (got - want).abs().max().item() < 3e-3 * want.abs().max().item(), (variant, sb, spw)
    # one real Adam step (lr > 0): both variants move the parameters identically.  Adam turns a gradient into a step
    # of ~lr whatever its size, so elements whose gradient is rounding noise (the MST root's pose: the loss is
    # invariant to a global rigid motion) are left out of the comparison.
    steps = [iteration(emu, rc, t, meta, p, variant, mode, gamma, 0.05)[2] for variant in (0, 1)]
    solid = want.abs() > 1e-3 * want.abs().max()
    cols = {"pps": slice(0, 2), "log_focals": slice(2, 3), "quats": slice(3, 7), "trans": slice(7, 10), "log_sizes": slice(10, 11)}
    moved = 0
    for k, sl in cols.items():
        a, b, m = steps[0][k].reshape(pb.N, -1), steps[1][k].reshape(pb.N, -1), solid[:, sl]
        if k == "quats":       # re-normalised after the step: compare rows whose four gradients are all solid
            m = m.all(dim=1, keepdim=True).expand_as(a)
        assert torch.allclose(a[m], b[m], rtol=1e-4, atol=1e-5), k
        moved += int((a[m] != p[k].float().reshape(pb.N, -1)[m]).sum())
    assert moved > 0
    assert torch.allclose(steps[1]["quats"].norm(dim=1), torch.ones(pb.N), atol=1e-5)
