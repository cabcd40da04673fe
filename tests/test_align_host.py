"""CPU: the ALIGN math the CUDA kernels use (align_math.cuh, compiled for the host) and the problem flattening
(reconstruct.flatten_problem) against autograd through the oracle, which is itself pinned to the reference."""
import ctypes
import os
import subprocess

import pytest
import torch

from oracle import align_oracle as ao
from starst3r_b200 import reconstruct as rc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = tmp_path_factory.mktemp("host") / "libalign_math_host.so"
    src = os.path.join(ROOT, "tests", "host", "align_math_host.cpp")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def perturbed_params(pb, seed):
    g = torch.Generator().manual_seed(seed)
    p = pb.init_params()
    p["quats"] = torch.nn.functional.normalize(p["quats"] + 0.2 * torch.randn(pb.N, 4, generator=g), dim=1)
    p["trans"] = 0.3 * torch.randn(pb.N, 3, generator=g)
    p["log_sizes"] = 0.2 * torch.randn(pb.N, generator=g)
    p["log_focals"] = p["log_focals"] + 0.1 * torch.randn(pb.N, generator=g)
    p["pps"] = p["pps"] + 0.03 * torch.randn(pb.N, 2, generator=g)
    return p


@pytest.mark.parametrize("name,mode", [("align_match3.pt", 0), ("align_match3.pt", 1), ("align_dust3r3.pt", 0),
                                       ("align_dust3r3.pt", 1)])
def test_loss_and_gradients_vs_autograd(hostlib, name, mode):
    fx = torch.load(os.path.join(GOLD, name), weights_only=False)
    inp = fx["inputs"]
    pb = ao.Problem(inp)
    t, meta = rc.flatten_problem(inp["imgs"], inp["imsizes"], inp["pps"], inp["base_focals"], inp["core_depth"],
                                 inp["anchors"], inp["corres"], inp["corres2d"], inp["preds_21"], inp["mst"], 5.0, "cpu")
    prob = rc.problem_struct(t, meta)
    gamma = 1.1 if mode == 0 else 0.4
    for seed in (0, 1, 2):
        p = perturbed_params(pb, seed)
        if seed == 2:      # tied minimum sizes: torch splits d(min)/d(size) evenly between the ties
            p["log_sizes"][:] = p["log_sizes"].min()
        q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        loss, (K, cam2w, depth, pts3d) = pb.total_loss(q, mode, gamma)
        loss.backward()
        out_loss = ctypes.c_float(0)
        grad = torch.zeros(pb.N, 11)
        cam = torch.zeros(pb.N, 20)
        hostlib.host_align_eval(ctypes.byref(prob), P(p["pps"]), P(p["log_focals"]), P(p["quats"]), P(p["trans"]),
                                P(p["log_sizes"]), mode, ctypes.c_float(gamma), ctypes.c_float(1.1),
                                ctypes.c_float(0.01), ctypes.byref(out_loss), P(grad), P(cam))
        assert abs(out_loss.value - loss.item()) < 2e-5 * max(1.0, abs(loss.item()))
        assert torch.allclose(cam[:, :9].reshape(-1, 3, 3), cam2w[:, :3, :3].detach(), atol=1e-5)
        assert torch.allclose(cam[:, 9:12], cam2w[:, :3, 3].detach(), atol=1e-5)
        want = torch.cat([q["pps"].grad, q["log_focals"].grad[:, None], q["quats"].grad, q["trans"].grad,
                          q["log_sizes"].grad[:, None]], dim=1)
        scale = want.abs().max().item()
        assert (grad - want).abs().max().item() < 2e-3 * scale, (grad - want).abs().max().item() / scale
