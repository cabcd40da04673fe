"""Generates tests/golden/raster_*.npz - the pin the RASTER oracle lacks - from the REAL dependencies of the reference:
gsplat (requirements.txt:1; the restatement follows 1.4.0) and torchmetrics (starster/gs.py:8).  Neither is vendored under
/root/reference nor installable in the build container, so this script cannot run there; a maintainer with
`pip install gsplat==1.4.0 torchmetrics` and any CUDA GPU runs

    python oracle/gen_golden_raster.py

and commits the files it writes.  tests/test_gs_second_derivation.py::test_golden_vectors_from_gsplat_when_present then
holds oracle/gs_oracle.py to them on the CPU, tests/test_gs_gpu.py::test_golden_vectors_from_gsplat_when_present the CUDA
path on the B200.  The scenes are the seeded synthetic ones the other tests use (starst3r_b200/synth.py): a few hundred
Gaussians, 2-3 views, well under 100 kB per file.  Calls mirror starster/gs.py:76-87 (rasterization) and :126-131 (SSIM)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [dict(tag="a", n=300, C=2, W=80, H=48, seed=0, scale_mult=8.0),
         dict(tag="b", n=120, C=3, W=70, H=37, seed=5, scale_mult=25.0)]


def main():
    try:
        import gsplat
        from torchmetrics.image import StructuralSimilarityIndexMeasure
    except ImportError as e:                                    # the build container ends here
        print(f"gen_golden_raster: {e}; install gsplat==1.4.0 and torchmetrics to generate the RASTER golden vectors")
        return 1
    from starst3r_b200 import synth
    dev = torch.device("cuda")
    for case in CASES:
        sp = synth.random_splats(case["n"], seed=case["seed"], scale_mode="rand")
        sp["scales"] = sp["scales"] * case["scale_mult"]
        viewmats, Ks = synth.look_at_cameras(case["C"], case["W"], case["H"])
        means, quats = sp["means"].to(dev), sp["quats"].to(dev)
        scales, opac = torch.exp(sp["scales"]).to(dev), torch.sigmoid(sp["opacities"]).to(dev)
        colors = sp["shN"].to(dev)
        render, alpha, info = gsplat.rasterization(means, quats, scales, opac, colors, viewmats.to(dev), Ks.to(dev),
                                                   case["W"], case["H"], sh_degree=1, packed=True)
        g = torch.Generator().manual_seed(case["seed"])
        truth = torch.rand(case["C"], case["H"], case["W"], 3, generator=g)
        ssim = StructuralSimilarityIndexMeasure(data_range=1.0).to(dev)
        s = ssim(render[:1].permute(0, 3, 1, 2).clamp(0, 1), truth[:1].permute(0, 3, 1, 2).to(dev)).item()
        out = dict(means=sp["means"].numpy(), quats=sp["quats"].numpy(), scales=scales.cpu().numpy(), opacities=opac.cpu().numpy(),
                   colors=sp["shN"].numpy(), viewmats=viewmats.numpy(), Ks=Ks.numpy(), width=case["W"], height=case["H"],
                   render=render.cpu().numpy(), alpha=alpha.cpu().numpy(), radii=info["radii"].cpu().numpy(),
                   isect_offsets=info["isect_offsets"].cpu().numpy(), flatten_ids=info["flatten_ids"].cpu().numpy(),
                   camera_ids=info["camera_ids"].cpu().numpy(), gaussian_ids=info["gaussian_ids"].cpu().numpy(),
                   truth=truth.numpy(), ssim=np.float64(s), gsplat_version=str(gsplat.__version__))
        path = os.path.join(ROOT, "tests", "golden", f"raster_{case['tag']}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
