"""Repeatability of the reconstruct leg of bench.py (stage times of three consecutive passes)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
for i in range(3):
    out = bench.reconstruct_leg(dev, cpu_legs=False)
    print(json.dumps({"pass": i, "seconds": out["seconds"], "stages_s": out["stages_s"],
                      "align_us_per_iteration": out["align"]["us_per_iteration"],
                      "smooth_ms_per_pair": out["match_smooth_descriptors"]["ms_per_pair"]}))
