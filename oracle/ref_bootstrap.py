"""Imports the UNMODIFIED reference (read-only mount /root/reference) in the build container.

Used only by oracle/gen_golden.py (fixture generation) and by tests that are
skipped when /root/reference is absent (it does not exist on the GPU box).
The only shim is `roma.unitquat_to_rotmat` (XYZW), the single roma function
the hot path calls (starster/reconstruct.py:229); roma itself is not installed.
"""
import os
import sys
import types

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "mast3r", "mast3r"))


def _roma_shim():
    import torch

    def unitquat_to_rotmat(q):
        x, y, z, w = q.unbind(-1)
        R = torch.stack([
            1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1)
        return R.reshape(q.shape[:-1] + (3, 3))

    def rotmat_to_unitquat(R):
        raise NotImplementedError("dead code in the reference hot path (reconstruct.py:183-187)")

    m = types.ModuleType("roma")
    m.unitquat_to_rotmat = unitquat_to_rotmat
    m.rotmat_to_unitquat = rotmat_to_unitquat
    return m


def bootstrap():
    """Returns (fast_nn module, sparse_ga module) of the reference."""
    if not available():
        raise RuntimeError("reference tree not mounted at /root/reference")
    for p in (os.path.join(REF, "mast3r", "dust3r"), os.path.join(REF, "mast3r")):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules.setdefault("roma", _roma_shim())
    import mast3r.fast_nn as fast_nn
    from mast3r.cloud_opt import sparse_ga
    return fast_nn, sparse_ga
