"""CPU: the fused tile binning (starst3r_b200/csrc/gs_bin.cu: tile_hist / tile_emit / tile_sort, i.e. gsplat's
isect_tiles + SortPairs + isect_offset_encode) run from its own source on the SIMT emulator
and compared BIT FOR BIT with the oracle's integer path: isect_ids, flatten_ids, isect_offsets - including tiles
whose lists exceed the 4096-element shared-memory sort and the global-atomics fallback of the counting passes."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import gs_oracle as go
from starst3r_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("emu") / "libbin_emu.so"
    src = os.path.join(ROOT, "tests", "host", "bin_emu_host.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("reg_sort", [2, 1, 0])
@pytest.mark.parametrize("n,C,W,H,scale_mult,use_smem", [(600, 2, 80, 48, 8.0, 1), (600, 2, 80, 48, 8.0, 0),
                                                         (2500, 1, 96, 64, 6.0, 1), (6000, 1, 160, 112, 8.0, 1),
                                                         (9000, 1, 40, 40, 60.0, 1), (2501, 1, 96, 64, 6.0, 1),
                                                         (2502, 1, 96, 64, 6.0, 1)])
def test_fused_binning_bit_exact_on_the_emulator(emu, n, C, W, H, scale_mult, use_smem, reg_sort):
    """reg_sort = 2: segments up to 2048 pairs are sorted in registers through 32-bit surrogate keys and repaired by
    odd-even passes (n = 2501: depths quantised to 1/64, runs of equal leading bits AND exact ties that the entry index
    must break; n = 2502: four distinct depths, which exhausts the passes and takes the fallback); reg_sort = 1: the
    64-bit register-resident network (1 / 2 / 4 / 8 elements per thread); longer segments and reg_sort = 0 the
    shared-memory / in-place network."""
    sp = synth.random_splats(n, seed=1, scale_mode="rand")
    sp["scales"] = sp["scales"] * scale_mult
    viewmats, Ks = synth.look_at_cameras(C, W, H)
    radii, means2d, depths, conics = go.project(sp["means"], sp["quats"], sp["scales"], viewmats, Ks, W, H)
    if n == 2501:
        depths = torch.round(depths * 64.0) / 64.0 + 1.0 / 128.0
    if n == 2502:
        depths = torch.round(depths * 0.5) * 2.0 + 1.0
    tw, th = (W + 15) // 16, (H + 15) // 16
    b = go.isect_tiles(means2d, radii, depths, 16, tw, th)
    want_offsets = go.isect_offset_encode(b["isect_ids"], C, tw, th, b["tile_n_bits"]).reshape(-1)
    n_isect = len(b["isect_ids"])
    per_cell = np.diff(np.r_[want_offsets, n_isect])
    if n == 9000:
        assert per_cell.max() > 4096                  # a segment that is sorted in place in global memory
    # between them the cases cover every size class of the register-resident sort (1 / 2 / 4 / 8 elements per thread)
    classes = {600: ((1, 256), (257, 512)), 2500: ((257, 512), (513, 1024), (1025, 2048)), 6000: ((1025, 2048), (2049, 4096))}
    for lo, hi in classes.get(n, ()):
        assert ((per_cell >= lo) & (per_cell <= hi)).any(), (lo, hi, sorted(per_cell))
    geomA = np.zeros((C * n, 4), np.float32)
    geomA[:, :2] = means2d.reshape(-1, 2).numpy()
    geomA[:, 3] = depths.reshape(-1).numpy()
    rad = np.ascontiguousarray(radii.reshape(-1).numpy().astype(np.int32))
    offsets = np.zeros(C * tw * th, np.int32)
    keys = np.zeros(max(n_isect, 1), np.uint64)
    vals = np.zeros(max(n_isect, 1), np.uint32)
    total = emu.emu_bin_tiles(P(rad), P(geomA), n, C, W, H, int(b["tile_n_bits"]), P(offsets), P(keys), P(vals), n_isect,
                              use_smem, reg_sort)
    assert total == n_isect
    assert np.array_equal(offsets, want_offsets.astype(np.int32))
    assert np.array_equal(keys[:n_isect].astype(np.int64), b["isect_ids"])
    dense_of_packed = b["camera_ids"] * n + b["gaussian_ids"]
    assert np.array_equal(vals[:n_isect].astype(np.int64), dense_of_packed[b["flatten_ids"]])
