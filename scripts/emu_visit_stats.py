"""Development aid (CPU only): visit statistics of the blend backward on a realistic frame, measured by running the
kernel source on the SIMT emulator (tests/host/simt_emu.h) over a sample of tiles of ONE view of the headline workload
(BASELINE.json configs[1]: 200 k Gaussians, 512 x 512, 3e-3 initial scale; the oracle projects and bins the view).
Prints how many (warp, Gaussian) visits a tile needs, how many of them have no contributing lane, and how many
contributing lanes the others have - the quantities that decide what the record-queue variant and a tighter cull buy.

  python scripts/emu_visit_stats.py [n_gaussians] [tile_stride]"""
import ctypes
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gs_oracle as go  # noqa: E402
from starst3r_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    stride = int(sys.argv[2]) if len(sys.argv) > 2 else 37
    W = H = 512
    so = os.path.join(tempfile.mkdtemp(), "libraster_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                    os.path.join(ROOT, "tests", "host", "raster_emu_host.cpp"), "-o", so], check=True)
    emu = ctypes.CDLL(so)
    sp = synth.random_splats(n, seed=0, scale_mode="init")
    viewmats, Ks = synth.look_at_cameras(8, W, H)
    viewmats, Ks = viewmats[:1], Ks[:1]
    radii, means2d, depths, conics = go.project(sp["means"], sp["quats"], sp["scales"], viewmats, Ks, W, H)
    rgb = go.sh_colors(sp["means"], torch.linalg.inv(viewmats)[:, :3, 3], sp["shN"])
    tw = th = W // 16
    b = go.isect_tiles(means2d, radii, depths, 16, tw, th)
    offsets = go.isect_offset_encode(b["isect_ids"], 1, tw, th, b["tile_n_bits"]).reshape(-1).astype(np.int32)
    cam, gau = torch.from_numpy(b["camera_ids"]), torch.from_numpy(b["gaussian_ids"])
    E = len(gau)
    A = np.zeros((E, 4), np.float32)
    B = np.zeros((E, 4), np.float32)
    col = np.zeros((E, 4), np.float32)
    A[:, :2] = means2d[cam, gau].numpy()
    A[:, 2] = torch.sigmoid(sp["opacities"][gau]).numpy() if sp["opacities"].min() < 0 else sp["opacities"][gau].numpy()
    A[:, 3] = depths[cam, gau].numpy()
    B[:, :3] = conics[cam, gau].numpy()
    col[:, :3] = rgb[cam, gau].numpy()
    flatten = np.r_[b["flatten_ids"].astype(np.uint32), np.uint32(0)]
    n_isect = np.asarray([len(b["flatten_ids"])], np.int32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    render = np.zeros((1, H, W, 3), np.float32)
    alphas = np.zeros((1, H, W), np.float32)
    last_ids = np.zeros((1, H, W), np.int32)
    n_blend = np.zeros(1, np.uint64)
    emu.emu_set_tile_stride(stride)
    assert emu.emu_raster_fwd(P(offsets), P(n_isect), P(flatten), P(A), P(B), P(col), 1, W, H, P(render), P(alphas),
                              P(last_ids), P(n_blend)) == 0
    rng = np.random.default_rng(0)
    v_render = rng.standard_normal((1, H, W, 3)).astype(np.float32)
    vA, vB, vC = np.zeros_like(A), np.zeros_like(B), np.zeros_like(col)
    counts = (ctypes.c_long * 16)()
    emu.emu_counts(counts, 1)
    assert emu.emu_raster_bwd(1, P(offsets), P(n_isect), P(flatten), P(A), P(B), P(col), 1, W, H, P(alphas), P(last_ids),
                              P(v_render), None, P(vA), P(vB), P(vC)) == 0
    emu.emu_counts(counts, 1)
    tiles = len(range(0, tw * th, stride))
    per_tile = np.diff(np.r_[offsets, n_isect])[::stride]
    visits, contrib, pairs = counts[3] / 32, counts[4] / 32, counts[5]
    dense, queued = counts[0] / 32, counts[2] / 32
    print(f"{tiles} tiles sampled, {per_tile.mean():.0f} intersections per tile ({n_isect[0]} in the view), "
          f"{int(n_blend[0]) / tiles / 256:.1f} blends per pixel")
    print(f"per tile: {visits / tiles:.0f} (warp, Gaussian) visits = {visits / max(per_tile.sum(), 1):.2f} per intersection; "
          f"{100 * (1 - contrib / max(visits, 1)):.0f} % of them without a contributing lane; "
          f"{pairs / max(contrib, 1):.2f} contributing lanes per contributing visit "
          f"({100 * dense / max(contrib, 1):.0f} % dense, {100 * queued / max(contrib, 1):.0f} % queued)")
    empty = visits - contrib
    default = empty * 50 + contrib * 190
    queue = empty * 50 + dense * 190 + queued * 110 + max(pairs - dense * 20, 0) * 5.5
    print(f"static-count model (DESIGN.md §10): default {default / tiles:.0f}, queue variant {queue / tiles:.0f} warp "
          f"instructions per tile -> x{default / queue:.2f}")
    # variant 2 (fragment pool)
    vA[:], vB[:], vC[:] = 0, 0, 0
    assert emu.emu_raster_bwd(2, P(offsets), P(n_isect), P(flatten), P(A), P(B), P(col), 1, W, H, P(alphas), P(last_ids),
                              P(v_render), None, P(vA), P(vB), P(vC)) == 0
    emu.emu_counts(counts, 1)
    tested, visits2, dense2, frags = counts[8], counts[9] / 32, counts[6] / 32, counts[7]
    print(f"fragment pool: {tested / max(per_tile.sum(), 1):.1f} box pixels tested per intersection in phase A, "
          f"{visits2 / tiles:.0f} phase-B visits per tile ({dense2 / tiles:.1f} dense), {frags / tiles:.0f} pool fragments per tile")


if __name__ == "__main__":
    main()
