// TEST-ONLY: the dense-geometry kernels of starst3r_b200/csrc/align_dense.cu on the SIMT emulator (simt_emu.h): the
// single-CTA Weiszfeld focal kernel and its thread-block-cluster variant (8 CTAs per image exchanging partial sums
// through distributed shared memory), canonical_view and clean_pointcloud.  Never linked into the product library.
#include "simt_emu.h"
#define ST3R_HOST_EMU 1
#include "../../starst3r_b200/csrc/align_dense.cu"

extern "C" {

// variant 0: focal_weiszfeld_kernel (one CTA per image), 1: focal_weiszfeld_cluster_kernel (one cluster per image)
int emu_focal_weiszfeld(int variant, const float* canon, int n_img, int H, int W, float min_focal, float max_focal,
                        float* out) {
  emu::g_blockDim = dim3(1024, 1, 1);
  for (int img = 0; img < n_img; ++img) {
    bool ok;
    if (variant == 1) {
      emu::g_gridDim = dim3(n_img * WZ_CLUSTER, 1, 1);
      emu::g_blockIdx = uint3{(unsigned)(img * WZ_CLUSTER), 0, 0};
      ok = emu::run_cluster(WZ_CLUSTER, 1024, [&]() { focal_weiszfeld_cluster_kernel(canon, H, W, min_focal, max_focal, out); },
                            64 * 1024);
    } else {
      emu::g_gridDim = dim3(n_img, 1, 1);
      emu::g_blockIdx = uint3{(unsigned)img, 0, 0};
      ok = emu::run_cluster(1, 1024, [&]() { focal_weiszfeld_kernel(canon, H, W, min_focal, max_focal, out); }, 64 * 1024);
    }
    if (!ok) return -1;
  }
  return 0;
}

int emu_canonical_view(const float* ptmaps, const float* confs, int P, int H, int W, int S, float* canon, float* canon2,
                       float* cconf) {
  emu::g_blockDim = dim3(256, 1, 1);
  emu::g_gridDim = dim3((W + 255) / 256, H, 1);
  for (int y = 0; y < H; ++y)
    for (int bx = 0; bx < (W + 255) / 256; ++bx) {
      emu::g_blockIdx = uint3{(unsigned)bx, (unsigned)y, 0};
      if (!emu::run_cta(256, [&]() { canonical_view_kernel(ptmaps, confs, P, H, W, S, canon, canon2, cconf); })) return -1;
    }
  return 0;
}
}
