// TEST-ONLY: the blend kernels of starst3r_b200/csrc/gs_raster.cu (fragment-pool pair and visit-list pair)
// compiled for the host and executed thread by thread by the SIMT emulator in simt_emu.h, so that the CPU test-suite
// runs the kernels' own source - indexing, visit lists, fragment pools, block scans, barriers - against the oracle.
// Never linked into the product library.
#include "simt_emu.h"
#define ST3R_HOST_EMU 1
static long g_emu_count[16];     // per-lane hits of the marked code paths of the kernels
#define ST3R_EMU_COUNT(i) (++g_emu_count[i])
#include "../../starst3r_b200/csrc/gs_raster.cu"

static int g_tile_stride = 1;   // > 1: only every g_tile_stride-th tile runs (statistics on samples of large frames)

namespace {
template <typename K>
int run_grid(int tiles, int C, K&& kernel_body) {
  emu::g_blockDim = dim3(BLOCK, 1, 1);
  emu::g_gridDim = dim3(tiles, C, 1);
  for (int c = 0; c < C; ++c)
    for (int t = 0; t < tiles; ++t) {
      if (t % g_tile_stride) continue;
      emu::g_blockIdx = uint3{(unsigned)t, (unsigned)c, 0};
      if (!emu::run_cta(BLOCK, kernel_body)) return -1;
    }
  return 0;
}
}  // namespace

extern "C" {

void emu_set_tile_stride(int stride) { g_tile_stride = stride > 0 ? stride : 1; }

void emu_counts(long* out16, int reset) {
  for (int i = 0; i < 16; ++i) { out16[i] = g_emu_count[i]; if (reset) g_emu_count[i] = 0; }
}

// variant 0: fragment-pool kernels (default), 1: visit-list kernels
int emu_raster_fwd(int variant, const int32_t* offsets, const int32_t* n_isect, const uint32_t* flatten, const float* geomA,
                   const float* geomB, const float* rgb, int C, int W, int H, float* render, float* alphas,
                   int32_t* last_ids, unsigned long long* n_blend) {
  const int tile_w = (W + TILE - 1) / TILE, tile_h = (H + TILE - 1) / TILE;
  return run_grid(tile_w * tile_h, C, [&]() {
    if (variant == 0)
      raster_fwd_pool_kernel(offsets, n_isect, flatten, (const float4*)geomA, (const float4*)geomB, (const float4*)rgb, C,
                             W, H, tile_w, tile_h, render, alphas, last_ids, n_blend);
    else
      raster_fwd_kernel(offsets, n_isect, flatten, (const float4*)geomA, (const float4*)geomB, (const float4*)rgb, C, W, H,
                        tile_w, tile_h, render, alphas, last_ids, n_blend);
  });
}

int emu_raster_bwd(int variant, const int32_t* offsets, const int32_t* n_isect, const uint32_t* flatten,
                   const float* geomA, const float* geomB, const float* rgb, int C, int W, int H, const float* alphas,
                   const int32_t* last_ids, const float* v_render, const float* v_alphas, float* v_geomA, float* v_geomB,
                   float* v_rgb) {
  const int tile_w = (W + TILE - 1) / TILE, tile_h = (H + TILE - 1) / TILE;
  return run_grid(tile_w * tile_h, C, [&]() {
    if (variant == 0)
      raster_bwd_pool_kernel(offsets, n_isect, flatten, (const float4*)geomA, (const float4*)geomB, (const float4*)rgb, C,
                             W, H, tile_w, tile_h, alphas, last_ids, v_render, v_alphas, (float4*)v_geomA,
                             (float4*)v_geomB, (float4*)v_rgb);
    else
      raster_bwd_kernel(offsets, n_isect, flatten, (const float4*)geomA, (const float4*)geomB, (const float4*)rgb, C, W, H,
                        tile_w, tile_h, alphas, last_ids, v_render, v_alphas, (float4*)v_geomA, (float4*)v_geomB,
                        (float4*)v_rgb);
  });
}
}
