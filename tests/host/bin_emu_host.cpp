// TEST-ONLY: the fused tile binning of starst3r_b200/csrc/gs_bin.cu (counting sort by (camera, tile) with privatised
// shared-memory counters + per-tile flip-bitonic sort) on the SIMT emulator (simt_emu.h), in the launch sequence of
// st3r_gs_bin_tiles; the exclusive scan between the passes is done on the host here.  Never linked into the product.
#include "simt_emu.h"
#define ST3R_HOST_EMU 1
static int32_t g_dyn_smem_i32[2 * 10240];
#define ST3R_DYN_SMEM_I32(name) int32_t* name = g_dyn_smem_i32
#include "../../starst3r_b200/csrc/gs_bin.cu"

extern "C" {

// keys / vals: [n_cap]; offsets: [C * tiles]; returns the intersection total, or -1 on an emulator deadlock.
int emu_bin_tiles(const int32_t* radii, const float* geomA, int N, int C, int W, int H, int tile_n_bits, int32_t* offsets,
                  uint64_t* keys, uint32_t* vals, int n_cap, int use_smem, int reg_sort) {
  const int tile_size = 16;
  const int tile_w = (W + tile_size - 1) / tile_size, tile_h = (H + tile_size - 1) / tile_size;
  const int n_tiles = tile_w * tile_h, n_cells = C * n_tiles;
  std::vector<int32_t> counts(n_cells, 0), cursor(n_cells, 0);
  std::vector<uint64_t> pairs(n_cap > 0 ? n_cap : 1, 0);
  const float4* gA = (const float4*)geomA;
  const int gx = (N + MIN_ENTRIES_PER_CTA - 1) / MIN_ENTRIES_PER_CTA;
  auto grid2 = [&](const std::function<void()>& body) {
    emu::g_blockDim = dim3(BIN_THREADS, 1, 1);
    emu::g_gridDim = dim3(gx, C, 1);
    for (int c = 0; c < C; ++c)
      for (int x = 0; x < gx; ++x) {
        emu::g_blockIdx = uint3{(unsigned)x, (unsigned)c, 0};
        if (!emu::run_cta(BIN_THREADS, body)) return false;
      }
    return true;
  };
  if (!grid2([&]() { tile_hist_kernel(radii, gA, N, tile_size, tile_w, tile_h, counts.data(), use_smem, MIN_ENTRIES_PER_CTA); })) return -1;
  int32_t total = 0;
  for (int i = 0; i < n_cells; ++i) { offsets[i] = total; total += counts[i]; }
  if (!grid2([&]() { tile_emit_kernel(radii, gA, N, tile_size, tile_w, tile_h, offsets, cursor.data(), pairs.data(), n_cap, use_smem, MIN_ENTRIES_PER_CTA); }))
    return -1;
  emu::g_blockDim = dim3(SORT_THREADS, 1, 1);
  emu::g_gridDim = dim3(n_cells, 1, 1);
  for (int cell = 0; cell < n_cells; ++cell) {
    emu::g_blockIdx = uint3{(unsigned)cell, 0, 0};
    if (!emu::run_cta(SORT_THREADS, [&]() { tile_sort_kernel(offsets, &total, n_cells, n_tiles, tile_n_bits, pairs.data(), keys, vals, n_cap, reg_sort); }))
      return -1;
  }
  return total;
}
}
