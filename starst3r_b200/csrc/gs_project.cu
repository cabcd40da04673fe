// 3DGS front end: projection + SH colour + tile counting, tile-intersection key emission,
// per-tile offsets.  (gsplat fully_fused_projection / spherical_harmonics / isect_tiles /
// isect_offset_encode; called from starster/gs.py:76-87.)
//
// This translation unit is compiled with -fmad=false: every multiply/add below is a single
// IEEE operation in the order written in gs_math.cuh, which makes radii, tile ranges and the
// depth bits inside the sort keys bit-identical to the CPU restatement.
//
// Layout in HBM (dense, entry e = camera * N + gaussian, gaussian fastest => coalesced):
//   radii   int32  [C*N]          0 = culled
//   geomA   float4 [C*N]          (mean2d.x, mean2d.y, opacity, depth)
//   geomB   float4 [C*N]          (conic a, b, c, unused)
//   rgb     float4 [C*N]          (r, g, b, unused)
//   tiles   int32  [C*N]          tiles touched (input of the exclusive scan)
// Algorithmic bytes: 92 B read per (Gaussian, view), 44 B written per visible one.
#include "common.cuh"
#include "gs.cuh"
#include "gs_math.cuh"

namespace {

// A thread projects ONE Gaussian into up to PROJ_CPT consecutive cameras: the 92 B of parameters are read once per
// group of cameras instead of once per camera (with a thread per (Gaussian, camera) the kernel moved 147 MB of parameter
// re-reads through L2 at configs[1], more than the 90 MB it writes), and what does not depend on the camera - the 3-D
// covariance - is computed once (same operations in the same order: identical bits).
#ifndef ST3R_PROJ_CPT
#define ST3R_PROJ_CPT 4
#endif
constexpr int PROJ_CPT = ST3R_PROJ_CPT;

__global__ void __launch_bounds__(256)
gs_project_kernel(const float* __restrict__ means, const float* __restrict__ quats,
                  const float* __restrict__ scales, const float* __restrict__ opacities,
                  const float* __restrict__ shN, int sh_stride, const GsCam* __restrict__ cams, int N, int C,
                  float W, float H, int tile_size, int tile_w, int tile_h, float eps2d, float near_plane,
                  float far_plane, float radius_clip, int32_t* __restrict__ radii, float4* __restrict__ geomA,
                  float4* __restrict__ geomB, float4* __restrict__ rgb, int32_t* __restrict__ tiles) {
  __shared__ GsCam cam_s[PROJ_CPT];
  const int c0 = blockIdx.y * PROJ_CPT, nc = min(PROJ_CPT, C - c0);
  constexpr int CAMF = (int)(sizeof(GsCam) / 4);
  for (int i = threadIdx.x; i < nc * CAMF; i += blockDim.x)
    reinterpret_cast<float*>(cam_s)[i] = reinterpret_cast<const float*>(cams + c0)[i];
  __syncthreads();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  float mean[3] = {means[3 * g], means[3 * g + 1], means[3 * g + 2]};
  float4 q4 = reinterpret_cast<const float4*>(quats)[g];
  float quat[4] = {q4.x, q4.y, q4.z, q4.w};
  float scale[3] = {scales[3 * g], scales[3 * g + 1], scales[3 * g + 2]};
  const float opac = opacities[g];
  float shl[12];
  {
    const float4* sh4 = reinterpret_cast<const float4*>(shN + (size_t)g * sh_stride);
    float4 s0 = sh4[0], s1 = sh4[1], s2 = sh4[2];
    shl[0] = s0.x; shl[1] = s0.y; shl[2] = s0.z; shl[3] = s0.w; shl[4] = s1.x; shl[5] = s1.y; shl[6] = s1.z;
    shl[7] = s1.w; shl[8] = s2.x; shl[9] = s2.y; shl[10] = s2.z; shl[11] = s2.w;
  }
#pragma unroll 1
  for (int ci = 0; ci < nc; ++ci) {
    const GsCam& cam = cam_s[ci];
    const size_t e = (size_t)(c0 + ci) * N + g;
    GsProj o;
    GsProjTmp t;
    bool vis = gs_project(mean, quat, scale, cam, W, H, eps2d, near_plane, far_plane, radius_clip, o, t);
    int ntiles = 0;
    if (vis) {
      const float ts = (float)tile_size;
      float txc = o.m2x / ts, tyc = o.m2y / ts, tr = (float)o.radius / ts;
      int x0 = min(max(0, (int)floorf(txc - tr)), tile_w), x1 = min(max(0, (int)ceilf(txc + tr)), tile_w);
      int y0 = min(max(0, (int)floorf(tyc - tr)), tile_h), y1 = min(max(0, (int)ceilf(tyc + tr)), tile_h);
      ntiles = (y1 - y0) * (x1 - x0);
      float col[3];
      gs_sh_color(mean, cam.pos, shl, col, nullptr, nullptr, nullptr);
      geomA[e] = make_float4(o.m2x, o.m2y, opac, o.depth);
      geomB[e] = make_float4(o.ca, o.cb, o.cc, 0.f);
      rgb[e] = make_float4(col[0], col[1], col[2], 0.f);
    }
    radii[e] = vis ? o.radius : 0;
    tiles[e] = ntiles;
  }
}

// One thread per entry writes its (key, value) pairs: key = cam << (32 + tile_bits) | tile << 32 | depth bits.
__global__ void __launch_bounds__(256)
gs_isect_kernel(const int32_t* __restrict__ radii, const float4* __restrict__ geomA,
                const int32_t* __restrict__ cum_tiles, int N, int C, int tile_size, int tile_w, int tile_h,
                int tile_n_bits, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int n_cap) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)C * N) return;
  const int r = radii[e];
  if (r <= 0) return;
  const int c = (int)(e / N);
  float4 a = geomA[e];
  const float ts = (float)tile_size;
  float txc = a.x / ts, tyc = a.y / ts, tr = (float)r / ts;
  int x0 = min(max(0, (int)floorf(txc - tr)), tile_w), x1 = min(max(0, (int)ceilf(txc + tr)), tile_w);
  int y0 = min(max(0, (int)floorf(tyc - tr)), tile_h), y1 = min(max(0, (int)ceilf(tyc + tr)), tile_h);
  uint64_t hi = (uint64_t)c << tile_n_bits;
  uint64_t depth_bits = (uint64_t)__float_as_uint(a.w);
  int pos = cum_tiles[e];
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      if (pos < n_cap) {
        keys[pos] = ((hi | (uint64_t)(y * tile_w + x)) << 32) | depth_bits;
        vals[pos] = (uint32_t)e;
      }
      ++pos;
    }
}

// offsets[t] = first sorted position whose (camera, tile) >= t  (t = camera * tiles + tile).
__global__ void __launch_bounds__(256)
gs_offsets_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ n_ptr, int n_cap, int C,
                  int n_tiles, int tile_n_bits, int32_t* __restrict__ offsets) {
  const int n = min(*n_ptr, n_cap);
  const int total = C * n_tiles;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) {
    if (i < total) offsets[i] = 0;
    return;
  }
  if (i >= n) return;
  const uint64_t tmask = (1ull << tile_n_bits) - 1;
  uint64_t k = keys[i] >> 32;
  int cur = (int)(k >> tile_n_bits) * n_tiles + (int)(k & tmask);
  if (i == 0) {
    for (int t = 0; t <= cur; ++t) offsets[t] = 0;
  } else {
    uint64_t kp = keys[i - 1] >> 32;
    int prev = (int)(kp >> tile_n_bits) * n_tiles + (int)(kp & tmask);
    for (int t = prev + 1; t <= cur; ++t) offsets[t] = i;
  }
  if (i == n - 1)
    for (int t = cur + 1; t < total; ++t) offsets[t] = n;
}

}  // namespace

int gs_tile_bits(int n_tiles) {
  int b = 0;
  while ((1 << (b + 1)) <= n_tiles) ++b;  // floor(log2(n_tiles))
  return b + 1;
}

extern "C" {

int st3r_gs_project(const float* means, const float* quats, const float* scales, const float* opacities,
                    const float* shN, int sh_coeffs, const float* cams, int N, int C, int width, int height,
                    int tile_size, float eps2d, float near_plane, float far_plane, float radius_clip,
                    int32_t* radii, float* geomA, float* geomB, float* rgb, int32_t* tiles, cudaStream_t stream) {
  ST3R_CHECK_ARG(N >= 0 && C >= 0 && width > 0 && height > 0 && tile_size > 0, "st3r_gs_project: bad sizes");
  ST3R_CHECK_ARG(sh_coeffs >= 4, "st3r_gs_project: need at least 4 SH coefficients per Gaussian (degree 1)");
  if (N == 0 || C == 0) return ST3R_OK;
  ST3R_CHECK_ARG(means && quats && scales && opacities && shN && cams && radii && geomA && geomB && rgb && tiles,
                 "st3r_gs_project: null pointer");
  ST3R_CHECK_ARG(((uintptr_t)quats % 16) == 0 && ((uintptr_t)shN % 16) == 0 && (sh_coeffs * 3) % 4 == 0,
                 "st3r_gs_project: quats / shN must be 16-byte aligned");
  const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
  dim3 grid((N + 255) / 256, (C + PROJ_CPT - 1) / PROJ_CPT);
  gs_project_kernel<<<grid, 256, 0, stream>>>(means, quats, scales, opacities, shN, sh_coeffs * 3,
                                              reinterpret_cast<const GsCam*>(cams), N, C, (float)width, (float)height,
                                              tile_size, tile_w, tile_h, eps2d, near_plane, far_plane, radius_clip,
                                              radii, reinterpret_cast<float4*>(geomA), reinterpret_cast<float4*>(geomB),
                                              reinterpret_cast<float4*>(rgb), tiles);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_gs_cam_floats(void) { return (int)(sizeof(GsCam) / sizeof(float)); }

int st3r_gs_isect(const int32_t* radii, const float* geomA, const int32_t* cum_tiles, int N, int C, int width,
                  int height, int tile_size, uint64_t* keys, uint32_t* vals, int n_cap, cudaStream_t stream) {
  if (N == 0 || C == 0 || n_cap == 0) return ST3R_OK;
  ST3R_CHECK_ARG(radii && geomA && cum_tiles && keys && vals, "st3r_gs_isect: null pointer");
  const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
  size_t total = (size_t)C * N;
  gs_isect_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      radii, reinterpret_cast<const float4*>(geomA), cum_tiles, N, C, tile_size, tile_w, tile_h,
      gs_tile_bits(tile_w * tile_h), keys, vals, n_cap);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_gs_sort_bits(int C, int width, int height, int tile_size) {
  const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
  int cam_bits = 0;
  while ((1 << cam_bits) < C) ++cam_bits;
  return 32 + gs_tile_bits(tile_w * tile_h) + cam_bits;
}

int st3r_gs_offsets(const uint64_t* keys, const int32_t* n_isect, int n_cap, int C, int width, int height,
                    int tile_size, int32_t* offsets, cudaStream_t stream) {
  ST3R_CHECK_ARG(n_isect && offsets, "st3r_gs_offsets: null pointer");
  const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
  int n_tiles = tile_w * tile_h;
  int work = max(n_cap, C * n_tiles);
  gs_offsets_kernel<<<(work + 255) / 256, 256, 0, stream>>>(keys, n_isect, n_cap, C, n_tiles, gs_tile_bits(n_tiles),
                                                           offsets);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

}  // extern "C"
