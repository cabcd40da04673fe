// Projection + SH backward (gsplat fully_fused_projection_bwd / spherical_harmonics_bwd, SURVEY.md
// Appendix A.6 last paragraph), plus the gradients of the two regularisers of starster/gs.py:132-134.
//
// A CTA handles 32 consecutive Gaussians: warp w takes the cameras w, w + 8, ... (one thread per (camera,
// Gaussian) entry, coalesced over Gaussians), re-derives the forward intermediates from the 92 B of parameters and
// turns the per-entry screen-space gradients written by the blend backward into gradients of means / quats / scales /
// opacities / SH coefficients; the warps' partial sums meet in shared memory and are added in warp order - no
// atomics, fixed summation order => deterministic - and the 736 output floats of the CTA are written coalesced.
// (The first version looped over the cameras inside one thread per Gaussian: 91 us at configs[1], latency-bound on the
// dependent radii -> gradient loads of the serial camera loop; the algorithmic traffic is ~120 MB, 18 us.)
// Compiled with -fmad=false like gs_project.cu so the recomputed forward is the same bit pattern.
#include "common.cuh"
#include "gs.cuh"
#include "gs_math.cuh"

namespace {

constexpr int PB_GAUSS = 32;      // Gaussians per CTA (one per lane)
constexpr int PB_WARPS = 8;       // camera-parallel warps per CTA
constexpr int PB_NG = 23;         // gradient values per Gaussian: means 3, quats 4, scales 3, opacity 1, SH 12

__global__ void __launch_bounds__(PB_GAUSS * PB_WARPS)
gs_project_bwd_kernel(const float* __restrict__ means, const float* __restrict__ quats,
                      const float* __restrict__ scales, const float* __restrict__ opacities,
                      const float* __restrict__ shN, int sh_stride, const GsCam* __restrict__ cams, int N, int C,
                      float W, float H, float eps2d, float near_plane, float far_plane, float radius_clip,
                      const int32_t* __restrict__ radii, const float4* __restrict__ v_geomA,
                      const float4* __restrict__ v_geomB, const float4* __restrict__ v_rgb, float reg_opac,
                      float reg_scale, float* __restrict__ v_means, float* __restrict__ v_quats,
                      float* __restrict__ v_scales, float* __restrict__ v_opacities, float* __restrict__ v_sh,
                      float* __restrict__ reg_sums) {
  __shared__ float part[PB_WARPS][PB_GAUSS * PB_NG];     // [warp][Gaussian][value]: odd pitch, conflict-free both ways
  __shared__ float red[2][PB_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g0 = blockIdx.x * PB_GAUSS, g = g0 + lane;
  float acc[PB_NG];      // vm 0..2, vq 3..6, vs 7..9, vo 10, vsh 11..22
#pragma unroll
  for (int k = 0; k < PB_NG; ++k) acc[k] = 0.f;
  if (g < N && warp < C) {
    float mean[3] = {means[3 * g], means[3 * g + 1], means[3 * g + 2]};
    float4 q4 = reinterpret_cast<const float4*>(quats)[g];
    float quat[4] = {q4.x, q4.y, q4.z, q4.w};
    float scale[3] = {scales[3 * g], scales[3 * g + 1], scales[3 * g + 2]};
    float sh[12];
    const float4* sh4 = reinterpret_cast<const float4*>(shN + (size_t)g * sh_stride);
    float4 s0 = sh4[0], s1 = sh4[1], s2 = sh4[2];
    sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
    sh[8] = s2.x; sh[9] = s2.y; sh[10] = s2.z; sh[11] = s2.w;
    for (int c = warp; c < C; c += nwarps) {
      const size_t e = (size_t)c * N + g;
      if (radii[e] <= 0) continue;
      const float4 gA = v_geomA[e], gB = v_geomB[e], gC = v_rgb[e];
      GsCam cam = cams[c];
      GsProj o;
      GsProjTmp t;
      if (!gs_project(mean, quat, scale, cam, W, H, eps2d, near_plane, far_plane, radius_clip, o, t)) continue;
      acc[10] += gA.z;
      float col[3], raw[3], dirn[3], inv_len;
      gs_sh_color(mean, cam.pos, sh, col, raw, dirn, &inv_len);
      const float vrgb[3] = {gC.x, gC.y, gC.z};
      gs_sh_color_vjp(sh, raw, dirn, inv_len, vrgb, acc + 11, acc + 0);
      const float vcon[3] = {gB.x, gB.y, gB.z};
      gs_project_vjp(scale, cam, o, t, gA.x, gA.y, vcon, acc + 0, acc + 3, acc + 7);
    }
  }
#pragma unroll
  for (int k = 0; k < PB_NG; ++k) part[warp][lane * PB_NG + k] = acc[k];
  __syncthreads();
  // Output pass: the CTA's 32 Gaussians own contiguous chunks of the five gradient arrays (96 + 128 + 96 + 32 + 384
  // floats); thread f adds the warps' partial sums of one output float in warp order.
  const int nvalid = min(PB_GAUSS, N - g0);
  float sig_sum = 0.f, exp_sum = 0.f;
  for (int f = threadIdx.x; f < PB_GAUSS * PB_NG; f += blockDim.x) {
    int gl, k;
    float* dst;
    if (f < 96) { gl = f / 3; k = f - 3 * gl; dst = v_means + (size_t)g0 * 3 + f; }
    else if (f < 224) { const int q = f - 96; gl = q >> 2; k = 3 + (q & 3); dst = v_quats + (size_t)g0 * 4 + q; }
    else if (f < 320) { const int q = f - 224; gl = q / 3; k = 7 + q - 3 * gl; dst = v_scales + (size_t)g0 * 3 + q; }
    else if (f < 352) { gl = f - 320; k = 10; dst = v_opacities + g0 + gl; }
    else { const int q = f - 352; gl = q / 12; k = 11 + q - 12 * gl; dst = v_sh + (size_t)g0 * 12 + q; }
    if (gl >= nvalid) continue;
    float v = part[0][gl * PB_NG + k];
    for (int w = 1; w < nwarps; ++w) v += part[w][gl * PB_NG + k];
    // regularisers: fac * mean|sigmoid(opacity)| and fac * mean|exp(scale)| summed over the C views
    if (k == 10) {
      const float sg = 1.0f / (1.0f + expf(-opacities[g0 + gl]));
      v += reg_opac * sg * (1.0f - sg);
      sig_sum += sg;
    } else if (k >= 7 && k < 10) {
      const float ex = expf(scales[(size_t)g0 * 3 + (f - 224)]);
      v += reg_scale * ex;
      exp_sum += ex;
    }
    *dst = v;
  }
  if (reg_sums) {
    for (int off = 16; off; off >>= 1) {
      sig_sum += __shfl_xor_sync(0xffffffffu, sig_sum, off);
      exp_sum += __shfl_xor_sync(0xffffffffu, exp_sum, off);
    }
    if (lane == 0) { red[0][warp] = sig_sum; red[1][warp] = exp_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < nwarps; ++w) { a += red[0][w]; b += red[1][w]; }
      atomicAdd(reg_sums, a);
      atomicAdd(reg_sums + 1, b);
    }
  }
}

}  // namespace

extern "C" int st3r_gs_project_bwd(const float* means, const float* quats, const float* scales, const float* opacities,
                                   const float* shN, int sh_coeffs, const float* cams, int N, int C, int width,
                                   int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                                   const int32_t* radii, const float* v_geomA, const float* v_geomB,
                                   const float* v_rgb, float reg_opac, float reg_scale, float* v_means,
                                   float* v_quats, float* v_scales, float* v_opacities, float* v_sh, float* reg_sums,
                                   cudaStream_t stream) {
  ST3R_CHECK_ARG(N >= 0 && C >= 0 && sh_coeffs >= 4, "st3r_gs_project_bwd: bad sizes");
  if (N == 0) return ST3R_OK;
  ST3R_CHECK_ARG(means && quats && scales && opacities && shN && cams && radii && v_geomA && v_geomB && v_rgb &&
                     v_means && v_quats && v_scales && v_opacities && v_sh,
                 "st3r_gs_project_bwd: null pointer");
  const int warps = C < PB_WARPS ? (C < 1 ? 1 : C) : PB_WARPS;
  gs_project_bwd_kernel<<<(N + PB_GAUSS - 1) / PB_GAUSS, 32 * warps, 0, stream>>>(
      means, quats, scales, opacities, shN, sh_coeffs * 3, reinterpret_cast<const GsCam*>(cams), N, C, (float)width,
      (float)height, eps2d, near_plane, far_plane, radius_clip, radii, reinterpret_cast<const float4*>(v_geomA),
      reinterpret_cast<const float4*>(v_geomB), reinterpret_cast<const float4*>(v_rgb), reg_opac, reg_scale, v_means,
      v_quats, v_scales, v_opacities, v_sh, reg_sums);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
