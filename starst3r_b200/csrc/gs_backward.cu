// Projection + SH backward (gsplat fully_fused_projection_bwd / spherical_harmonics_bwd, SURVEY.md
// Appendix A.6 last paragraph), plus the gradients of the two regularisers of starster/gs.py:132-134.
//
// One thread per Gaussian loops over the C cameras (coalesced over Gaussians, no atomics, fixed
// summation order => deterministic), re-derives the forward intermediates from the 92 B of
// parameters and turns the per-(camera, Gaussian) screen-space gradients written by the blend
// backward into gradients of means / quats / scales / opacities / SH coefficients.
// Compiled with -fmad=false like gs_project.cu so the recomputed forward is the same bit pattern.
// (Measured alternative, round 2: a warp per camera with the partial sums meeting in shared memory - 8x the threads,
// no serial camera loop - ran at 141 us against this kernel's 92 us at configs[1] (profiles/r02w_*): the kernel is bound
// by its ~250 non-contracted FMUL / FADD per (camera, Gaussian), not by the latency of the loop, and the parallel
// form re-derives the per-Gaussian part of the projection once per camera.  Dropped.)
#include "common.cuh"
#include "gs.cuh"
#include "gs_math.cuh"

namespace {

// 4 CTAs per SM (128 registers instead of 165) and the next camera's screen-space gradients requested before the current
// camera's are consumed: 0.094 -> 0.079 ms at configs[1] (B200 A/B of 1 / 4 / 5 / 6 CTAs per SM with and without the
// prefetch: profiles/r02ag_libvar.jsonl; 5 and 6 spill and lose).
__global__ void __launch_bounds__(128, 4)
gs_project_bwd_kernel(const float* __restrict__ means, const float* __restrict__ quats,
                      const float* __restrict__ scales, const float* __restrict__ opacities,
                      const float* __restrict__ shN, int sh_stride, const GsCam* __restrict__ cams, int N, int C,
                      float W, float H, float eps2d, float near_plane, float far_plane, float radius_clip,
                      const int32_t* __restrict__ radii, const float4* __restrict__ v_geomA,
                      const float4* __restrict__ v_geomB, const float4* __restrict__ v_rgb, float reg_opac,
                      float reg_scale, float* __restrict__ v_means, float* __restrict__ v_quats,
                      float* __restrict__ v_scales, float* __restrict__ v_opacities, float* __restrict__ v_sh,
                      float* __restrict__ reg_sums) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  float sig_sum = 0.f, exp_sum = 0.f;
  if (g < N) {
    float mean[3] = {means[3 * g], means[3 * g + 1], means[3 * g + 2]};
    float4 q4 = reinterpret_cast<const float4*>(quats)[g];
    float quat[4] = {q4.x, q4.y, q4.z, q4.w};
    float scale[3] = {scales[3 * g], scales[3 * g + 1], scales[3 * g + 2]};
    const float opac = opacities[g];
    float sh[12];
    const float4* sh4 = reinterpret_cast<const float4*>(shN + (size_t)g * sh_stride);
    float4 s0 = sh4[0], s1 = sh4[1], s2 = sh4[2];
    sh[0] = s0.x; sh[1] = s0.y; sh[2] = s0.z; sh[3] = s0.w; sh[4] = s1.x; sh[5] = s1.y; sh[6] = s1.z; sh[7] = s1.w;
    sh[8] = s2.x; sh[9] = s2.y; sh[10] = s2.z; sh[11] = s2.w;

    float vm[3] = {0.f, 0.f, 0.f}, vq[4] = {0.f, 0.f, 0.f, 0.f}, vs[3] = {0.f, 0.f, 0.f}, vo = 0.f;
    float vsh[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) vsh[k] = 0.f;

    // the next camera's screen-space gradients are in flight while this camera's are turned into parameter gradients
    int r_nx = C > 0 ? radii[g] : 0;
    float4 a_nx = make_float4(0.f, 0.f, 0.f, 0.f), b_nx = a_nx, c_nx = a_nx;
    if (C > 0) { a_nx = v_geomA[g]; b_nx = v_geomB[g]; c_nx = v_rgb[g]; }
    for (int c = 0; c < C; ++c) {
      const int r_cur = r_nx;
      const float4 gA = a_nx, gB = b_nx, gC = c_nx;
      if (c + 1 < C) {
        const size_t e1 = (size_t)(c + 1) * N + g;
        r_nx = radii[e1]; a_nx = v_geomA[e1]; b_nx = v_geomB[e1]; c_nx = v_rgb[e1];
      }
      if (r_cur <= 0) continue;
      GsCam cam = cams[c];
      GsProj o;
      GsProjTmp t;
      if (!gs_project(mean, quat, scale, cam, W, H, eps2d, near_plane, far_plane, radius_clip, o, t)) continue;
      vo += gA.z;
      float col[3], raw[3], dirn[3], inv_len;
      gs_sh_color(mean, cam.pos, sh, col, raw, dirn, &inv_len);
      const float vrgb[3] = {gC.x, gC.y, gC.z};
      gs_sh_color_vjp(sh, raw, dirn, inv_len, vrgb, vsh, vm);
      const float vcon[3] = {gB.x, gB.y, gB.z};
      gs_project_vjp(scale, cam, o, t, gA.x, gA.y, vcon, vm, vq, vs);
    }
    // regularisers: fac * mean|sigmoid(opacity)| and fac * mean|exp(scale)| summed over the C views
    const float sg = 1.0f / (1.0f + expf(-opac));
    vo += reg_opac * sg * (1.0f - sg);
    sig_sum = sg;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float ex = expf(scale[k]);
      vs[k] += reg_scale * ex;
      exp_sum += ex;
    }
    v_means[3 * g] = vm[0]; v_means[3 * g + 1] = vm[1]; v_means[3 * g + 2] = vm[2];
    reinterpret_cast<float4*>(v_quats)[g] = make_float4(vq[0], vq[1], vq[2], vq[3]);
    v_scales[3 * g] = vs[0]; v_scales[3 * g + 1] = vs[1]; v_scales[3 * g + 2] = vs[2];
    v_opacities[g] = vo;
    float4* o4 = reinterpret_cast<float4*>(v_sh + (size_t)g * 12);
    o4[0] = make_float4(vsh[0], vsh[1], vsh[2], vsh[3]);
    o4[1] = make_float4(vsh[4], vsh[5], vsh[6], vsh[7]);
    o4[2] = make_float4(vsh[8], vsh[9], vsh[10], vsh[11]);
  }
  if (reg_sums) {
    for (int off = 16; off; off >>= 1) {
      sig_sum += __shfl_xor_sync(0xffffffffu, sig_sum, off);
      exp_sum += __shfl_xor_sync(0xffffffffu, exp_sum, off);
    }
    if (lane_id() == 0) {
      atomicAdd(reg_sums, sig_sum);
      atomicAdd(reg_sums + 1, exp_sum);
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
  gs_project_bwd_kernel<<<(N + 127) / 128, 128, 0, stream>>>(
      means, quats, scales, opacities, shN, sh_coeffs * 3, reinterpret_cast<const GsCam*>(cams), N, C, (float)width,
      (float)height, eps2d, near_plane, far_plane, radius_clip, radii, reinterpret_cast<const float4*>(v_geomA),
      reinterpret_cast<const float4*>(v_geomB), reinterpret_cast<const float4*>(v_rgb), reg_opac, reg_scale, v_means,
      v_quats, v_scales, v_opacities, v_sh, reg_sums);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
