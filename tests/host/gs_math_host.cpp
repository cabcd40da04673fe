// TEST-ONLY host build of the per-Gaussian math header (starst3r_b200/csrc/gs_math.cuh): lets the CPU
// test-suite check the analytic projection / SH backward against autograd through the oracle without a
// GPU.  Never linked into the product library.
#include <stdint.h>
#include "../../starst3r_b200/csrc/gs_math.cuh"

extern "C" {

void host_project(const float* means, const float* quats, const float* scales, const float* shN, int sh_stride,
                  const float* cams, int N, int C, float W, float H, int32_t* radii, float* geom /*[C*N*6]*/,
                  float* rgb /*[C*N*3]*/) {
  for (int c = 0; c < C; ++c) {
    GsCam cam = reinterpret_cast<const GsCam*>(cams)[c];
    for (int g = 0; g < N; ++g) {
      GsProj o; GsProjTmp t;
      bool vis = gs_project(means + 3 * g, quats + 4 * g, scales + 3 * g, cam, W, H, 0.3f, 0.01f, 1e10f, 0.0f, o, t);
      size_t e = (size_t)c * N + g;
      radii[e] = vis ? o.radius : 0;
      float* q = geom + e * 6;
      q[0] = o.m2x; q[1] = o.m2y; q[2] = o.depth; q[3] = o.ca; q[4] = o.cb; q[5] = o.cc;
      float raw[3], dirn[3], il;
      gs_sh_color(means + 3 * g, cam.pos, shN + (size_t)g * sh_stride, rgb + e * 3, raw, dirn, &il);
    }
  }
}

void host_project_bwd(const float* means, const float* quats, const float* scales, const float* shN, int sh_stride,
                      const float* cams, int N, int C, float W, float H, const int32_t* radii, const float* v_m2,
                      const float* v_conic, const float* v_rgb, float* v_means, float* v_quats, float* v_scales,
                      float* v_sh /*[N*12]*/) {
  for (int g = 0; g < N; ++g) {
    float* vm = v_means + 3 * g; float* vq = v_quats + 4 * g; float* vs = v_scales + 3 * g; float* vsh = v_sh + 12 * g;
    for (int k = 0; k < 3; ++k) vm[k] = vs[k] = 0.f;
    for (int k = 0; k < 4; ++k) vq[k] = 0.f;
    for (int k = 0; k < 12; ++k) vsh[k] = 0.f;
    for (int c = 0; c < C; ++c) {
      size_t e = (size_t)c * N + g;
      if (radii[e] <= 0) continue;
      GsCam cam = reinterpret_cast<const GsCam*>(cams)[c];
      GsProj o; GsProjTmp t;
      if (!gs_project(means + 3 * g, quats + 4 * g, scales + 3 * g, cam, W, H, 0.3f, 0.01f, 1e10f, 0.0f, o, t)) continue;
      float col[3], raw[3], dirn[3], il;
      gs_sh_color(means + 3 * g, cam.pos, shN + (size_t)g * sh_stride, col, raw, dirn, &il);
      gs_sh_color_vjp(shN + (size_t)g * sh_stride, raw, dirn, il, v_rgb + e * 3, vsh, vm);
      gs_project_vjp(scales + 3 * g, cam, o, t, v_m2[e * 2], v_m2[e * 2 + 1], v_conic + e * 3, vm, vq, vs);
    }
  }
}
}
