// TEST-ONLY host build of the ALIGN math header (starst3r_b200/csrc/align_math.cuh): evaluates one iteration's
// loss and parameter gradients single-threaded with exactly the functions the CUDA kernels call, so the CPU
// test-suite can compare them with autograd through the oracle.  Never linked into the product library.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../starst3r_b200/csrc/align_math.cuh"
#include "../../include/starst3r_b200.h"

static void anchor(const St3rAlignProblem& pb, const AlignImgConst* ic, int a, int* img, float* u, float* v, float* core,
                   float* off) {
  *img = pb.anc_img[a]; *u = pb.anc_uv[2 * a]; *v = pb.anc_uv[2 * a + 1];
  *core = pb.core[ic[*img].core_off + pb.anc_k[a]]; *off = pb.anc_off[a];
}

extern "C" void host_align_eval(const St3rAlignProblem* prob, const float* pp, const float* log_focal, const float* quat,
                                const float* trans, const float* log_size, int mode, float gamma, float gamma_d,
                                float dust3r_w, float* out_loss, float* out_grad /*[N*11]*/, float* out_cam /*[N*20]*/) {
  const St3rAlignProblem pb = *prob;
  const int N = pb.n_img;
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  std::vector<AlignCamTmp> tmp(N);
  std::vector<AlignCam> cam(N);
  std::vector<AlignCamGrad> cg(N);
  std::vector<float> gcam((size_t)N * 17, 0.f);
  float smin = INFINITY;
  for (int i = 0; i < N; ++i) {
    al_cam_local_fwd(ic[i], log_focal[i], log_size[i], quat + 4 * i, tmp[i]);
    if (tmp[i].s < smin) smin = tmp[i].s;
  }
  int ties = 0;
  for (int i = 0; i < N; ++i) ties += tmp[i].s == smin;
  const float g = 1.0f / smin;
  al_chain_fwd(N, pb.root, pb.edges, tmp.data(), trans);
  for (int i = 0; i < N; ++i) al_cam_final_fwd(ic[i], pp + 2 * i, g, tmp[i], cam[i]);
  memcpy(out_cam, cam.data(), (size_t)N * sizeof(AlignCam));

  auto consts = [](float gm, float* off, float* offp) {
    if (gm == 1.f) { *off = 0; *offp = 0; return; }
    double o = pow(1.0 / gm, 1.0 / (gm - 1.0)); *off = (float)o; *offp = (float)pow(o, (double)gm);
  };
  float off_m, offp_m, off_d, offp_d;
  consts(gamma, &off_m, &offp_m); consts(gamma_d, &off_d, &offp_d);
  double loss = 0.0;
  if (mode == 0 && pb.n3 > 0) {
    float scale = 1.0f / pb.norm3;
    for (int m = 0; m < pb.n3; ++m) {
      int i1, i2; float u1, v1, c1, o1, u2, v2, c2, o2;
      anchor(pb, ic, pb.e3_a1[m], &i1, &u1, &v1, &c1, &o1);
      anchor(pb, ic, pb.e3_a2[m], &i2, &u2, &v2, &c2, &o2);
      float P1[3], P2[3], pc1[3], pc2[3], z1, z2, D1, D2, op1, op2;
      al_anchor_point(cam[i1], u1, v1, c1, o1, P1, pc1, &z1, &D1, &op1);
      al_anchor_point(cam[i2], u2, v2, c2, o2, P2, pc2, &z2, &D2, &op2);
      float d[3] = {P1[0] - P2[0], P1[1] - P2[1], P1[2] - P2[2]};
      float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), dl;
      float l = al_gamma_loss(dist, gamma, off_m, offp_m, &dl);
      float cw = pb.e3_conf[m] * scale;
      loss += cw * l;
      float k = dist > 0 ? cw * dl / dist : 0.f;
      float G1[3] = {k * d[0], k * d[1], k * d[2]}, G2[3] = {-k * d[0], -k * d[1], -k * d[2]};
      al_anchor_point_vjp(cam[i1], u1, v1, c1, o1, pc1, z1, D1, op1, G1, gcam.data() + 17 * i1);
      al_anchor_point_vjp(cam[i2], u2, v2, c2, o2, pc2, z2, D2, op2, G2, gcam.data() + 17 * i2);
    }
  }
  if (mode == 1 && pb.n2 > 0) {
    float scale = 1.0f / pb.norm2;
    for (int m = 0; m < pb.n2; ++m) {
      int i1 = pb.e2_img1[m], i2; float u2, v2, c2, o2;
      anchor(pb, ic, pb.e2_a2[m], &i2, &u2, &v2, &c2, &o2);
      float P2[3], pc2[3], z2, D2, op2;
      al_anchor_point(cam[i2], u2, v2, c2, o2, P2, pc2, &z2, &D2, &op2);
      float uv[2]; AlignReproj q;
      al_reproj(cam[i1], P2, uv, q);
      float d[2] = {pb.e2_pix[2 * m] - uv[0], pb.e2_pix[2 * m + 1] - uv[1]};
      float dist = sqrtf(d[0] * d[0] + d[1] * d[1]), dl;
      float l = al_gamma_loss(dist, gamma, off_m, offp_m, &dl);
      float cw = pb.e2_conf[m] * scale;
      loss += cw * l;
      float k = dist > 0 ? cw * dl / dist : 0.f;
      float Guv[2] = {-k * d[0], -k * d[1]}, GP[3];
      al_reproj_vjp(cam[i1], P2, q, Guv, gcam.data() + 17 * i1, GP);
      al_anchor_point_vjp(cam[i2], u2, v2, c2, o2, pc2, z2, D2, op2, GP, gcam.data() + 17 * i2);
    }
  }
  if (pb.nd > 0 && dust3r_w != 0.f) {
    float scale = dust3r_w / pb.normd;
    for (int m = 0; m < pb.nd; ++m) {
      int i1, i2 = pb.ed_img2[m]; float u1, v1, c1, o1;
      anchor(pb, ic, pb.ed_a1[m], &i1, &u1, &v1, &c1, &o1);
      float P1[3], pc1[3], z1, D1, op1;
      al_anchor_point(cam[i1], u1, v1, c1, o1, P1, pc1, &z1, &D1, &op1);
      const float* tg = pb.ed_tgt + 3 * m;
      float T[3];
      al_mat3_vec(cam[i2].R, tg, T);
      T[0] += cam[i2].t[0]; T[1] += cam[i2].t[1]; T[2] += cam[i2].t[2];
      float d[3] = {P1[0] - T[0], P1[1] - T[1], P1[2] - T[2]};
      float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), dl;
      float l = al_gamma_loss(dist, gamma_d, off_d, offp_d, &dl);
      float cw = pb.ed_conf[m] * scale;
      loss += cw * l;
      float k = dist > 0 ? cw * dl / dist : 0.f;
      float G1[3] = {k * d[0], k * d[1], k * d[2]};
      al_anchor_point_vjp(cam[i1], u1, v1, c1, o1, pc1, z1, D1, op1, G1, gcam.data() + 17 * i1);
      float* g2 = gcam.data() + 17 * i2;
      for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) g2[3 * a + b] += -G1[a] * tg[b]; g2[9 + a] += -G1[a]; }
    }
  }
  *out_loss = (float)loss;
  float gg = 0.f;
  for (int i = 0; i < N; ++i) { al_cam_final_bwd(ic[i], pp + 2 * i, g, tmp[i], gcam.data() + 17 * i, cg[i]); gg += cg[i].g_g; }
  al_chain_bwd(N, pb.root, pb.edges, tmp.data(), trans, cg.data());
  for (int i = 0; i < N; ++i) {
    float* o = out_grad + 11 * i;
    al_cam_local_bwd(tmp[i], cg[i], tmp[i].s == smin ? -gg * g * g / ties : 0.f, o, o + 2, o + 3, o + 7, o + 10);
  }
}
