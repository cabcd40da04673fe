// Per-image camera math of the sparse global alignment optimiser, forward and analytic backward.
//
// Restates make_K_cam_depth of starster/reconstruct.py:209-261:
//   f = clip(exp(log_focal)), K = [[f,0,pp.x W],[0,f,pp.y H],[0,0,1]], s = exp(log_size),
//   g = 1 / min_i s_i, z = s * median * f / base_focal, Rel = [R_xyzw(normalize(quat)) | trans],
//   kinematic chain along the MST, t' = g (T.t - T.R off), off = z (W/f (0.5 - pp.x), H/f (0.5 - pp.y), 1),
//   depth[k] = g (z + (core[k] - 1) median s)  =  A + B core[k].
// and make_pts3d / proj3d / reproj2d of mast3r/cloud_opt/sparse_ga.py:469-501,977-981.
// The optimiser kernels (align.cu) split every iteration in: per-image "local" stage (parallel over
// images), the MST chain (sequential in tree order, tiny), a per-correspondence stage (parallel over
// matches, accumulates d loss / d (R, t', f, cx, cy, A, B) per image), and the mirror-image backward.
// Host+device so tests/host/align_math_host.cpp can check the backward against autograd through the
// unmodified reference on CPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define AL_HD __host__ __device__ __forceinline__
#else
#define AL_HD inline
#endif

struct AlignImgConst {
  float W, H, base_focal, median, min_focal, max_focal;
  int core_off, n_core;   // slice of the concatenated (median-normalised) core depth array
};

// What the per-correspondence stage needs of one image.  cam2w = [R | t].
struct AlignCam {
  float R[9];
  float t[3];
  float f, cx, cy, A, B, bf;
  float inv_f;     // 1 / f: the per-correspondence stage multiplies instead of dividing (14 IEEE divisions per entry before)
  float pad;
};
constexpr int ALIGN_CAM_GRADS = 17;  // R(9) t(3) f cx cy A B

// Saved by the forward for the backward.
struct AlignCamTmp {
  float relR[9];      // R(normalize(quat))
  float qn[4];        // normalised quaternion (x, y, z, w)
  float inv_norm;
  float s, z, f;
  float TR[9], Tt[3]; // chained pose before the re-parameterisation
  float off[3];
  int f_clipped;
};

// roma.unitquat_to_rotmat, XYZW convention (identity = (0,0,0,1)).
AL_HD void al_quat_xyzw_to_rotmat(const float* q, float* R, float* qn, float* inv_norm) {
  float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  float inv = 1.0f / fmaxf(n, 1e-12f);  // F.normalize eps
  float x = q[0] * inv, y = q[1] * inv, z = q[2] * inv, w = q[3] * inv;
  qn[0] = x; qn[1] = y; qn[2] = z; qn[3] = w;
  *inv_norm = inv;
  R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - z * w);       R[2] = 2.f * (x * z + y * w);
  R[3] = 2.f * (x * y + z * w);       R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - x * w);
  R[6] = 2.f * (x * z - y * w);       R[7] = 2.f * (y * z + x * w);       R[8] = 1.f - 2.f * (x * x + y * y);
}

// G = dL/dR (row-major) -> dL/dq for the un-normalised XYZW quaternion.
AL_HD void al_quat_xyzw_vjp(const float* qn, float inv_norm, const float* G, float* vq) {
  float x = qn[0], y = qn[1], z = qn[2], w = qn[3];
  float vx = 2.f * (-2.f * x * (G[4] + G[8]) + y * (G[1] + G[3]) + z * (G[2] + G[6]) + w * (G[7] - G[5]));
  float vy = 2.f * (x * (G[1] + G[3]) - 2.f * y * (G[0] + G[8]) + z * (G[5] + G[7]) + w * (G[2] - G[6]));
  float vz = 2.f * (x * (G[2] + G[6]) + y * (G[5] + G[7]) - 2.f * z * (G[0] + G[4]) + w * (G[3] - G[1]));
  float vw = 2.f * (x * (G[7] - G[5]) + y * (G[2] - G[6]) + z * (G[3] - G[1]));
  float d = vx * x + vy * y + vz * z + vw * w;
  vq[0] = (vx - d * x) * inv_norm;
  vq[1] = (vy - d * y) * inv_norm;
  vq[2] = (vz - d * z) * inv_norm;
  vq[3] = (vw - d * w) * inv_norm;
}

AL_HD void al_mat3_mul(const float* A, const float* B, float* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
AL_HD void al_mat3_vec(const float* A, const float* v, float* o) {
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
AL_HD void al_mat3T_vec(const float* A, const float* v, float* o) {
  for (int j = 0; j < 3; ++j) o[j] = A[j] * v[0] + A[3 + j] * v[1] + A[6 + j] * v[2];
}

// ---- forward, stage 1 (per image): focal, size, depth-plane distance, relative pose -------------
AL_HD void al_cam_local_fwd(const AlignImgConst& ic, float log_focal, float log_size, const float* quat,
                            AlignCamTmp& t) {
  float fe = expf(log_focal);
  t.f_clipped = (fe < ic.min_focal || fe > ic.max_focal) ? 1 : 0;
  t.f = fminf(fmaxf(fe, ic.min_focal), ic.max_focal);
  t.s = expf(log_size);
  t.z = t.s * ic.median * t.f / ic.base_focal;
  al_quat_xyzw_to_rotmat(quat, t.relR, t.qn, &t.inv_norm);
}

// ---- forward, stage 2 (sequential over the tree): T_child = T_parent * Rel_child ------------------
AL_HD void al_chain_fwd(int N, int root, const int* edges, AlignCamTmp* tmp, const float* trans) {
  for (int k = 0; k < 9; ++k) tmp[root].TR[k] = tmp[root].relR[k];
  for (int k = 0; k < 3; ++k) tmp[root].Tt[k] = trans[3 * root + k];
  for (int e = 0; e < N - 1; ++e) {
    int i = edges[2 * e], j = edges[2 * e + 1];
    al_mat3_mul(tmp[i].TR, tmp[j].relR, tmp[j].TR);
    float v[3];
    al_mat3_vec(tmp[i].TR, trans + 3 * j, v);
    for (int k = 0; k < 3; ++k) tmp[j].Tt[k] = v[k] + tmp[i].Tt[k];
  }
}

// ---- forward, stage 3 (per image): re-parameterised translation, intrinsics, depth affine map ------
AL_HD void al_cam_final_fwd(const AlignImgConst& ic, const float* pp, float g, AlignCamTmp& t, AlignCam& c) {
  t.off[0] = t.z * (ic.W / t.f * (0.5f - pp[0]));
  t.off[1] = t.z * (ic.H / t.f * (0.5f - pp[1]));
  t.off[2] = t.z;
  float Ro[3];
  al_mat3_vec(t.TR, t.off, Ro);
  for (int k = 0; k < 9; ++k) c.R[k] = t.TR[k];
  for (int k = 0; k < 3; ++k) c.t[k] = g * (t.Tt[k] - Ro[k]);
  c.f = t.f;
  c.inv_f = 1.0f / t.f;
  c.pad = 0.f;
  c.cx = pp[0] * ic.W;
  c.cy = pp[1] * ic.H;
  c.A = g * (t.z - ic.median * t.s);
  c.B = g * (ic.median * t.s);
  c.bf = ic.base_focal;
}

// ---- backward, stage 3: gcam (dL/d R,t,f,cx,cy,A,B) -> chain gradients + partial scalar gradients ----
struct AlignCamGrad {
  float GTR[9], GTt[3];       // dL/d(T.R), dL/d(T.t)
  float g_g;                  // dL/dg contribution of this image
  float g_f, g_s, g_pp[2];    // direct contributions (completed by the local backward)
};

AL_HD void al_cam_final_bwd(const AlignImgConst& ic, const float* pp, float g, const AlignCamTmp& t,
                            const float* gc /*[17]*/, AlignCamGrad& o) {
  const float* GR = gc;
  const float* Gt = gc + 9;
  float Gf = gc[12], Gcx = gc[13], Gcy = gc[14], GA = gc[15], GB = gc[16];
  float Ro[3];
  al_mat3_vec(t.TR, t.off, Ro);
  // t' = g (Tt - TR off)
  o.g_g = Gt[0] * (t.Tt[0] - Ro[0]) + Gt[1] * (t.Tt[1] - Ro[1]) + Gt[2] * (t.Tt[2] - Ro[2]);
  for (int k = 0; k < 3; ++k) o.GTt[k] = g * Gt[k];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) o.GTR[3 * a + b] = GR[3 * a + b] - g * Gt[a] * t.off[b];
  float Goff[3];
  al_mat3T_vec(t.TR, Gt, Goff);
  for (int k = 0; k < 3; ++k) Goff[k] *= -g;
  // off = z (W/f (0.5 - ppx), H/f (0.5 - ppy), 1)
  float ux = ic.W / t.f * (0.5f - pp[0]), uy = ic.H / t.f * (0.5f - pp[1]);
  float Gz = Goff[0] * ux + Goff[1] * uy + Goff[2];
  o.g_f = Gf + t.z * (-ux / t.f * Goff[0] - uy / t.f * Goff[1]);
  o.g_pp[0] = -t.z * ic.W / t.f * Goff[0] + ic.W * Gcx;
  o.g_pp[1] = -t.z * ic.H / t.f * Goff[1] + ic.H * Gcy;
  // A = g (z - med s), B = g med s
  o.g_g += GA * (t.z - ic.median * t.s) + GB * ic.median * t.s;
  Gz += g * GA;
  o.g_s = -g * ic.median * GA + g * ic.median * GB;
  // z = s med f / bf
  o.g_s += Gz * ic.median * t.f / ic.base_focal;
  o.g_f += Gz * t.s * ic.median / ic.base_focal;
}

// ---- backward, stage 2 (sequential, reverse tree order): chain -> dL/d(Rel.R), dL/d(trans) ---------
// On exit grads[j].GTR holds dL/d(Rel_j.R) and grads[j].GTt holds dL/d(trans_j).
AL_HD void al_chain_bwd(int N, int root, const int* edges, const AlignCamTmp* tmp, const float* trans,
                        AlignCamGrad* grads) {
  for (int e = N - 2; e >= 0; --e) {
    int i = edges[2 * e], j = edges[2 * e + 1];
    const float* GTRj = grads[j].GTR;
    const float* GTtj = grads[j].GTt;
    // parent accumulates: G_TR_i += G_TR_j Rel_j.R^T + G_Tt_j (x) trans_j ; G_Tt_i += G_Tt_j
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        float s = GTtj[a] * trans[3 * j + b];
        for (int k = 0; k < 3; ++k) s += GTRj[3 * a + k] * tmp[j].relR[3 * b + k];
        grads[i].GTR[3 * a + b] += s;
      }
    for (int k = 0; k < 3; ++k) grads[i].GTt[k] += GTtj[k];
    // child: dL/dRel_j.R = T_i.R^T G_TR_j ; dL/dtrans_j = T_i.R^T G_Tt_j
    float GR[9], Gt[3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        GR[3 * a + b] = tmp[i].TR[a] * GTRj[b] + tmp[i].TR[3 + a] * GTRj[3 + b] + tmp[i].TR[6 + a] * GTRj[6 + b];
    al_mat3T_vec(tmp[i].TR, GTtj, Gt);
    for (int k = 0; k < 9; ++k) grads[j].GTR[k] = GR[k];
    for (int k = 0; k < 3; ++k) grads[j].GTt[k] = Gt[k];
  }
  (void)root;  // the root's GTR / GTt already are dL/dRel_root
}

// ---- backward, stage 1 (per image): -> gradients of the raw parameters -----------------------------
// g_s_extra = contribution through g = 1 / min s (non-zero only for the arg-min image).
AL_HD void al_cam_local_bwd(const AlignCamTmp& t, const AlignCamGrad& gr, float g_s_extra, float* g_pp,
                            float* g_log_focal, float* g_quat, float* g_trans, float* g_log_size) {
  g_pp[0] = gr.g_pp[0];
  g_pp[1] = gr.g_pp[1];
  *g_log_focal = t.f_clipped ? 0.f : gr.g_f * t.f;
  *g_log_size = (gr.g_s + g_s_extra) * t.s;
  al_quat_xyzw_vjp(t.qn, t.inv_norm, gr.GTR, g_quat);
  for (int k = 0; k < 3; ++k) g_trans[k] = gr.GTt[k];
}

// ============================ per-correspondence stage =============================================
AL_HD float al_gamma_loss(float d, float gamma, float offset, float off_pow, float* dloss_dd) {
  // (d + o)^gamma - o^gamma  (cloud_opt/utils/losses.py:19-28); gamma == 1 -> d
  if (gamma == 1.0f) { *dloss_dd = 1.0f; return d; }
  float b = d + offset;
#ifdef __CUDA_ARCH__
  float p = __powf(b, gamma - 1.0f);        // ex2.approx(lg2.approx(b) (gamma - 1)): b >= offset > 0, relative error ~1e-6
#else
  float p = powf(b, gamma - 1.0f);
#endif
  *dloss_dd = gamma * p;
  return p * b - off_pow;
}

// World point of one anchor (sparse_ga.py:469-501): returns p_cam in pc and the world point in P.
AL_HD void al_anchor_point(const AlignCam& c, float u, float v, float core, float off, float* P, float* pc,
                           float* z_out, float* D_out, float* op_out) {
  float D = c.A + c.B * core;
  float op = 1.0f + (off - 1.0f) * (c.bf * c.inv_f);
  float z = D * op;
  pc[0] = z * ((u - c.cx) * c.inv_f);
  pc[1] = z * ((v - c.cy) * c.inv_f);
  pc[2] = z;
  al_mat3_vec(c.R, pc, P);
  P[0] += c.t[0]; P[1] += c.t[1]; P[2] += c.t[2];
  *z_out = z; *D_out = D; *op_out = op;
}

// dL/dP -> gradient of the owning image's 17 local camera quantities (accumulated into g[17]).
AL_HD void al_anchor_point_vjp(const AlignCam& c, float u, float v, float core, float off, const float* pc, float z,
                               float D, float op, const float* GP, float* g) {
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) g[3 * a + b] += GP[a] * pc[b];
  g[9] += GP[0]; g[10] += GP[1]; g[11] += GP[2];
  float Gpc[3];
  al_mat3T_vec(c.R, GP, Gpc);
  float rx = (u - c.cx) * c.inv_f, ry = (v - c.cy) * c.inv_f;
  float Gz = Gpc[0] * rx + Gpc[1] * ry + Gpc[2];
  float Grx = z * Gpc[0], Gry = z * Gpc[1];
  g[13] += -Grx * c.inv_f;
  g[14] += -Gry * c.inv_f;
  float Gf = -(Grx * rx + Gry * ry) * c.inv_f;
  float GD = Gz * op, Gop = Gz * D;
  Gf += -Gop * (off - 1.0f) * c.bf * (c.inv_f * c.inv_f);
  g[12] += Gf;
  g[15] += GD;
  g[16] += GD * core;
}

// reproj2d(K w2cam, P) (sparse_ga.py:977-981) for the image owning the pixel: returns uv, fills intermediates.
struct AlignReproj { float r[3]; float zc, inv_zc; float uh, vh; int clip_u, clip_v; };
AL_HD void al_reproj(const AlignCam& c, const float* P, float* uv, AlignReproj& q) {
  float d[3] = {P[0] - c.t[0], P[1] - c.t[1], P[2] - c.t[2]};
  al_mat3T_vec(c.R, d, q.r);                 // camera-frame point  R^T (P - t)
  q.zc = fmaxf(q.r[2], 1e-3f);
  q.uh = c.f * q.r[0] + c.cx * q.r[2];
  q.vh = c.f * q.r[1] + c.cy * q.r[2];
  q.inv_zc = 1.0f / q.zc;
  float u = q.uh * q.inv_zc, v = q.vh * q.inv_zc;
  q.clip_u = (u < -1000.f || u > 2000.f);
  q.clip_v = (v < -1000.f || v > 2000.f);
  uv[0] = fminf(fmaxf(u, -1000.f), 2000.f);
  uv[1] = fminf(fmaxf(v, -1000.f), 2000.f);
}
// dL/duv -> g1[17] (pixel-owning image: R, t, f, cx, cy) and GP (world point).
AL_HD void al_reproj_vjp(const AlignCam& c, const float* P, const AlignReproj& q, const float* Guv, float* g1,
                         float* GP) {
  float Gu = q.clip_u ? 0.f : Guv[0], Gv = q.clip_v ? 0.f : Guv[1];
  float Guh = Gu * q.inv_zc, Gvh = Gv * q.inv_zc;
  float Gzc = -(Gu * q.uh + Gv * q.vh) * (q.inv_zc * q.inv_zc);
  float Gr[3];
  Gr[0] = Guh * c.f;
  Gr[1] = Gvh * c.f;
  Gr[2] = Guh * c.cx + Gvh * c.cy + (q.r[2] > 1e-3f ? Gzc : 0.f);
  g1[12] += Guh * q.r[0] + Gvh * q.r[1];
  g1[13] += Guh * q.r[2];
  g1[14] += Gvh * q.r[2];
  // r = R^T (P - t):  dL/dP = R Gr, dL/dt = -R Gr, dL/dR[a][b] = (P - t)[a] Gr[b]
  float RG[3];
  al_mat3_vec(c.R, Gr, RG);
  for (int k = 0; k < 3; ++k) { GP[k] = RG[k]; g1[9 + k] += -RG[k]; }
  float d[3] = {P[0] - c.t[0], P[1] - c.t[1], P[2] - c.t[2]};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) g1[3 * a + b] += d[a] * Gr[b];
}
