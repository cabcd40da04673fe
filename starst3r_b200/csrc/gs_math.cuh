// Per-Gaussian math of the 3DGS rasteriser (projection, SH colour) and its analytic backward.
//
// Follows the published gsplat 1.4 algorithm: quat/scale ->
// Sigma, world -> camera, perspective Jacobian with principal-point-aware tangent clamp,
// eps2d blur, conic, radius = ceil(3 sqrt(lambda_max)), screen culling; degree-1 SH.
// The forward is written as explicit single operations in a fixed order; the translation
// unit that instantiates it for the product (gs_project.cu) is compiled with -fmad=false
// so that radii / tile ranges are bit-identical to the CPU restatement (oracle/gs_oracle.py).
// Functions are host+device so the test-only harness (tests/host/gs_math_host.cpp) can
// check the backward against autograd without a GPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD inline
#endif

#define GS_SH_C0 0.2820947917738781f
#define GS_SH_C1 0.48860251190292f

struct GsCam {
  float R[9];  // world->camera rotation, row-major
  float t[3];
  float fx, fy, cx, cy;
  float pos[3];  // camera centre in world space (inverse(viewmat)[:3,3])
};

struct GsProj {
  float m2x, m2y, depth, ca, cb, cc;
  int radius;  // 0 => culled
};

GS_HD float gs_dot3(float a0, float a1, float a2, float b0, float b1, float b2) { return (a0 * b0 + a1 * b1) + a2 * b2; }

// wxyz quaternion (normalised here) -> rotation (row-major)
GS_HD void gs_quat_to_rotmat(const float* q, float* R, float* qn, float* inv_norm) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  float inv = 1.0f / sqrtf(((w * w + x * x) + y * y) + z * z);
  w *= inv; x *= inv; y *= inv; z *= inv;
  if (qn) { qn[0] = w; qn[1] = x; qn[2] = y; qn[3] = z; }
  if (inv_norm) *inv_norm = inv;
  float x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
  R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
  R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
}

// Intermediates shared by forward and backward.
struct GsProjTmp {
  float R[9], qn[4], inv_norm;
  float M[9];      // R * diag(s)
  float S[9];      // Sigma (symmetric, full)
  float Sc[9];     // Sigma in camera frame
  float x, y, z, rz, rz2, tx, ty;
  float j00, j02, j11, j12;
  float c00, c01, c11, det;  // blurred 2D covariance
  bool clamp_x, clamp_y;
};

GS_HD bool gs_project(const float* mean, const float* quat, const float* scale, const GsCam& cam, float W, float H,
                      float eps2d, float near_plane, float far_plane, float radius_clip, GsProj& o, GsProjTmp& t) {
  o.radius = 0;
  const float* Rc = cam.R;
  t.x = gs_dot3(Rc[0], Rc[1], Rc[2], mean[0], mean[1], mean[2]) + cam.t[0];
  t.y = gs_dot3(Rc[3], Rc[4], Rc[5], mean[0], mean[1], mean[2]) + cam.t[1];
  t.z = gs_dot3(Rc[6], Rc[7], Rc[8], mean[0], mean[1], mean[2]) + cam.t[2];
  o.depth = t.z;
  if (!(t.z >= near_plane) || !(t.z <= far_plane)) return false;

  gs_quat_to_rotmat(quat, t.R, t.qn, &t.inv_norm);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t.M[3 * i + j] = t.R[3 * i + j] * scale[j];
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      float e = gs_dot3(t.M[3 * i], t.M[3 * i + 1], t.M[3 * i + 2], t.M[3 * j], t.M[3 * j + 1], t.M[3 * j + 2]);
      t.S[3 * i + j] = e;
      t.S[3 * j + i] = e;
    }
  float T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      T[3 * i + j] = gs_dot3(Rc[3 * i], Rc[3 * i + 1], Rc[3 * i + 2], t.S[j], t.S[3 + j], t.S[6 + j]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      t.Sc[3 * i + j] = gs_dot3(T[3 * i], T[3 * i + 1], T[3 * i + 2], Rc[3 * j], Rc[3 * j + 1], Rc[3 * j + 2]);

  const float fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
  float tan_fovx = (0.5f * W) / fx, tan_fovy = (0.5f * H) / fy;
  float lim_x_pos = (W - cx) / fx + 0.3f * tan_fovx, lim_x_neg = cx / fx + 0.3f * tan_fovx;
  float lim_y_pos = (H - cy) / fy + 0.3f * tan_fovy, lim_y_neg = cy / fy + 0.3f * tan_fovy;
  t.rz = 1.0f / t.z;
  t.rz2 = t.rz * t.rz;
  float xr = t.x * t.rz, yr = t.y * t.rz;
  t.clamp_x = !(xr <= lim_x_pos && xr >= -lim_x_neg);
  t.clamp_y = !(yr <= lim_y_pos && yr >= -lim_y_neg);
  t.tx = t.z * fminf(lim_x_pos, fmaxf(-lim_x_neg, xr));
  t.ty = t.z * fminf(lim_y_pos, fmaxf(-lim_y_neg, yr));
  t.j00 = fx * t.rz;
  t.j02 = -(fx * t.tx) * t.rz2;
  t.j11 = fy * t.rz;
  t.j12 = -(fy * t.ty) * t.rz2;
  const float* Sc = t.Sc;
  float a0 = t.j00 * Sc[0] + t.j02 * Sc[6];
  float a2 = t.j00 * Sc[2] + t.j02 * Sc[8];
  float a1 = t.j00 * Sc[1] + t.j02 * Sc[7];
  float b1 = t.j11 * Sc[4] + t.j12 * Sc[7];
  float b2 = t.j11 * Sc[5] + t.j12 * Sc[8];
  float c00 = a0 * t.j00 + a2 * t.j02;
  float c01 = a1 * t.j11 + a2 * t.j12;
  float c11 = b1 * t.j11 + b2 * t.j12;
  o.m2x = (fx * t.x) * t.rz + cx;
  o.m2y = (fy * t.y) * t.rz + cy;

  c00 = c00 + eps2d;
  c11 = c11 + eps2d;
  float det = c00 * c11 - c01 * c01;
  t.c00 = c00; t.c01 = c01; t.c11 = c11; t.det = det;
  if (!(det > 0.0f)) return false;
  float inv_det = 1.0f / det;
  o.ca = c11 * inv_det;
  o.cb = -c01 * inv_det;
  o.cc = c00 * inv_det;
  float b = 0.5f * (c00 + c11);
  float v1 = b + sqrtf(fmaxf(b * b - det, 0.01f));
  float radius = ceilf(3.0f * sqrtf(v1));
  if (!(radius > radius_clip)) return false;
  if (o.m2x + radius <= 0.0f || o.m2x - radius >= W || o.m2y + radius <= 0.0f || o.m2y - radius >= H) return false;
  o.radius = (int)radius;
  return true;
}

// Degree-1 SH colour of one Gaussian seen from cam.pos; sh = first 4 coefficients x 3 channels ([4][3]).
// Returns the pre-clamp value in `raw` (needed by the backward) and the clamped colour in rgb.
GS_HD void gs_sh_color(const float* mean, const float* campos, const float* sh, float* rgb, float* raw, float* dirn,
                       float* inv_len) {
  float dx = mean[0] - campos[0], dy = mean[1] - campos[1], dz = mean[2] - campos[2];
  float il = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
  dx *= il; dy *= il; dz *= il;
  if (dirn) { dirn[0] = dx; dirn[1] = dy; dirn[2] = dz; }
  if (inv_len) *inv_len = il;
  for (int c = 0; c < 3; ++c) {
    float r = GS_SH_C0 * sh[c] + GS_SH_C1 * (-dy * sh[3 + c] + dz * sh[6 + c] - dx * sh[9 + c]);
    if (raw) raw[c] = r;
    rgb[c] = fmaxf(r + 0.5f, 0.0f);
  }
}

// Backward of gs_sh_color.  v_rgb -> v_sh[12] (+=), v_mean[3] (+=).
GS_HD void gs_sh_color_vjp(const float* sh, const float* raw, const float* dirn, float inv_len, const float* v_rgb,
                           float* v_sh, float* v_mean) {
  float vdx = 0.f, vdy = 0.f, vdz = 0.f;
  for (int c = 0; c < 3; ++c) {
    float v = (raw[c] + 0.5f >= 0.0f) ? v_rgb[c] : 0.0f;
    v_sh[c] += GS_SH_C0 * v;
    v_sh[3 + c] += -GS_SH_C1 * dirn[1] * v;
    v_sh[6 + c] += GS_SH_C1 * dirn[2] * v;
    v_sh[9 + c] += -GS_SH_C1 * dirn[0] * v;
    vdx += -GS_SH_C1 * sh[9 + c] * v;
    vdy += -GS_SH_C1 * sh[3 + c] * v;
    vdz += GS_SH_C1 * sh[6 + c] * v;
  }
  float dotp = vdx * dirn[0] + vdy * dirn[1] + vdz * dirn[2];
  v_mean[0] += (vdx - dotp * dirn[0]) * inv_len;
  v_mean[1] += (vdy - dotp * dirn[1]) * inv_len;
  v_mean[2] += (vdz - dotp * dirn[2]) * inv_len;
}

// Backward of gs_project for one (camera, Gaussian): (v_m2x, v_m2y, v_conic[3]) ->
// v_mean[3], v_quat[4], v_scale[3] (all +=).  `t` is the forward's intermediate record.
GS_HD void gs_project_vjp(const float* scale, const GsCam& cam, const GsProj& o, const GsProjTmp& t, float v_m2x,
                          float v_m2y, const float* v_conic, float* v_mean, float* v_quat, float* v_scale) {
  // conic = inverse(cov2d):  v_cov = -X * V * X  with X = conic, V = sym(v_conic) (off-diagonal halved)
  float Xa = o.ca, Xb = o.cb, Xc = o.cc;
  float Va = v_conic[0], Vb = 0.5f * v_conic[1], Vc = v_conic[2];
  // P = X * V
  float p00 = Xa * Va + Xb * Vb, p01 = Xa * Vb + Xb * Vc;
  float p10 = Xb * Va + Xc * Vb, p11 = Xb * Vb + Xc * Vc;
  // G = -(P * X)  (symmetric 2x2 gradient w.r.t. the full cov2d matrix)
  float g00 = -(p00 * Xa + p01 * Xb);
  float g01 = -(p00 * Xb + p01 * Xc);
  float g10 = -(p10 * Xa + p11 * Xb);
  float g11 = -(p10 * Xb + p11 * Xc);

  const float* Sc = t.Sc;
  // J = [[j00, 0, j02], [0, j11, j12]]
  // v_Sc = J^T G J (3x3)
  float J[6] = {t.j00, 0.f, t.j02, 0.f, t.j11, t.j12};
  float GJ[6];  // G * J (2x3)
  for (int k = 0; k < 3; ++k) {
    GJ[k] = g00 * J[k] + g01 * J[3 + k];
    GJ[3 + k] = g10 * J[k] + g11 * J[3 + k];
  }
  float vSc[9];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) vSc[3 * i + k] = J[i] * GJ[k] + J[3 + i] * GJ[3 + k];
  // v_J = G J Sc^T + G^T J Sc   (2x3)
  float GtJ[6];
  for (int k = 0; k < 3; ++k) {
    GtJ[k] = g00 * J[k] + g10 * J[3 + k];
    GtJ[3 + k] = g01 * J[k] + g11 * J[3 + k];
  }
  float vJ[6];
  for (int r = 0; r < 2; ++r)
    for (int k = 0; k < 3; ++k) {
      float s = 0.f;
      for (int m = 0; m < 3; ++m) s += GJ[3 * r + m] * Sc[3 * k + m] + GtJ[3 * r + m] * Sc[3 * m + k];
      vJ[3 * r + k] = s;
    }

  const float fx = cam.fx, fy = cam.fy;
  float rz = t.rz, rz2 = t.rz2, rz3 = rz2 * rz;
  float vx = fx * rz * v_m2x;
  float vy = fy * rz * v_m2y;
  float vz = -(fx * t.x * v_m2x + fy * t.y * v_m2y) * rz2;
  if (!t.clamp_x) vx += -fx * rz2 * vJ[2]; else vz += -fx * rz3 * vJ[2] * t.tx;
  if (!t.clamp_y) vy += -fy * rz2 * vJ[5]; else vz += -fy * rz3 * vJ[5] * t.ty;
  vz += -fx * rz2 * vJ[0] - fy * rz2 * vJ[4] + 2.f * fx * t.tx * rz3 * vJ[2] + 2.f * fy * t.ty * rz3 * vJ[5];

  const float* Rc = cam.R;
  // v_mean += Rc^T v_meanc
  v_mean[0] += Rc[0] * vx + Rc[3] * vy + Rc[6] * vz;
  v_mean[1] += Rc[1] * vx + Rc[4] * vy + Rc[7] * vz;
  v_mean[2] += Rc[2] * vx + Rc[5] * vy + Rc[8] * vz;
  // v_S = Rc^T vSc Rc
  float A[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[3 * i + j] = Rc[i] * vSc[j] + Rc[3 + i] * vSc[3 + j] + Rc[6 + i] * vSc[6 + j];
  float vS[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) vS[3 * i + j] = A[3 * i] * Rc[j] + A[3 * i + 1] * Rc[3 + j] + A[3 * i + 2] * Rc[6 + j];
  // Sigma = M M^T:  v_M = (v_S + v_S^T) M
  float vM[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s += (vS[3 * i + k] + vS[3 * k + i]) * t.M[3 * k + j];
      vM[3 * i + j] = s;
    }
  // M = R diag(s)
  float G[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G[3 * i + j] = vM[3 * i + j] * scale[j];
  for (int j = 0; j < 3; ++j) v_scale[j] += t.R[j] * vM[j] + t.R[3 + j] * vM[3 + j] + t.R[6 + j] * vM[6 + j];
  float w = t.qn[0], x = t.qn[1], y = t.qn[2], z = t.qn[3];
  float vq[4];
  vq[0] = 2.f * (x * (G[7] - G[5]) + y * (G[2] - G[6]) + z * (G[3] - G[1]));
  vq[1] = 2.f * (-2.f * x * (G[4] + G[8]) + y * (G[1] + G[3]) + z * (G[2] + G[6]) + w * (G[7] - G[5]));
  vq[2] = 2.f * (x * (G[1] + G[3]) - 2.f * y * (G[0] + G[8]) + z * (G[5] + G[7]) + w * (G[2] - G[6]));
  vq[3] = 2.f * (x * (G[2] + G[6]) + y * (G[5] + G[7]) - 2.f * z * (G[0] + G[4]) + w * (G[3] - G[1]));
  float dotp = vq[0] * w + vq[1] * x + vq[2] * y + vq[3] * z;
  v_quat[0] += (vq[0] - dotp * w) * t.inv_norm;
  v_quat[1] += (vq[1] - dotp * x) * t.inv_norm;
  v_quat[2] += (vq[2] - dotp * y) * t.inv_norm;
  v_quat[3] += (vq[3] - dotp * z) * t.inv_norm;
}
