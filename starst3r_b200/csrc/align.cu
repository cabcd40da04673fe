// Fused sparse global alignment optimiser (ALIGN hot loop).
//
// Replaces optimize_loop of starster/reconstruct.py:371-406 with its nested make_K_cam_depth (:209-261),
// loss_3d (:325-353), loss_2d (:355-369), loss_dust3r (:311-323), make_pts3d / reproj2d
// (mast3r/cloud_opt/sparse_ga.py:469-501,977-981), gamma_loss (cloud_opt/utils/losses.py:19-28) and the
// torch.optim.Adam(lr=1, betas=(.9,.9)) step + quaternion re-normalisation (:373-395).
// The reference spends ~200 tiny autograd kernels and one host sync per iteration; here one iteration is
// three launches and the whole niter loop is enqueued without touching the host:
//   align_cam_fwd   (1 CTA)  raw parameters -> per-image camera record (R, t', f, cx, cy, depth affine map)
//   align_loss_*    (grid)   per-correspondence unproject -> transform -> gamma loss -> analytic gradient
//                            w.r.t. the two images' camera records (warp-shuffle + shared-memory reduction)
//   align_cam_bwd   (1 CTA)  camera-record gradients -> MST chain backward -> parameter gradients -> Adam
// HBM traffic per iteration is the correspondence list (~40 B per correspondence slot); the
// loop is latency-bound, so the figure of merit is iterations/s.
// ST3R_HOST_EMU: test builds that run this file on a CPU SIMT emulator (tests/host/): align_emu_host.cpp includes the
// kernels only, build_emu_lib.py (ST3R_EMU_WHOLE) compiles the entry points too, with their launches rewritten.
#include <vector>
#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
#include "common.cuh"
#endif
#ifndef ST3R_HOST_EMU
#define ST3R_DYN_SMEM(name) extern __shared__ float name[]
#endif
#include "align_math.cuh"
#include "../../include/starst3r_b200.h"

namespace {

constexpr int CAM_THREADS = 256;
constexpr int LOSS_THREADS = 256;
constexpr int NG = ALIGN_CAM_GRADS;
// Variant 1 of the loss kernels ("segmented", st3r_align_set_variant): replicas of the per-image gradient table in
// HBM (CTA b adds into replica b % ALIGN_REPL) and the ranges one warp walks.
constexpr int ALIGN_REPL = 8;
constexpr int SEG_MIN_PER_WARP = 128;   // entries per warp (4 rows of 32) before the grid stops growing
constexpr int SEG_MAX_CTAS = 592;       // 4 per SM on a 148-SM part
static_assert(sizeof(AlignImgConst) == sizeof(St3rAlignImgConst), "public / internal image record mismatch");

struct Params { float* pp; float* log_focal; float* quat; float* trans; float* log_size; };
struct AdamState { float* m; float* v; };   // laid out like the flat parameter vector [11 * N]

struct Work {
  AlignCam* cam; AlignCamTmp* tmp; AlignCamGrad* cgrad;
  float* gcam;        // [ALIGN_REPL][N * 17] (variant 0 uses the first replica only)
  float* sums;        // [4]: main loss numerator, dust3r numerator, spare, stop flag
  float* gscal;       // [2]: g (global scaling), number of images attaining the minimum size, the minimum size
};

// kStage (ALIGN variant bit 0, N <= CAM_STAGE_MAX): the per-image records the single-thread MST chain walks live in
// shared memory for the duration of the kernel.  In HBM every edge of the chain is a store -> load round trip through
// L2 (~10 us for 7 edges in the launch list of profiles/r01u); the serial scans also stop at min(N, block size).
constexpr int CAM_STAGE_MAX = 128;

template <bool kStage>
__global__ void __launch_bounds__(CAM_THREADS)
align_cam_fwd_kernel(St3rAlignProblem pb, Params p, Work w) {
  const int N = pb.n_img;
  if (w.sums[3] != 0.f) return;   // NaN loss seen: the reference breaks out of the loop (reconstruct.py:398-399)
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  __shared__ float s_min[CAM_THREADS];
  __shared__ AlignCamTmp s_tmp[kStage ? CAM_STAGE_MAX : 1];
  AlignCamTmp* tmp = kStage ? s_tmp : w.tmp;
  float best = INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    al_cam_local_fwd(ic[i], p.log_focal[i], p.log_size[i], p.quat + 4 * i, tmp[i]);
    float s = tmp[i].s;
    best = fminf(best, s);
  }
  s_min[threadIdx.x] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = INFINITY;
    const int nscan = kStage ? min(N, (int)blockDim.x) : (int)blockDim.x;   // threads >= N hold +inf
    for (int t = 0; t < nscan; ++t) b = fminf(b, s_min[t]);
    // torch's sizes.min() backward splits the gradient evenly between tied minima (all sizes tie at the
    // first iteration, when every log_size is 0), so remember how many images attain the minimum.
    int ties = 0;
    for (int i = 0; i < N; ++i) ties += (tmp[i].s == b) ? 1 : 0;
    w.gscal[0] = 1.0f / b;
    w.gscal[1] = (float)ties;
    w.gscal[2] = b;
    al_chain_fwd(N, pb.root, pb.edges, tmp, p.trans);
  }
  __syncthreads();
  const float g = w.gscal[0];
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    al_cam_final_fwd(ic[i], p.pp + 2 * i, g, tmp[i], w.cam[i]);
    if (kStage) w.tmp[i] = tmp[i];     // the backward kernel reads the records from HBM
  }
}

// Adds a thread's 17 camera-record gradients of image `img` into the CTA-wide shared table
// (warp-shuffle reduction when the whole warp works on the same image, which is the common case:
// correspondences of one image pair are contiguous).
__device__ __forceinline__ void accum_image(float* table, int img, bool active, const float* g) {
  const unsigned full = 0xffffffffu;
  int key = active ? img : -1;
  unsigned peers = __match_any_sync(full, key);
  if (peers == full) {
    if (key < 0) return;
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      float v = g[k];
      for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(full, v, off);
      if (lane_id() == 0 && v != 0.f) atomicAdd(table + img * NG + k, v);
    }
  } else if (active) {
#pragma unroll
    for (int k = 0; k < NG; ++k)
      if (g[k] != 0.f) atomicAdd(table + img * NG + k, g[k]);
  }
}

__device__ __forceinline__ void flush_table(const float* table, int N, float* gcam, float loss, float* loss_out) {
  __shared__ float red[LOSS_THREADS / 32];
  __syncthreads();
  for (int e = threadIdx.x; e < N * NG; e += blockDim.x) {
    float v = table[e];
    if (v != 0.f) atomicAdd(gcam + e, v);
  }
  for (int off = 16; off; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
  if (lane_id() == 0) red[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < LOSS_THREADS / 32; ++k) s += red[k];
    if (s != 0.f) atomicAdd(loss_out, s);
  }
}

struct Anchor { int img; float u, v, core, off; };
__device__ __forceinline__ Anchor load_anchor(const St3rAlignProblem& pb, const AlignImgConst* ic, int a) {
  Anchor r;
  r.img = pb.anc_img[a];
  r.u = pb.anc_uv[2 * a];
  r.v = pb.anc_uv[2 * a + 1];
  r.core = pb.core[ic[r.img].core_off + pb.anc_k[a]];
  r.off = pb.anc_off[a];
  return r;
}

// loss_3d: sum conf * gamma(|P1 - P2|) / norm   (reconstruct.py:325-353)
__global__ void __launch_bounds__(LOSS_THREADS)
align_loss3d_kernel(St3rAlignProblem pb, Work w, float gamma, float offset, float off_pow, float scale) {
  ST3R_DYN_SMEM(table);
  if (w.sums[3] != 0.f) return;
  const int N = pb.n_img;
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  for (int e = threadIdx.x; e < N * NG; e += blockDim.x) table[e] = 0.f;
  __syncthreads();
  float loss = 0.f;
  const int n = pb.n3;
  const int n_round = (n + 31) / 32 * 32;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n_round; m += gridDim.x * blockDim.x) {
    const bool active = m < n;
    float g1[NG], g2[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) g1[k] = g2[k] = 0.f;
    int i1 = 0, i2 = 0;
    if (active) {
      Anchor a1 = load_anchor(pb, ic, pb.e3_a1[m]), a2 = load_anchor(pb, ic, pb.e3_a2[m]);
      i1 = a1.img; i2 = a2.img;
      const AlignCam c1 = w.cam[i1], c2 = w.cam[i2];
      float P1[3], P2[3], pc1[3], pc2[3], z1, z2, D1, D2, o1, o2;
      al_anchor_point(c1, a1.u, a1.v, a1.core, a1.off, P1, pc1, &z1, &D1, &o1);
      al_anchor_point(c2, a2.u, a2.v, a2.core, a2.off, P2, pc2, &z2, &D2, &o2);
      float d[3] = {P1[0] - P2[0], P1[1] - P2[1], P1[2] - P2[2]};
      float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      float dl;
      float l = al_gamma_loss(dist, gamma, offset, off_pow, &dl);
      float cw = pb.e3_conf[m] * scale;
      loss += cw * l;
      float k = dist > 0.f ? cw * dl / dist : 0.f;
      float GP1[3] = {k * d[0], k * d[1], k * d[2]}, GP2[3] = {-k * d[0], -k * d[1], -k * d[2]};
      al_anchor_point_vjp(c1, a1.u, a1.v, a1.core, a1.off, pc1, z1, D1, o1, GP1, g1);
      al_anchor_point_vjp(c2, a2.u, a2.v, a2.core, a2.off, pc2, z2, D2, o2, GP2, g2);
    }
    accum_image(table, i1, active, g1);
    accum_image(table, i2, active, g2);
  }
  flush_table(table, N, w.gcam, loss, w.sums + 0);
}

// loss_2d: sum conf * gamma(|pix1 - reproj(K1 w2cam1, P2)|) / norm   (reconstruct.py:355-369)
__global__ void __launch_bounds__(LOSS_THREADS)
align_loss2d_kernel(St3rAlignProblem pb, Work w, float gamma, float offset, float off_pow, float scale) {
  ST3R_DYN_SMEM(table);
  if (w.sums[3] != 0.f) return;
  const int N = pb.n_img;
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  for (int e = threadIdx.x; e < N * NG; e += blockDim.x) table[e] = 0.f;
  __syncthreads();
  float loss = 0.f;
  const int n = pb.n2;
  const int n_round = (n + 31) / 32 * 32;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n_round; m += gridDim.x * blockDim.x) {
    const bool active = m < n;
    float g1[NG], g2[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) g1[k] = g2[k] = 0.f;
    int i1 = 0, i2 = 0;
    if (active) {
      i1 = pb.e2_img1[m];
      Anchor a2 = load_anchor(pb, ic, pb.e2_a2[m]);
      i2 = a2.img;
      const AlignCam c1 = w.cam[i1], c2 = w.cam[i2];
      float P2[3], pc2[3], z2, D2, o2;
      al_anchor_point(c2, a2.u, a2.v, a2.core, a2.off, P2, pc2, &z2, &D2, &o2);
      float uv[2];
      AlignReproj q;
      al_reproj(c1, P2, uv, q);
      float d[2] = {pb.e2_pix[2 * m] - uv[0], pb.e2_pix[2 * m + 1] - uv[1]};
      float dist = sqrtf(d[0] * d[0] + d[1] * d[1]);
      float dl;
      float l = al_gamma_loss(dist, gamma, offset, off_pow, &dl);
      float cw = pb.e2_conf[m] * scale;
      loss += cw * l;
      float k = dist > 0.f ? cw * dl / dist : 0.f;
      float Guv[2] = {-k * d[0], -k * d[1]};
      float GP[3];
      al_reproj_vjp(c1, P2, q, Guv, g1, GP);
      al_anchor_point_vjp(c2, a2.u, a2.v, a2.core, a2.off, pc2, z2, D2, o2, GP, g2);
    }
    accum_image(table, i1, active, g1);
    accum_image(table, i2, active, g2);
  }
  flush_table(table, N, w.gcam, loss, w.sums + 0);
}

// loss_dust3r: sum conf * gamma(|P1 - cam2w[img2] tgt|) / norm   (reconstruct.py:311-323)
__global__ void __launch_bounds__(LOSS_THREADS)
align_lossd_kernel(St3rAlignProblem pb, Work w, float gamma, float offset, float off_pow, float scale) {
  ST3R_DYN_SMEM(table);
  if (w.sums[3] != 0.f) return;
  const int N = pb.n_img;
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  for (int e = threadIdx.x; e < N * NG; e += blockDim.x) table[e] = 0.f;
  __syncthreads();
  float loss = 0.f;
  const int n = pb.nd;
  const int n_round = (n + 31) / 32 * 32;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n_round; m += gridDim.x * blockDim.x) {
    const bool active = m < n;
    float g1[NG], g2[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) g1[k] = g2[k] = 0.f;
    int i1 = 0, i2 = 0;
    if (active) {
      Anchor a1 = load_anchor(pb, ic, pb.ed_a1[m]);
      i1 = a1.img;
      i2 = pb.ed_img2[m];
      const AlignCam c1 = w.cam[i1], c2 = w.cam[i2];
      float P1[3], pc1[3], z1, D1, o1;
      al_anchor_point(c1, a1.u, a1.v, a1.core, a1.off, P1, pc1, &z1, &D1, &o1);
      const float tg[3] = {pb.ed_tgt[3 * m], pb.ed_tgt[3 * m + 1], pb.ed_tgt[3 * m + 2]};
      float T[3];
      al_mat3_vec(c2.R, tg, T);
      T[0] += c2.t[0]; T[1] += c2.t[1]; T[2] += c2.t[2];
      float d[3] = {P1[0] - T[0], P1[1] - T[1], P1[2] - T[2]};
      float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      float dl;
      float l = al_gamma_loss(dist, gamma, offset, off_pow, &dl);
      float cw = pb.ed_conf[m] * scale;
      loss += cw * l;
      float k = dist > 0.f ? cw * dl / dist : 0.f;
      float GP1[3] = {k * d[0], k * d[1], k * d[2]};
      al_anchor_point_vjp(c1, a1.u, a1.v, a1.core, a1.off, pc1, z1, D1, o1, GP1, g1);
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) g2[3 * a + b] += -GP1[a] * tg[b];
        g2[9 + a] += -GP1[a];
      }
    }
    accum_image(table, i1, active, g1);
    accum_image(table, i2, active, g2);
  }
  flush_table(table, N, w.gcam, loss, w.sums + 1);
}


// ---- variant 1: segmented accumulation -------------------------------------------------------------------------
// The kernels above reduce 2 x 17 gradient values across the warp for EVERY row of 32 entries (170 shuffles + 34
// shared-memory atomics) and every CTA of a ~1100-CTA grid then adds its table to the same N x 17 words of HBM
// (profiles/r01u_launches_reconstruct.csv: 92 / 109 us per launch for ~0.6 M entries, where the entry lists are a
// ~30 MB stream).  Entries of one image pair are contiguous (flatten_problem emits them slice by slice), so here a
// warp walks a contiguous range and keeps the two images' gradients in REGISTERS while the pair does not change -
// the vjp helpers accumulate straight into them - and reduces across lanes only when the pair changes or its range
// ends.  The grid is ~4x smaller and CTA b adds its table into replica b % ALIGN_REPL, which align_cam_bwd folds.
template <int KIND> struct SegEntry {
  Anchor a1, a2; int i1, i2;
  float x[3];       // KIND 1: the pixel in image 1; KIND 2: the regression target in image 2's camera frame
  float conf;
};

template <int KIND>
__device__ __forceinline__ void seg_load(const St3rAlignProblem& pb, const AlignImgConst* ic, int m, SegEntry<KIND>& e) {
  e.x[0] = e.x[1] = e.x[2] = 0.f;
  if (KIND == 0) {
    e.a1 = load_anchor(pb, ic, pb.e3_a1[m]); e.a2 = load_anchor(pb, ic, pb.e3_a2[m]);
    e.i1 = e.a1.img; e.i2 = e.a2.img;
    e.conf = pb.e3_conf[m];
  } else if (KIND == 1) {
    e.i1 = pb.e2_img1[m];
    e.a2 = load_anchor(pb, ic, pb.e2_a2[m]);
    e.i2 = e.a2.img;
    e.x[0] = pb.e2_pix[2 * m]; e.x[1] = pb.e2_pix[2 * m + 1];
    e.conf = pb.e2_conf[m];
  } else {
    e.a1 = load_anchor(pb, ic, pb.ed_a1[m]);
    e.i1 = e.a1.img;
    e.i2 = pb.ed_img2[m];
    e.x[0] = pb.ed_tgt[3 * m]; e.x[1] = pb.ed_tgt[3 * m + 1]; e.x[2] = pb.ed_tgt[3 * m + 2];
    e.conf = pb.ed_conf[m];
  }
}

// Loss of one entry; ADDS its gradients w.r.t. the two camera records into g1 / g2.  Same arithmetic as the bodies of
// align_loss3d / align_loss2d / align_lossd above.  `cams`: the camera table (global memory or a shared-memory copy).
template <int KIND>
__device__ __forceinline__ float seg_eval(const AlignCam* cams, const SegEntry<KIND>& e, float gamma, float offset,
                                          float off_pow, float scale, float* g1, float* g2) {
  const AlignCam c1 = cams[e.i1], c2 = cams[e.i2];
  float dl;
  const float cw = e.conf * scale;
  if (KIND == 0) {
    float P1[3], P2[3], pc1[3], pc2[3], z1, z2, D1, D2, o1, o2;
    al_anchor_point(c1, e.a1.u, e.a1.v, e.a1.core, e.a1.off, P1, pc1, &z1, &D1, &o1);
    al_anchor_point(c2, e.a2.u, e.a2.v, e.a2.core, e.a2.off, P2, pc2, &z2, &D2, &o2);
    const float d[3] = {P1[0] - P2[0], P1[1] - P2[1], P1[2] - P2[2]};
    const float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float l = al_gamma_loss(dist, gamma, offset, off_pow, &dl);
    const float k = dist > 0.f ? cw * dl / dist : 0.f;
    const float GP1[3] = {k * d[0], k * d[1], k * d[2]}, GP2[3] = {-k * d[0], -k * d[1], -k * d[2]};
    al_anchor_point_vjp(c1, e.a1.u, e.a1.v, e.a1.core, e.a1.off, pc1, z1, D1, o1, GP1, g1);
    al_anchor_point_vjp(c2, e.a2.u, e.a2.v, e.a2.core, e.a2.off, pc2, z2, D2, o2, GP2, g2);
    return cw * l;
  } else if (KIND == 1) {
    float P2[3], pc2[3], z2, D2, o2;
    al_anchor_point(c2, e.a2.u, e.a2.v, e.a2.core, e.a2.off, P2, pc2, &z2, &D2, &o2);
    float uv[2];
    AlignReproj q;
    al_reproj(c1, P2, uv, q);
    const float d[2] = {e.x[0] - uv[0], e.x[1] - uv[1]};
    const float dist = sqrtf(d[0] * d[0] + d[1] * d[1]);
    const float l = al_gamma_loss(dist, gamma, offset, off_pow, &dl);
    const float k = dist > 0.f ? cw * dl / dist : 0.f;
    const float Guv[2] = {-k * d[0], -k * d[1]};
    float GP[3];
    al_reproj_vjp(c1, P2, q, Guv, g1, GP);
    al_anchor_point_vjp(c2, e.a2.u, e.a2.v, e.a2.core, e.a2.off, pc2, z2, D2, o2, GP, g2);
    return cw * l;
  } else {
    float P1[3], pc1[3], z1, D1, o1;
    al_anchor_point(c1, e.a1.u, e.a1.v, e.a1.core, e.a1.off, P1, pc1, &z1, &D1, &o1);
    const float tg[3] = {e.x[0], e.x[1], e.x[2]};
    float T[3];
    al_mat3_vec(c2.R, tg, T);
    T[0] += c2.t[0]; T[1] += c2.t[1]; T[2] += c2.t[2];
    const float d[3] = {P1[0] - T[0], P1[1] - T[1], P1[2] - T[2]};
    const float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float l = al_gamma_loss(dist, gamma, offset, off_pow, &dl);
    const float k = dist > 0.f ? cw * dl / dist : 0.f;
    const float GP1[3] = {k * d[0], k * d[1], k * d[2]};
    al_anchor_point_vjp(c1, e.a1.u, e.a1.v, e.a1.core, e.a1.off, pc1, z1, D1, o1, GP1, g1);
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) g2[3 * a + b] += -GP1[a] * tg[b];
      g2[9 + a] += -GP1[a];
    }
    return cw * l;
  }
}

// Warp-wide sum of a lane-private gradient record into the CTA table; the record restarts at zero.
__device__ __forceinline__ void seg_flush(float* table, int img, float* acc) {
  if (img < 0) return;   // warp-uniform
  // All shuffles first, then ONE shared-memory update pass by 17 lanes: a compare-and-swap loop between two shuffles
  // would make the compiler wrap every later shuffle in a WARPSYNC / ENDCOLLECTIVE pair.
  const int lane = lane_id();
  float mine = 0.f;
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    float v = acc[k];
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == k) mine = v;
    acc[k] = 0.f;
  }
  if (lane < NG && mine != 0.f) atomicAdd(table + img * NG + lane, mine);
}

// One row of 32 entries into the register accumulators acc1 / acc2 of the image pair (cur1, cur2).  Entries are
// emitted pair by pair, so a row almost always belongs to ONE image pair; a row that straddles two (or more) pairs is
// walked once per pair, the other lanes redoing the pair's first entry with weight zero - the first version handed such a row to per-lane
// shared-memory atomics (34 CAS loops per lane, 32-way contended: ~100 k cycles for ONE row, measured with the
// per-phase counters of the persistent kernel, and every CTA of the grid waits for it at the barrier).
template <int KIND>
__device__ __forceinline__ float seg_row(const AlignCam* cams, float* table, SegEntry<KIND>& e, float gamma, float offset,
                                         float off_pow, float sc, float* acc1, float* acc2, int& cur1, int& cur2) {
  const int key = (e.i1 << 16) | e.i2;                       // n_img < 2^15 (the shared table caps it far lower)
  unsigned todo = 0xffffffffu;
  float loss = 0.f;
  while (todo) {                                             // warp-uniform: one pass per distinct image pair of the row
    const int leader = __ffs(todo) - 1;
    const int k = __shfl_sync(0xffffffffu, key, leader);
    const bool mine = key == k;
    const unsigned grp = __ballot_sync(0xffffffffu, mine);
    todo &= ~grp;
    const int f1 = k >> 16, f2 = k & 0xffff;
    if (f1 != cur1 || f2 != cur2) {
      seg_flush(table, cur1, acc1);
      seg_flush(table, cur2, acc2);
      cur1 = f1; cur2 = f2;
    }
    SegEntry<KIND> g = e;
    float wgt = sc;
    if (grp != 0xffffffffu) {
      // the lanes of another pair redo the leader's entry with weight zero: arithmetic that is known to be finite with
      // this pair's cameras (their own entry, seen from the wrong camera, could divide by a zero depth: 0 x inf = NaN)
      SegEntry<KIND> l;
      l.a1.u = __shfl_sync(0xffffffffu, e.a1.u, leader); l.a1.v = __shfl_sync(0xffffffffu, e.a1.v, leader);
      l.a1.core = __shfl_sync(0xffffffffu, e.a1.core, leader); l.a1.off = __shfl_sync(0xffffffffu, e.a1.off, leader);
      l.a2.u = __shfl_sync(0xffffffffu, e.a2.u, leader); l.a2.v = __shfl_sync(0xffffffffu, e.a2.v, leader);
      l.a2.core = __shfl_sync(0xffffffffu, e.a2.core, leader); l.a2.off = __shfl_sync(0xffffffffu, e.a2.off, leader);
      l.x[0] = __shfl_sync(0xffffffffu, e.x[0], leader); l.x[1] = __shfl_sync(0xffffffffu, e.x[1], leader);
      l.x[2] = __shfl_sync(0xffffffffu, e.x[2], leader);
      l.conf = __shfl_sync(0xffffffffu, e.conf, leader);
      l.a1.img = f1; l.a2.img = f2;
      if (!mine) { g = l; wgt = 0.f; }
    }
    g.i1 = f1; g.i2 = f2;
    loss += seg_eval<KIND>(cams, g, gamma, offset, off_pow, wgt, acc1, acc2);
  }
  return loss;
}

template <int KIND>
__global__ void __launch_bounds__(LOSS_THREADS)
align_loss_seg_kernel(St3rAlignProblem pb, Work w, float gamma, float offset, float off_pow, float scale, int per_warp,
                      float* loss_out) {
  ST3R_DYN_SMEM(table);
  if (w.sums[3] != 0.f) return;
  const int N = pb.n_img;
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  for (int e = threadIdx.x; e < N * NG; e += blockDim.x) table[e] = 0.f;
  __syncthreads();
  const int n = KIND == 0 ? pb.n3 : (KIND == 1 ? pb.n2 : pb.nd);
  const int lane = lane_id();
  const long long wbeg = ((long long)blockIdx.x * (LOSS_THREADS / 32) + (threadIdx.x >> 5)) * per_warp;
  // redux.sync / votes return warp-uniform values in a form the compiler's divergence analysis accepts, so the
  // shuffles of seg_flush below are not wrapped in WARPSYNC / ENDCOLLECTIVE pairs
  const int begin = __reduce_max_sync(0xffffffffu, (int)(wbeg < n ? wbeg : n));
  const int end = __reduce_max_sync(0xffffffffu, (int)(wbeg + per_warp < n ? wbeg + per_warp : n));
  float acc1[NG], acc2[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) acc1[k] = acc2[k] = 0.f;
  int cur1 = -1, cur2 = -1;      // image pair whose gradients acc1 / acc2 hold (warp-uniform)
  float loss = 0.f;
  for (int base = begin; base < end; base += 32) {
    // Lanes past the end of the range (last row only) redo the range's last entry with weight zero: the loop body has
    // no lane-dependent branch, so the warp provably stays converged for the shuffles of seg_flush.
    const bool active = base + lane < end;
    const int m = active ? base + lane : end - 1;
    const float sc = active ? scale : 0.f;
    SegEntry<KIND> e;
    seg_load<KIND>(pb, ic, m, e);
    loss += seg_row<KIND>(w.cam, table, e, gamma, offset, off_pow, sc, acc1, acc2, cur1, cur2);
  }
  seg_flush(table, cur1, acc1);
  seg_flush(table, cur2, acc2);
  flush_table(table, N, w.gcam + (size_t)(blockIdx.x % ALIGN_REPL) * N * NG, loss, loss_out);
}

__device__ __forceinline__ void adam_update(float* p, float g, float* m, float* v, float lr_over_bc1, float inv_sqrt_bc2,
                                            float omb1, float b2, float omb2, float eps) {
  float mm = *m + (g - *m) * omb1;
  float vv = *v * b2 + omb2 * g * g;
  *p = *p - lr_over_bc1 * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  *m = mm;
  *v = vv;
}

// camera-record gradients -> parameter gradients -> Adam -> quaternion re-normalisation; also finalises the
// iteration's loss and clears the accumulators for the next iteration.
template <bool kStage>
__global__ void __launch_bounds__(CAM_THREADS)
align_cam_bwd_kernel(St3rAlignProblem pb, Params p, AdamState ad, Work w, int train_mask, float lr_over_bc1,
                     float inv_sqrt_bc2, float omb1, float b2, float omb2, float eps, float dust3r_w,
                     float* loss_hist, int iter, float* grad_out /* optional [11 N], for tests */, int reps) {
  const int N = pb.n_img;
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  __shared__ float s_gg[CAM_THREADS];
  __shared__ AlignCamTmp s_tmp[kStage ? CAM_STAGE_MAX : 1];
  __shared__ AlignCamGrad s_cg[kStage ? CAM_STAGE_MAX : 1];
  if (w.sums[3] != 0.f) return;
  AlignCamTmp* tmp = kStage ? s_tmp : w.tmp;
  AlignCamGrad* cg = kStage ? s_cg : w.cgrad;
  const float g = w.gscal[0];
  const float ties = w.gscal[1], smin = w.gscal[2];
  float gg = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if (kStage) tmp[i] = w.tmp[i];
    const float* gsrc = w.gcam + i * NG;
    float folded[NG];
    if (reps > 1) {     // variant 1: the loss CTAs spread their sums over `reps` replicas of the table
#pragma unroll
      for (int k = 0; k < NG; ++k) folded[k] = gsrc[k];
      for (int r = 1; r < reps; ++r)
#pragma unroll
        for (int k = 0; k < NG; ++k) folded[k] += gsrc[(size_t)r * N * NG + k];
      gsrc = folded;
    }
    al_cam_final_bwd(ic[i], p.pp + 2 * i, g, tmp[i], gsrc, cg[i]);
    gg += cg[i].g_g;
  }
  s_gg[threadIdx.x] = gg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    const int nscan = kStage ? min(N, (int)blockDim.x) : (int)blockDim.x;   // threads >= N hold zero
    for (int k = 0; k < nscan; ++k) t += s_gg[k];
    s_gg[0] = t;
    al_chain_bwd(N, pb.root, pb.edges, tmp, p.trans, cg);
    float loss = w.sums[0] + dust3r_w * w.sums[1];
    if (loss_hist) loss_hist[iter] = loss;
    if (loss != loss) w.sums[3] = 1.0f;   // NaN -> stop flag (reconstruct.py:398-399)
    w.sums[0] = 0.f; w.sums[1] = 0.f;
  }
  __syncthreads();
  const float gg_total = s_gg[0];
  const bool stop = false;  // the NaN iteration itself still steps, like the reference
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float g_pp[2], g_lf, g_q[4], g_t[3], g_ls;
    float extra = (tmp[i].s == smin) ? -gg_total * g * g / ties : 0.f;
    al_cam_local_bwd(tmp[i], cg[i], extra, g_pp, &g_lf, g_q, g_t, &g_ls);
    if (grad_out) {
      float* o = grad_out + 11 * i;
      o[0] = g_pp[0]; o[1] = g_pp[1]; o[2] = g_lf; o[3] = g_q[0]; o[4] = g_q[1]; o[5] = g_q[2]; o[6] = g_q[3];
      o[7] = g_t[0]; o[8] = g_t[1]; o[9] = g_t[2]; o[10] = g_ls;
    }
    if (!stop) {
      float* m = ad.m + 11 * i;
      float* v = ad.v + 11 * i;
      if (train_mask & 1) {
        adam_update(p.pp + 2 * i, g_pp[0], m + 0, v + 0, lr_over_bc1, inv_sqrt_bc2, omb1, b2, omb2, eps);
        adam_update(p.pp + 2 * i + 1, g_pp[1], m + 1, v + 1, lr_over_bc1, inv_sqrt_bc2, omb1, b2, omb2, eps);
      }
      if (train_mask & 2) adam_update(p.log_focal + i, g_lf, m + 2, v + 2, lr_over_bc1, inv_sqrt_bc2, omb1, b2, omb2, eps);
      if (train_mask & 4) {
        for (int k = 0; k < 4; ++k)
          adam_update(p.quat + 4 * i + k, g_q[k], m + 3 + k, v + 3 + k, lr_over_bc1, inv_sqrt_bc2, omb1, b2, omb2, eps);
      }
      if (train_mask & 8) {
        for (int k = 0; k < 3; ++k)
          adam_update(p.trans + 3 * i + k, g_t[k], m + 7 + k, v + 7 + k, lr_over_bc1, inv_sqrt_bc2, omb1, b2, omb2, eps);
      }
      if (train_mask & 16) adam_update(p.log_size + i, g_ls, m + 10, v + 10, lr_over_bc1, inv_sqrt_bc2, omb1, b2, omb2, eps);
      // quats[i].data[:] /= quats[i].data.norm()   (reconstruct.py:394-395)
      float* q = p.quat + 4 * i;
      float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
    }
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < NG; ++k) w.gcam[(size_t)r * N * NG + i * NG + k] = 0.f;
  }
}

// ---- variant 4 (bit 2): the whole optimisation loop in ONE cooperative launch ------------------------------------
// Launch-per-iteration (above): three dependent launches per iteration, two of them single-CTA kernels whose work is
// a serial chain over the images, and loss kernels that chase four dependent loads per entry (entry -> anchor ->
// image record -> core depth): 89 / 106 us per coarse / fine iteration at 8 views 512 x 512 (profiles/r02w), for
// ~25 MB of entry data that should stream in a few microseconds.  Here:
//   * align_pack_kernel resolves the indirections ONCE per call (anchors and core depths are constants of the problem):
//     three float4 per entry, read with coalesced 128-bit loads, the next row prefetched while the current one is
//     evaluated;
//   * every CTA keeps its OWN copy of the optimiser state (parameters, Adam moments, camera records) in shared memory
//     and runs the camera forward, the MST chain, the camera backward and Adam redundantly - they are O(N) and
//     deterministic, so all copies stay bit-identical without any exchange;
//   * the only exchange per iteration is the per-image gradient table: each CTA writes the sums of its slice of the
//     entries to its own row of a [2][grid][17 N + 2] buffer, one grid-wide barrier (a monotonic counter in global
//     memory; the launch is cooperative, so all CTAs are resident), and every CTA adds the rows in the same order.
//     The two halves of the buffer alternate between iterations, which makes that single barrier sufficient.
// N <= PERSIST_MAX_IMG (the shared-memory copies are static arrays); larger problems take the path above.
constexpr int PERSIST_THREADS = 512;
constexpr int PERSIST_MAX_IMG = 64;

// Packed entry: a = anchor 1 (u, v, core, off) [KIND 1: the pixel], b = anchor 2 [KIND 2: the regression target],
// x = (conf, bits of (img1 << 16 | img2), -, -).
template <int KIND>
__global__ void __launch_bounds__(256)
align_pack_kernel(St3rAlignProblem pb, float4* __restrict__ out, int n) {
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  SegEntry<KIND> e;
  seg_load<KIND>(pb, ic, m, e);
  float4 a, b;
  if (KIND == 1) a = make_float4(e.x[0], e.x[1], 0.f, 0.f);
  else a = make_float4(e.a1.u, e.a1.v, e.a1.core, e.a1.off);
  if (KIND == 2) b = make_float4(e.x[0], e.x[1], e.x[2], 0.f);
  else b = make_float4(e.a2.u, e.a2.v, e.a2.core, e.a2.off);
  out[m] = a;
  out[(size_t)n + m] = b;
  out[2 * (size_t)n + m] = make_float4(e.conf, __int_as_float((e.i1 << 16) | e.i2), 0.f, 0.f);
}

template <int KIND>
__device__ __forceinline__ void seg_unpack(const float4& a, const float4& b, const float4& x, SegEntry<KIND>& e) {
  const int key = __float_as_int(x.y);
  e.i1 = key >> 16; e.i2 = key & 0xffff;
  e.conf = x.x;
  e.x[0] = e.x[1] = e.x[2] = 0.f;
  if (KIND == 1) { e.x[0] = a.x; e.x[1] = a.y; }
  else { e.a1.img = e.i1; e.a1.u = a.x; e.a1.v = a.y; e.a1.core = a.z; e.a1.off = a.w; }
  if (KIND == 2) { e.x[0] = b.x; e.x[1] = b.y; e.x[2] = b.z; }
  else { e.a2.img = e.i2; e.a2.u = b.x; e.a2.v = b.y; e.a2.core = b.z; e.a2.off = b.w; }
}

struct PersistArgs {
  St3rAlignProblem pb; Params p; AdamState ad; Work w;
  const float4* pk_main; int n_main;          // packed entries of the main loss (KIND 0 or 1)
  const float4* pk_d; int n_d;                // packed entries of the dust3r term (KIND 2)
  float gamma, off_m, offp_m, scale_main, gamma_d, off_d, offp_d, scale_d;
  int train_mask; float omb1, b2, omb2, eps;
  const float2* sched;                        // [niter]: (lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t))
  int niter; float* loss_hist; float* grad_out;
  float* partial;                             // [2][grid][17 N + 2]
  unsigned* bar;                              // grid barrier counter (zero at launch)
};

#ifndef ST3R_HOST_EMU
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                          // this CTA's partial sums are visible before its arrival
    atomicAdd(bar, 1u);
    unsigned v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spins > (1u << 27)) __trap();     // a protocol bug becomes a CUDA error instead of a hung GPU
    } while (v < target);
  }
  __syncthreads();
}
__device__ __forceinline__ float ld_l2(const float* p) { return __ldcg(p); }     // L2 is the point of coherence between SMs
#else
__device__ __forceinline__ void grid_barrier(unsigned*, unsigned) { emu::cluster_sync(); }   // the grid runs as one emulated cluster
__device__ __forceinline__ float ld_l2(const float* p) { return *p; }
#endif

// One warp walks rows [begin, end) of a packed entry array (32 entries per row) with the segmented register
// accumulation of align_loss_seg_kernel; returns the lane's loss share.
template <int KIND>
__device__ __forceinline__ float persist_walk(const float4* __restrict__ pk, int n, int begin, int end, const AlignCam* cams,
                                              float* table, float gamma, float offset, float off_pow, float scale) {
  const int lane = lane_id();
  float acc1[NG], acc2[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) acc1[k] = acc2[k] = 0.f;
  int cur1 = -1, cur2 = -1;
  float loss = 0.f;
  float4 na, nb, nx;
  {
    const int m = begin + lane < end ? begin + lane : end - 1;
    if (begin < end) { na = pk[m]; nb = pk[(size_t)n + m]; nx = pk[2 * (size_t)n + m]; }
  }
  for (int base = begin; base < end; base += 32) {
    const bool active = base + lane < end;
    const float4 a = na, b = nb, x = nx;
    if (base + 32 < end) {                     // prefetch the next row while this one is evaluated
      const int m = base + 32 + lane < end ? base + 32 + lane : end - 1;
      na = pk[m]; nb = pk[(size_t)n + m]; nx = pk[2 * (size_t)n + m];
    }
    const float sc = active ? scale : 0.f;     // lanes past the end redo the last entry with weight zero
    SegEntry<KIND> e;
    seg_unpack<KIND>(a, b, x, e);
    loss += seg_row<KIND>(cams, table, e, gamma, offset, off_pow, sc, acc1, acc2, cur1, cur2);
  }
  seg_flush(table, cur1, acc1);
  seg_flush(table, cur2, acc2);
  return loss;
}

// The MST chain of al_chain_fwd / al_chain_bwd with the 9 + 3 elements of an edge spread over the lanes of ONE warp
// (same expressions element by element): an edge costs one shared-memory round trip instead of ~100 dependent ones in
// a single thread - the chains were 7 k + 10 k of the persistent kernel's 170 k cycles per iteration at 8 images.
__device__ __forceinline__ void chain_fwd_warp(int N, int root, const int* edges, AlignCamTmp* tmp, const float* trans) {
  const int l = lane_id();
  if (l < 9) tmp[root].TR[l] = tmp[root].relR[l];
  else if (l < 12) tmp[root].Tt[l - 9] = trans[3 * root + l - 9];
  __syncwarp();
  for (int e = 0; e < N - 1; ++e) {
    const int i = edges[2 * e], j = edges[2 * e + 1];
    const float* A = tmp[i].TR;
    if (l < 9) {
      const int a = l / 3, b = l - 3 * a;
      const float* B = tmp[j].relR;
      tmp[j].TR[l] = A[3 * a] * B[b] + A[3 * a + 1] * B[3 + b] + A[3 * a + 2] * B[6 + b];
    } else if (l < 12) {
      const int k = l - 9;
      const float* v = trans + 3 * j;
      tmp[j].Tt[k] = (A[3 * k] * v[0] + A[3 * k + 1] * v[1] + A[3 * k + 2] * v[2]) + tmp[i].Tt[k];
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void chain_bwd_warp(int N, const int* edges, const AlignCamTmp* tmp, const float* trans,
                                               AlignCamGrad* grads) {
  const int l = lane_id();
  for (int e = N - 2; e >= 0; --e) {
    const int i = edges[2 * e], j = edges[2 * e + 1];
    const float* GTRj = grads[j].GTR;
    const float* GTtj = grads[j].GTt;
    float up = 0.f, down = 0.f;      // new value of the parent's / the child's element this lane owns
    if (l < 9) {
      const int a = l / 3, b = l - 3 * a;
      float s = GTtj[a] * trans[3 * j + b];
      for (int k = 0; k < 3; ++k) s += GTRj[3 * a + k] * tmp[j].relR[3 * b + k];
      up = grads[i].GTR[l] + s;
      down = tmp[i].TR[a] * GTRj[b] + tmp[i].TR[3 + a] * GTRj[3 + b] + tmp[i].TR[6 + a] * GTRj[6 + b];
    } else if (l < 12) {
      const int k = l - 9;
      up = grads[i].GTt[k] + GTtj[k];
      down = tmp[i].TR[k] * GTtj[0] + tmp[i].TR[3 + k] * GTtj[1] + tmp[i].TR[6 + k] * GTtj[2];
    }
    __syncwarp();                    // everybody has read the child's old gradients
    if (l < 9) { grads[i].GTR[l] = up; grads[j].GTR[l] = down; }
    else if (l < 12) { grads[i].GTt[l - 9] = up; grads[j].GTt[l - 9] = down; }
    __syncwarp();
  }
}

// The CTA's copy of the optimiser state.  (File scope rather than inside the kernel template: the CPU emulator of
// tests/host keeps one copy of the shared-memory section per CTA, and function-local statics of a template are
// emitted outside that section.)
__shared__ AlignCam s_cam[PERSIST_MAX_IMG];
__shared__ AlignCamTmp s_tmp[PERSIST_MAX_IMG];
__shared__ AlignCamGrad s_cg[PERSIST_MAX_IMG];
__shared__ float s_par[11 * PERSIST_MAX_IMG], s_m[11 * PERSIST_MAX_IMG], s_v[11 * PERSIST_MAX_IMG];
__shared__ float s_table[NG * PERSIST_MAX_IMG + 32], s_gsum[NG * PERSIST_MAX_IMG];   // (the table doubles as the reduction scratch)
__shared__ int s_edges[2 * PERSIST_MAX_IMG];
__shared__ float s_red[PERSIST_THREADS / 32][2];
__shared__ float s_scal[8];                    // g, ties, smin, loss main, loss dust3r, stop flag, sum of g_g

template <int KIND_MAIN>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
align_persist_kernel(const PersistArgs a) {
  const St3rAlignProblem& pb = a.pb;
  const int N = pb.n_img, tid = threadIdx.x, nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  // shared copy of the optimiser state: [pp 2N | log_focal N | quat 4N | trans 3N | log_size N], moments [N][11]
  float* s_pp = s_par, *s_lf = s_par + 2 * N, *s_q = s_par + 3 * N, *s_tr = s_par + 7 * N, *s_ls = s_par + 10 * N;
  for (int i = tid; i < 2 * N; i += nthr) s_pp[i] = a.p.pp[i];
  for (int i = tid; i < N; i += nthr) { s_lf[i] = a.p.log_focal[i]; s_ls[i] = a.p.log_size[i]; }
  for (int i = tid; i < 4 * N; i += nthr) s_q[i] = a.p.quat[i];
  for (int i = tid; i < 3 * N; i += nthr) s_tr[i] = a.p.trans[i];
  for (int i = tid; i < 11 * N; i += nthr) { s_m[i] = a.ad.m[i]; s_v[i] = a.ad.v[i]; }
  for (int i = tid; i < 2 * (N - 1); i += nthr) s_edges[i] = pb.edges[i];
  if (tid == 0) s_scal[5] = 0.f;
  // this CTA's slice of the entry rows, split over its warps
  const long long gw = (long long)blockIdx.x * nwarps + warp, tw = (long long)gridDim.x * nwarps;
  auto slice = [&](int n, int& begin, int& end) {
    const long long per = ((n + tw - 1) / tw + 31) / 32 * 32;
    const long long b0 = gw * per, e0 = b0 + per;
    begin = __reduce_max_sync(0xffffffffu, (int)(b0 < n ? b0 : n));       // (warp-uniform values the compiler can see)
    end = __reduce_max_sync(0xffffffffu, (int)(e0 < n ? e0 : n));
  };
  int mb, me, db, de;
  slice(a.n_main, mb, me);
  slice(a.n_d, db, de);
  const int row_len = NG * N + 2;
  __syncthreads();

#ifdef ALIGN_PERSIST_TIMING      // development aid (scripts/build_dbg.sh): cycles per phase seen by thread 0 of the first / last CTA
  long long tc[6] = {0, 0, 0, 0, 0, 0};
#define PT_MARK(k) do { if (tid == 0) { const long long now_ = clock64(); tc[k] += now_ - tlast; tlast = now_; } } while (0)
  long long tlast = clock64();
#else
#define PT_MARK(k) do { } while (0)
#endif
  for (int it = 0; it < a.niter; ++it) {
    if (s_scal[5] != 0.f) break;               // NaN loss seen (identical in every CTA): reconstruct.py:398-399
    // ---- camera forward (redundant in every CTA)
    float best = INFINITY;
    for (int i = tid; i < N; i += nthr) {
      al_cam_local_fwd(ic[i], s_lf[i], s_ls[i], s_q + 4 * i, s_tmp[i]);
      best = fminf(best, s_tmp[i].s);
    }
    for (int off = 16; off; off >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, off));
    if (lane == 0) s_red[warp][0] = best;
    for (int e = tid; e < NG * N; e += nthr) s_table[e] = 0.f;
    __syncthreads();
    if (warp == 0) {
      float b = lane < nwarps ? s_red[lane][0] : INFINITY;
      for (int off = 16; off; off >>= 1) b = fminf(b, __shfl_xor_sync(0xffffffffu, b, off));
      int ties = 0;                            // torch's min() backward splits the gradient between tied minima
      for (int i = lane; i < N; i += 32) ties += (s_tmp[i].s == b) ? 1 : 0;
      ties = __reduce_add_sync(0xffffffffu, ties);
      if (lane == 0) { s_scal[0] = 1.0f / b; s_scal[1] = (float)ties; s_scal[2] = b; }
      chain_fwd_warp(N, pb.root, s_edges, s_tmp, s_tr);
    }
    __syncthreads();
    const float g = s_scal[0];
    for (int i = tid; i < N; i += nthr) al_cam_final_fwd(ic[i], s_pp + 2 * i, g, s_tmp[i], s_cam[i]);
    __syncthreads();
    if (blockIdx.x == 0)                       // the records of the LAST forward are the result (reconstruct.py:379-380)
      for (int i = tid; i < N * (int)(sizeof(AlignCam) / 4); i += nthr)
        reinterpret_cast<float*>(a.w.cam)[i] = reinterpret_cast<const float*>(s_cam)[i];
    PT_MARK(0);
    // ---- loss + gradient table of this CTA's slice
    float l_main = 0.f, l_d = 0.f;
    if (a.n_main > 0)
      l_main = persist_walk<KIND_MAIN>(a.pk_main, a.n_main, mb, me, s_cam, s_table, a.gamma, a.off_m, a.offp_m, a.scale_main);
    if (a.n_d > 0)
      l_d = persist_walk<2>(a.pk_d, a.n_d, db, de, s_cam, s_table, a.gamma_d, a.off_d, a.offp_d, a.scale_d);
    for (int off = 16; off; off >>= 1) {
      l_main += __shfl_xor_sync(0xffffffffu, l_main, off);
      l_d += __shfl_xor_sync(0xffffffffu, l_d, off);
    }
    if (lane == 0) { s_red[warp][0] = l_main; s_red[warp][1] = l_d; }
    __syncthreads();
    float* mine = a.partial + ((size_t)(it & 1) * gridDim.x + blockIdx.x) * row_len;
    for (int e = tid; e < NG * N; e += nthr) mine[e] = s_table[e];
    if (tid == 0) {
      float s0 = 0.f, s1 = 0.f;
      for (int k = 0; k < nwarps; ++k) { s0 += s_red[k][0]; s1 += s_red[k][1]; }
      mine[NG * N] = s0; mine[NG * N + 1] = s1;
    }
    PT_MARK(1);
    grid_barrier(a.bar, (unsigned)(it + 1) * gridDim.x);
    PT_MARK(2);
    // ---- every CTA adds the rows in the same order: identical sums everywhere
    // The row's values are cut into chunks of 32 (one per lane: coalesced 128-byte reads of L2) and the warps of a
    // chunk share the CTAs' rows; eight loads are in flight per lane, the additions keep the row order.  (First
    // version: one thread per value walking all rows - a chain of gridDim.x dependent L2 round trips, 22 k cycles per
    // iteration; lanes across ROWS instead made every 4-byte read fetch its own 32-byte sector: 50 k.)
    const float* rows = a.partial + (size_t)(it & 1) * gridDim.x * row_len;
    const int nchunks = (row_len + 31) / 32;
    const int G = nwarps / nchunks > 0 ? nwarps / nchunks : 1;          // warps per chunk
    float* part = s_table;                                              // [G][row_len] scratch (the table is published)
    for (int w2 = warp; w2 < G * nchunks; w2 += nwarps) {
      const int chunk = w2 % nchunks, g2 = w2 / nchunks, e = chunk * 32 + lane;
      float v = 0.f;
      if (e < row_len) {
        unsigned c = g2;
        for (; c + 7 * G < gridDim.x; c += 8 * G) {
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] = ld_l2(rows + (size_t)(c + j * G) * row_len + e);
#pragma unroll
          for (int j = 0; j < 8; ++j) v += t[j];
        }
        for (; c < gridDim.x; c += G) v += ld_l2(rows + (size_t)c * row_len + e);
        part[g2 * row_len + e] = v;
      }
    }
    __syncthreads();
    for (int e = tid; e < row_len; e += nthr) {
      float v = part[e];
      for (int g2 = 1; g2 < G; ++g2) v += part[g2 * row_len + e];
      if (e < NG * N) s_gsum[e] = v; else s_scal[3 + (e - NG * N)] = v;
    }
    __syncthreads();
    PT_MARK(3);
    // ---- camera backward + Adam (redundant in every CTA)
    float gg = 0.f;
    for (int i = tid; i < N; i += nthr) {
      al_cam_final_bwd(ic[i], s_pp + 2 * i, g, s_tmp[i], s_gsum + i * NG, s_cg[i]);
      gg += s_cg[i].g_g;
    }
    for (int off = 16; off; off >>= 1) gg += __shfl_xor_sync(0xffffffffu, gg, off);
    if (lane == 0) s_red[warp][0] = gg;
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        float t = 0.f;
        for (int k = 0; k < nwarps; ++k) t += s_red[k][0];
        s_scal[6] = t;
        const float loss = s_scal[3] + s_scal[4];          // the dust3r weight is folded into scale_d
        if (blockIdx.x == 0 && a.loss_hist) a.loss_hist[it] = loss;
        if (loss != loss) s_scal[5] = 1.0f;                 // NaN: this iteration still steps, the next one stops
      }
      chain_bwd_warp(N, s_edges, s_tmp, s_tr, s_cg);
    }
    __syncthreads();
    const float gg_total = s_scal[6], ties = s_scal[1], smin = s_scal[2];
    const float2 sch = a.sched[it];
    for (int i = tid; i < N; i += nthr) {
      float g_pp[2], g_lf, g_q[4], g_t[3], g_ls;
      const float extra = (s_tmp[i].s == smin) ? -gg_total * g * g / ties : 0.f;
      al_cam_local_bwd(s_tmp[i], s_cg[i], extra, g_pp, &g_lf, g_q, g_t, &g_ls);
      if (a.grad_out && blockIdx.x == 0 && it == a.niter - 1) {
        float* o = a.grad_out + 11 * i;
        o[0] = g_pp[0]; o[1] = g_pp[1]; o[2] = g_lf; o[3] = g_q[0]; o[4] = g_q[1]; o[5] = g_q[2]; o[6] = g_q[3];
        o[7] = g_t[0]; o[8] = g_t[1]; o[9] = g_t[2]; o[10] = g_ls;
      }
      float* m = s_m + 11 * i;
      float* v = s_v + 11 * i;
      if (a.train_mask & 1) {
        adam_update(s_pp + 2 * i, g_pp[0], m + 0, v + 0, sch.x, sch.y, a.omb1, a.b2, a.omb2, a.eps);
        adam_update(s_pp + 2 * i + 1, g_pp[1], m + 1, v + 1, sch.x, sch.y, a.omb1, a.b2, a.omb2, a.eps);
      }
      if (a.train_mask & 2) adam_update(s_lf + i, g_lf, m + 2, v + 2, sch.x, sch.y, a.omb1, a.b2, a.omb2, a.eps);
      if (a.train_mask & 4)
        for (int k = 0; k < 4; ++k) adam_update(s_q + 4 * i + k, g_q[k], m + 3 + k, v + 3 + k, sch.x, sch.y, a.omb1, a.b2, a.omb2, a.eps);
      if (a.train_mask & 8)
        for (int k = 0; k < 3; ++k) adam_update(s_tr + 3 * i + k, g_t[k], m + 7 + k, v + 7 + k, sch.x, sch.y, a.omb1, a.b2, a.omb2, a.eps);
      if (a.train_mask & 16) adam_update(s_ls + i, g_ls, m + 10, v + 10, sch.x, sch.y, a.omb1, a.b2, a.omb2, a.eps);
      float* q = s_q + 4 * i;                                // quats[i].data[:] /= quats[i].data.norm()  (:394-395)
      const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
    }
    __syncthreads();
    PT_MARK(4);
  }
#ifdef ALIGN_PERSIST_TIMING
  if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1 || blockIdx.x == gridDim.x / 2))
    printf("align_persist cta %d niter %d cycles/iter: cam_fwd %lld walk+publish %lld barrier %lld reduce %lld bwd+adam %lld\n",
           (int)blockIdx.x, a.niter, tc[0] / a.niter, tc[1] / a.niter, tc[2] / a.niter, tc[3] / a.niter, tc[4] / a.niter);
#endif
  if (blockIdx.x == 0) {                                      // the state after the last step goes back to the caller
    for (int i = tid; i < 2 * N; i += nthr) a.p.pp[i] = s_pp[i];
    for (int i = tid; i < N; i += nthr) { a.p.log_focal[i] = s_lf[i]; a.p.log_size[i] = s_ls[i]; }
    for (int i = tid; i < 4 * N; i += nthr) a.p.quat[i] = s_q[i];
    for (int i = tid; i < 3 * N; i += nthr) a.p.trans[i] = s_tr[i];
    for (int i = tid; i < 11 * N; i += nthr) { a.ad.m[i] = s_m[i]; a.ad.v[i] = s_v[i]; }
    if (tid == 0 && s_scal[5] != 0.f) a.w.sums[3] = 1.0f;
  }
}

// pts3d of every anchor + dense depth maps from the last forward's camera records.
__global__ void align_outputs_kernel(St3rAlignProblem pb, Work w, float* pts3d, float* depthmaps, int n_core_total) {
  const AlignImgConst* ic = reinterpret_cast<const AlignImgConst*>(pb.img_const);
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < pb.n_anchor) {
    Anchor a = load_anchor(pb, ic, t);
    float P[3], pc[3], z, D, o;
    al_anchor_point(w.cam[a.img], a.u, a.v, a.core, a.off, P, pc, &z, &D, &o);
    pts3d[3 * t] = P[0]; pts3d[3 * t + 1] = P[1]; pts3d[3 * t + 2] = P[2];
  }
  if (t < n_core_total) {
    // find owning image (N is small)
    int img = 0;
    for (int i = 0; i < pb.n_img; ++i)
      if (t >= ic[i].core_off && t < ic[i].core_off + ic[i].n_core) img = i;
    depthmaps[t] = w.cam[img].A + w.cam[img].B * pb.core[t];
  }
}

#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
Work carve(void* ws, int N) {
  WsAlloc a(ws, (size_t)-1);
  Work w;
  w.cam = a.take<AlignCam>(N);
  w.tmp = a.take<AlignCamTmp>(N);
  w.cgrad = a.take<AlignCamGrad>(N);
  w.gcam = a.take<float>((size_t)N * NG * ALIGN_REPL);
  w.sums = a.take<float>(4);
  w.gscal = a.take<float>(4);
  return w;
}
#endif  // ST3R_HOST_EMU

}  // namespace

#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
// bit 0: segmented accumulation in the loss kernels (align_loss_seg_kernel); bit 1: one thread-block cluster per image
// in st3r_focal_weiszfeld (align_dense.cu).  0 = the variants every committed profile was measured with.
static int g_align_variant = 0;
int align_variant() { return g_align_variant; }

extern "C" {

int st3r_align_set_variant(int variant) {
  ST3R_CHECK_ARG(variant >= 0 && variant <= 7, "st3r_align_set_variant: unknown variant %d", variant);
  g_align_variant = variant;
  return ST3R_OK;
}

size_t st3r_align_ws_bytes(int n_img) {
  size_t n = (size_t)(n_img > 0 ? n_img : 1);
  return st3r_align_up(n * sizeof(AlignCam), 256) + st3r_align_up(n * sizeof(AlignCamTmp), 256) +
         st3r_align_up(n * sizeof(AlignCamGrad), 256) + st3r_align_up(n * NG * ALIGN_REPL * sizeof(float), 256) + 4096;
}

// Workspace that also lets st3r_align_optimize run its loop as ONE cooperative launch (variant bit 2): the packed
// entries of the loss terms, the per-CTA gradient rows, the schedule and the barrier word.  One carving routine serves
// the size query (null base) and the call.
struct PersistWs { float4* pk_main; float4* pk_d; float* partial; float2* sched; unsigned* bar; };
static size_t carve_persist(PersistWs* out, void* base, const St3rAlignProblem& pb, int niter) {
  WsAlloc a(base, (size_t)-1);
  const size_t n_main = (size_t)(pb.n3 > pb.n2 ? pb.n3 : pb.n2);
  const size_t row = (size_t)NG * (pb.n_img > 0 ? pb.n_img : 1) + 2;
  PersistWs w;
  w.pk_main = a.take<float4>(3 * n_main + 1);
  w.pk_d = a.take<float4>(3 * (size_t)(pb.nd > 0 ? pb.nd : 0) + 1);
  w.partial = a.take<float>(2 * row * (size_t)st3r_num_sms());
  w.sched = a.take<float2>((size_t)(niter > 0 ? niter : 1));
  w.bar = a.take<unsigned>(64);
  if (out) *out = w;
  return st3r_align_up(a.off, 256);
}
static size_t persist_extra_bytes(const St3rAlignProblem& pb, int niter) { return carve_persist(nullptr, nullptr, pb, niter); }
size_t st3r_align_ws_bytes_for(const St3rAlignProblem* prob, int niter) {
  if (!prob) return 0;
  return st3r_align_ws_bytes(prob->n_img) + persist_extra_bytes(*prob, niter);
}
int st3r_align_cam_floats(void) { return (int)(sizeof(AlignCam) / sizeof(float)); }
int st3r_align_img_const_bytes(void) { return (int)sizeof(AlignImgConst); }

int st3r_align_optimize(const St3rAlignProblem* prob, float* pp, float* log_focal, float* quat, float* trans,
                        float* log_size, float* adam_m, float* adam_v, int mode, int train_mask, float gamma,
                        float gamma_dust3r, float dust3r_w, const float* h_lr, int niter, double beta1,
                        double beta2, double eps, float* loss_hist, float* cam_out, float* pts3d_out,
                        float* depth_out, float* grad_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
  ST3R_CHECK_ARG(prob && pp && log_focal && quat && trans && log_size && adam_m && adam_v, "st3r_align_optimize: null");
  ST3R_CHECK_ARG(prob->n_img >= 1 && niter >= 0 && (mode == 0 || mode == 1), "st3r_align_optimize: bad sizes / mode");
  ST3R_CHECK_ARG(ws && ws_bytes >= st3r_align_ws_bytes(prob->n_img), "st3r_align_optimize: workspace too small");
  ST3R_CHECK_ARG(niter == 0 || h_lr, "st3r_align_optimize: missing lr schedule");
  const St3rAlignProblem pb = *prob;
  const int N = pb.n_img;
  const size_t table_bytes = (size_t)N * NG * sizeof(float);
  ST3R_CHECK_ARG(table_bytes <= 96 * 1024, "st3r_align_optimize: too many images for the shared-memory gradient table");
  static PerDeviceOnce attr;
  if (!attr.done()) {
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(align_loss3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(align_loss2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(align_lossd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(align_loss_seg_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(align_loss_seg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(align_loss_seg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr.mark();
  }
  Work w = carve(ws, N);
  Params p{pp, log_focal, quat, trans, log_size};
  AdamState ad{adam_m, adam_v};
  const int variant = g_align_variant & 1;
  const int reps = variant == 1 ? ALIGN_REPL : 1;
  const bool stage = variant == 1 && N <= CAM_STAGE_MAX;   // camera kernels with their records in shared memory
  ST3R_CHECK_CUDA(cudaMemsetAsync(w.gcam, 0, (size_t)N * NG * ALIGN_REPL * sizeof(float), stream));
  ST3R_CHECK_CUDA(cudaMemsetAsync(w.sums, 0, 4 * sizeof(float), stream));

  auto gamma_consts = [](float gm, float* off, float* offpow) {
    if (gm == 1.0f) { *off = 0.f; *offpow = 0.f; return; }
    double o = pow(1.0 / (double)gm, 1.0 / ((double)gm - 1.0));
    *off = (float)o;
    *offpow = (float)pow(o, (double)gm);
  };
  float off_m, offp_m, off_d, offp_d;
  gamma_consts(gamma, &off_m, &offp_m);
  gamma_consts(gamma_dust3r, &off_d, &offp_d);
  const int n_main = mode == 0 ? pb.n3 : pb.n2;
  const float norm_main = mode == 0 ? pb.norm3 : pb.norm2;
  const float scale_main = (n_main > 0 && norm_main != 0.f) ? 1.0f / norm_main : 0.f;
  const float scale_d = (pb.nd > 0 && pb.normd != 0.f) ? 1.0f / pb.normd : 0.f;
  auto blocks_for = [](int n) { int b = (n + LOSS_THREADS * 2 - 1) / (LOSS_THREADS * 2); return b < 1 ? 1 : (b > 1184 ? 1184 : b); };
  // variant 1: every warp walks `per_warp` consecutive entries (a multiple of 32), at most SEG_MAX_CTAS CTAs
  constexpr int WARPS = LOSS_THREADS / 32;
  auto seg_blocks = [](int n) {
    int b = (n + WARPS * SEG_MIN_PER_WARP - 1) / (WARPS * SEG_MIN_PER_WARP);
    return b < 1 ? 1 : (b > SEG_MAX_CTAS ? SEG_MAX_CTAS : b);
  };
  auto seg_per_warp = [](int n, int blocks) {
    const long long warps = (long long)blocks * WARPS;
    return (int)(((n + warps - 1) / warps + 31) / 32 * 32);
  };

  const int iters = niter > 0 ? niter : 1;
  const bool dust = pb.nd > 0 && dust3r_w != 0.f;
  const bool persist = (g_align_variant & 4) && niter > 0 && N <= PERSIST_MAX_IMG && N < 32768 &&
                       (n_main > 0 || dust) && ws_bytes >= st3r_align_ws_bytes(N) + persist_extra_bytes(pb, niter);
  if (persist) {
    // ---- one cooperative launch for the whole loop (see align_persist_kernel)
    PersistWs pw;
    carve_persist(&pw, static_cast<char*>(ws) + st3r_align_ws_bytes(N), pb, niter);
    float4* pk_main = pw.pk_main;
    float4* pk_d = pw.pk_d;
    float* partial = pw.partial;
    float2* sched = pw.sched;
    unsigned* bar = pw.bar;
    std::vector<float2> h_sched((size_t)niter);
    for (int it = 0; it < niter; ++it) {
      const double bc1 = 1.0 - pow(beta1, (double)(it + 1)), bc2 = 1.0 - pow(beta2, (double)(it + 1));
      h_sched[it] = make_float2((float)((double)h_lr[it] / bc1), (float)(1.0 / sqrt(bc2)));
    }
    // (h_sched is pageable: the call returns once the buffer has been staged, so the local vector may go away)
    ST3R_CHECK_CUDA(cudaMemcpyAsync(sched, h_sched.data(), sizeof(float2) * (size_t)niter, cudaMemcpyHostToDevice, stream));
    ST3R_CHECK_CUDA(cudaMemsetAsync(bar, 0, 64, stream));
    if (n_main > 0) {
      if (mode == 0) align_pack_kernel<0><<<(n_main + 255) / 256, 256, 0, stream>>>(pb, pk_main, n_main);
      else align_pack_kernel<1><<<(n_main + 255) / 256, 256, 0, stream>>>(pb, pk_main, n_main);
      ST3R_CHECK_LAUNCH();
    }
    if (dust) {
      align_pack_kernel<2><<<(pb.nd + 255) / 256, 256, 0, stream>>>(pb, pk_d, pb.nd);
      ST3R_CHECK_LAUNCH();
    }
    PersistArgs pa;
    pa.pb = pb; pa.p = p; pa.ad = ad; pa.w = w;
    pa.pk_main = pk_main; pa.n_main = n_main; pa.pk_d = pk_d; pa.n_d = dust ? pb.nd : 0;
    pa.gamma = gamma; pa.off_m = off_m; pa.offp_m = offp_m; pa.scale_main = scale_main;
    pa.gamma_d = gamma_dust3r; pa.off_d = off_d; pa.offp_d = offp_d; pa.scale_d = scale_d * dust3r_w;
    pa.train_mask = train_mask; pa.omb1 = (float)(1.0 - beta1); pa.b2 = (float)beta2; pa.omb2 = (float)(1.0 - beta2);
    pa.eps = (float)eps; pa.sched = sched; pa.niter = niter; pa.loss_hist = loss_hist; pa.grad_out = grad_out;
    pa.partial = partial; pa.bar = bar;
    // enough CTAs that a warp walks ~4 rows of 32 entries, at most one CTA per SM (all resident: the barrier needs it)
    const long long rows = ((long long)n_main + pa.n_d + 31) / 32;
#ifdef ST3R_HOST_EMU
    long long want = (rows + 7) / 8;                       // (test builds: several CTAs even on the small fixtures)
#else
    long long want = (rows + 4 * (PERSIST_THREADS / 32) - 1) / (4 * (PERSIST_THREADS / 32));
#endif
    int grid = (int)(want < 1 ? 1 : (want > st3r_num_sms() ? st3r_num_sms() : want));
#ifdef ST3R_HOST_EMU
    if (grid > 4) grid = 4;                                // the emulator runs the grid as one cluster of fibers
    if (mode == 0) emu_launch_cluster(grid, dim3(grid), dim3(PERSIST_THREADS), [&]() { align_persist_kernel<0>(pa); });
    else emu_launch_cluster(grid, dim3(grid), dim3(PERSIST_THREADS), [&]() { align_persist_kernel<1>(pa); });
#else
    void* kargs[] = {(void*)&pa};
    const void* fn = mode == 0 ? (const void*)align_persist_kernel<0> : (const void*)align_persist_kernel<1>;
    int per_sm = 0;
    ST3R_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PERSIST_THREADS, 0));
    ST3R_CHECK_ARG(per_sm >= 1, "st3r_align_optimize: the persistent kernel does not fit on an SM");
    ST3R_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(PERSIST_THREADS), kargs, 0, stream));
#endif
    ST3R_CHECK_LAUNCH();
  } else
  for (int it = 0; it < iters; ++it) {
    if (stage) align_cam_fwd_kernel<true><<<1, CAM_THREADS, 0, stream>>>(pb, p, w);
    else align_cam_fwd_kernel<false><<<1, CAM_THREADS, 0, stream>>>(pb, p, w);
    ST3R_CHECK_LAUNCH();
    if (niter == 0) break;
    if (n_main > 0) {
      if (variant == 1) {
        const int gb = seg_blocks(n_main), pw = seg_per_warp(n_main, gb);
        if (mode == 0)
          align_loss_seg_kernel<0><<<gb, LOSS_THREADS, table_bytes, stream>>>(pb, w, gamma, off_m, offp_m, scale_main, pw, w.sums + 0);
        else
          align_loss_seg_kernel<1><<<gb, LOSS_THREADS, table_bytes, stream>>>(pb, w, gamma, off_m, offp_m, scale_main, pw, w.sums + 0);
      } else if (mode == 0) {
        align_loss3d_kernel<<<blocks_for(n_main), LOSS_THREADS, table_bytes, stream>>>(pb, w, gamma, off_m, offp_m, scale_main);
      } else {
        align_loss2d_kernel<<<blocks_for(n_main), LOSS_THREADS, table_bytes, stream>>>(pb, w, gamma, off_m, offp_m, scale_main);
      }
      ST3R_CHECK_LAUNCH();
    }
    if (pb.nd > 0 && dust3r_w != 0.f) {
      // the dust3r term enters the total loss with weight dust3r_w (reconstruct.py:389)
      if (variant == 1) {
        const int gb = seg_blocks(pb.nd), pw = seg_per_warp(pb.nd, gb);
        align_loss_seg_kernel<2><<<gb, LOSS_THREADS, table_bytes, stream>>>(pb, w, gamma_dust3r, off_d, offp_d,
                                                                           scale_d * dust3r_w, pw, w.sums + 1);
      } else {
        align_lossd_kernel<<<blocks_for(pb.nd), LOSS_THREADS, table_bytes, stream>>>(pb, w, gamma_dust3r, off_d, offp_d,
                                                                                     scale_d * dust3r_w);
      }
      ST3R_CHECK_LAUNCH();
    }
    const int step = it + 1;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    if (stage)
      align_cam_bwd_kernel<true><<<1, CAM_THREADS, 0, stream>>>(pb, p, ad, w, train_mask, (float)((double)h_lr[it] / bc1),
                                                                (float)(1.0 / sqrt(bc2)), (float)(1.0 - beta1), (float)beta2,
                                                                (float)(1.0 - beta2), (float)eps, 1.0f, loss_hist, it,
                                                                (it == iters - 1) ? grad_out : nullptr, reps);
    else
      align_cam_bwd_kernel<false><<<1, CAM_THREADS, 0, stream>>>(pb, p, ad, w, train_mask, (float)((double)h_lr[it] / bc1),
                                                                 (float)(1.0 / sqrt(bc2)), (float)(1.0 - beta1), (float)beta2,
                                                                 (float)(1.0 - beta2), (float)eps, 1.0f, loss_hist, it,
                                                                 (it == iters - 1) ? grad_out : nullptr, reps);
    ST3R_CHECK_LAUNCH();
  }
  if (cam_out) ST3R_CHECK_CUDA(cudaMemcpyAsync(cam_out, w.cam, (size_t)N * sizeof(AlignCam), cudaMemcpyDeviceToDevice, stream));
  if (pts3d_out || depth_out) {
    ST3R_CHECK_ARG(pts3d_out && depth_out, "st3r_align_optimize: pts3d_out and depth_out go together");
    int total = pb.n_anchor > pb.n_core_total ? pb.n_anchor : pb.n_core_total;
    if (total > 0) {
      align_outputs_kernel<<<(total + 255) / 256, 256, 0, stream>>>(pb, w, pts3d_out, depth_out, pb.n_core_total);
      ST3R_CHECK_LAUNCH();
    }
  }
  return ST3R_OK;
}

}  // extern "C"
#endif  // ST3R_HOST_EMU
