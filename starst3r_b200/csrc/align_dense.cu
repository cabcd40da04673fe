// Dense-geometry kernels either side of the ALIGN optimiser (all HBM-bound streaming passes):
//   canonical_view  (mast3r/cloud_opt/sparse_ga.py:817-855, mode 'avg-angle'): confidence-weighted mean point map of an
//                   image over its pair entries + per 8x8 block angle-averaged relative depth; 16 B read per
//                   (pair entry, pixel), 20 B written per pixel; the pair entries are streamed, never stacked.
//   focal_weiszfeld (dust3r/post_process.py:36-58): closed-form init + 10 IRLS re-weightings, one CTA per image.
//   dense_points    (SparseGA.get_dense_pts3d, sparse_ga.py:70-93 + make_pts3d :475-501 with every pixel an anchor).
//   clean_pointcloud(dust3r/cloud_opt/base_opt.py:369-405): lower the confidence of points that lie in front of a
//                   more confident view's depth map; images are processed in order because image i reads the
//                   already cleaned confidences of images j < i.
// ST3R_HOST_EMU: test builds that run this file on a CPU SIMT emulator (tests/host/): dense_emu_host.cpp includes the
// kernels only, build_emu_lib.py (ST3R_EMU_WHOLE) compiles the entry points too, with their launches rewritten.
#ifndef ST3R_HOST_EMU
#include <cooperative_groups.h>
#endif
#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
#include "common.cuh"
#include "../../include/starst3r_b200.h"

int align_variant();   // align.cu (st3r_align_set_variant)
#endif

namespace cg = cooperative_groups;

namespace {

// ptmaps [P,H,W,3], confs [P,H,W] -> canon [H,W,3], canon2 [H,W], cconf [H,W].  One thread per pixel.
__global__ void __launch_bounds__(256)
canonical_view_kernel(const float* __restrict__ ptmaps, const float* __restrict__ confs, int P, int H, int W, int S,
                      float* __restrict__ canon, float* __restrict__ canon2, float* __restrict__ cconf) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)y * W + x;
  const size_t cpix = (size_t)((y / S) * S + S / 2) * W + ((x / S) * S + S / 2);   // block centre
  float sw = 0.f, sw2 = 0.f, sx = 0.f, sy = 0.f, sz = 0.f, swa = 0.f, srad = 0.f, scz = 0.f, swc = 0.f;
  for (int p = 0; p < P; ++p) {
    const float* X = ptmaps + ((size_t)p * HW + pix) * 3;
    const float* Xc = ptmaps + ((size_t)p * HW + cpix) * 3;
    const float w = confs[(size_t)p * HW + pix] - 0.999f;
    const float wc = confs[(size_t)p * HW + cpix] - 0.999f;
    const float px = X[0], py = X[1], pz = X[2];
    sw += w; sw2 += w * w;
    sx += w * px; sy += w * py; sz += w * pz;
    const float cz = fmaxf(Xc[2], 1.1920929e-07f);
    const float dx = px - Xc[0], dy = py - Xc[1];
    const float rad = fmaxf(sqrtf(dx * dx + dy * dy), 1e-8f);
    swa += w * atanf((pz - cz) / rad);
    srad += rad;
    scz += wc * Xc[2];   // canon z at the block centre = weighted mean with the centre pixel's weights
    swc += wc;
  }
  canon[pix * 3] = sx / sw; canon[pix * 3 + 1] = sy / sw; canon[pix * 3 + 2] = sz / sw;
  cconf[pix] = sw2 / sw;
  const float depth = (srad / (float)P) * tanf(swa / sw);
  canon2[pix] = 1.0f + depth / (scz / swc);
}

// One CTA per image: focal = argmin sum |pixel - focal * xy/z| by Weiszfeld iterations.
__global__ void __launch_bounds__(1024)
focal_weiszfeld_kernel(const float* __restrict__ canon, int H, int W, float min_focal, float max_focal,
                       float* __restrict__ out) {
  __shared__ float red[2][32];
  __shared__ float s_f;
  const float* X = canon + (size_t)blockIdx.x * H * W * 3;
  const int n = H * W;
  const float cx = 0.5f * W, cy = 0.5f * H;
  float focal = 0.f;
  for (int iter = 0; iter <= 10; ++iter) {
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float u = (float)(i % W) - cx, v = (float)(i / W) - cy;
      float qx = X[3 * i] / X[3 * i + 2], qy = X[3 * i + 1] / X[3 * i + 2];
      if (isinf(qx)) qx = 0.f;
      if (isinf(qy)) qy = 0.f;
      if (isnan(qx)) qx = 0.f;
      if (isnan(qy)) qy = 0.f;
      const float dxp = qx * u + qy * v, dxx = qx * qx + qy * qy;
      float w = 1.f;
      if (iter > 0) {
        const float ex = u - focal * qx, ey = v - focal * qy;
        w = 1.0f / fmaxf(sqrtf(ex * ex + ey * ey), 1e-8f);
      }
      a += w * dxp;
      b += w * dxx;
    }
    for (int off = 16; off; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if (lane_id() == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float sa = 0.f, sb = 0.f;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { sa += red[0][k]; sb += red[1][k]; }
      s_f = sa / sb;   // the means' 1/n cancels
    }
    __syncthreads();
    focal = s_f;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float fb = (float)max(H, W) / (2.0f * tanf(0.5235987755982988f));
    out[blockIdx.x] = fminf(fmaxf(focal, min_focal * fb), max_focal * fb);
  }
}

// Variant with a thread-block cluster per image (st3r_align_set_variant bit 1).  The single-CTA kernel above streams
// the point map eleven times through one SM (1.5 ms per 512 x 512 image, profiles/r01u_launches_reconstruct.csv);
// here WZ_CLUSTER CTAs split the pixels, publish their partial sums in their own shared memory, and after a cluster
// barrier every CTA adds the partials of all ranks in rank order over distributed shared memory, so all of them
// continue with the same focal.  Two slots alternate between iterations: one cluster barrier per iteration.
constexpr int WZ_CLUSTER = 8;

__global__ void __cluster_dims__(WZ_CLUSTER, 1, 1) __launch_bounds__(1024)
focal_weiszfeld_cluster_kernel(const float* __restrict__ canon, int H, int W, float min_focal, float max_focal,
                               float* __restrict__ out) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float red[2][32];
  __shared__ float part[2][2];
  __shared__ float s_f;
  const int img = blockIdx.x / WZ_CLUSTER;
  const int rank = (int)cluster.block_rank();
  const float* X = canon + (size_t)img * H * W * 3;
  const int n = H * W;
  const float cx = 0.5f * W, cy = 0.5f * H;
  float focal = 0.f;
  for (int iter = 0; iter <= 10; ++iter) {
    float a = 0.f, b = 0.f;
    for (int i = rank * blockDim.x + threadIdx.x; i < n; i += WZ_CLUSTER * blockDim.x) {
      const float u = (float)(i % W) - cx, v = (float)(i / W) - cy;
      float qx = X[3 * i] / X[3 * i + 2], qy = X[3 * i + 1] / X[3 * i + 2];
      if (isinf(qx)) qx = 0.f;
      if (isinf(qy)) qy = 0.f;
      if (isnan(qx)) qx = 0.f;
      if (isnan(qy)) qy = 0.f;
      const float dxp = qx * u + qy * v, dxx = qx * qx + qy * qy;
      float w = 1.f;
      if (iter > 0) {
        const float ex = u - focal * qx, ey = v - focal * qy;
        w = 1.0f / fmaxf(sqrtf(ex * ex + ey * ey), 1e-8f);
      }
      a += w * dxp;
      b += w * dxx;
    }
    for (int off = 16; off; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      b += __shfl_xor_sync(0xffffffffu, b, off);
    }
    if (lane_id() == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float sa = 0.f, sb = 0.f;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { sa += red[0][k]; sb += red[1][k]; }
      part[iter & 1][0] = sa;
      part[iter & 1][1] = sb;
    }
    cluster.sync();                       // every rank's partial of this iteration is published
    if (threadIdx.x == 0) {
      float sa = 0.f, sb = 0.f;
      for (int r = 0; r < WZ_CLUSTER; ++r) {
        const float* rp = cluster.map_shared_rank(&part[iter & 1][0], r);
        sa += rp[0];
        sb += rp[1];
      }
      s_f = sa / sb;
    }
    __syncthreads();
    focal = s_f;
  }
  cluster.sync();                         // nobody exits while a peer may still read its shared memory
  if (rank == 0 && threadIdx.x == 0) {
    const float fb = (float)max(H, W) / (2.0f * tanf(0.5235987755982988f));
    out[img] = fminf(fmaxf(focal, min_focal * fb), max_focal * fb);
  }
}

struct DenseCam { float R[9]; float t[3]; float f, cx, cy, bf; };

// Every pixel as an anchor of its 8x8 block: pts3d [HW,3] (world) and depth [HW] (camera z).
__global__ void __launch_bounds__(256)
dense_points_kernel(const float* __restrict__ canon2, const float* __restrict__ core_depth, DenseCam cam, int H, int W,
                    int S, float* __restrict__ pts3d, float* __restrict__ depth_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int y = i / W, x = i - y * W;
  const int W2 = (W - S / 2 + S - 1) / S;
  const int k = (y / S) * W2 + (x / S);
  const int cyy = (y / S) * S + S / 2, cxx = (x / S) * S + S / 2;
  const float off = canon2[i] / canon2[(size_t)cyy * W + cxx];
  const float op = 1.0f + (off - 1.0f) * (cam.bf / cam.f);
  const float z = core_depth[k] * op;
  const float pc[3] = {z * (((float)x - cam.cx) / cam.f), z * (((float)y - cam.cy) / cam.f), z};
  depth_out[i] = z;
  for (int a = 0; a < 3; ++a)
    pts3d[3 * (size_t)i + a] = cam.R[3 * a] * pc[0] + cam.R[3 * a + 1] * pc[1] + cam.R[3 * a + 2] * pc[2] + cam.t[a];
}

struct CleanCam { float Rw[9]; float tw[3]; float K[9]; };   // world->camera and intrinsics of view j

__global__ void __launch_bounds__(256)
clean_pointcloud_kernel(const float* __restrict__ pts_i, float* __restrict__ conf_all, const float* __restrict__ depth_all,
                        const CleanCam* __restrict__ cams, int N, int i, int H, int W, float tol, float bad_conf) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (p >= HW) return;
  const float X[3] = {pts_i[3 * (size_t)p], pts_i[3 * (size_t)p + 1], pts_i[3 * (size_t)p + 2]};
  float c = conf_all[(size_t)i * HW + p];
  for (int j = 0; j < N; ++j) {
    if (j == i) continue;
    const CleanCam cj = cams[j];
    float q[3];
    for (int a = 0; a < 3; ++a) q[a] = cj.Rw[3 * a] * X[0] + cj.Rw[3 * a + 1] * X[1] + cj.Rw[3 * a + 2] * X[2] + cj.tw[a];
    const float uh = cj.K[0] * q[0] + cj.K[1] * q[1] + cj.K[2] * q[2];
    const float vh = cj.K[3] * q[0] + cj.K[4] * q[1] + cj.K[5] * q[2];
    const float wh = cj.K[6] * q[0] + cj.K[7] * q[1] + cj.K[8] * q[2];
    const float uf = rintf(uh / wh), vf = rintf(vh / wh);
    if (!(q[2] > 0.f) || !(uf >= 0.f) || !(uf < (float)W) || !(vf >= 0.f) || !(vf < (float)H)) continue;
    const int u = (int)uf, v = (int)vf;
    const size_t t = (size_t)j * HW + (size_t)v * W + u;
    if (q[2] < (1.0f - tol) * depth_all[t] && c < conf_all[t]) c = fminf(c, bad_conf);
  }
  conf_all[(size_t)i * HW + p] = c;
}

}  // namespace

#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
extern "C" {

int st3r_canonical_view(const float* ptmaps, const float* confs, int n_entries, int H, int W, int subsample,
                        float* canon, float* canon2, float* cconf, cudaStream_t stream) {
  ST3R_CHECK_ARG(n_entries > 0 && H > 0 && W > 0 && subsample > 0, "st3r_canonical_view: not a single view-1 point map");
  ST3R_CHECK_ARG(H % subsample == 0 && W % subsample == 0, "st3r_canonical_view: H, W must be multiples of subsample");
  ST3R_CHECK_ARG(ptmaps && confs && canon && canon2 && cconf, "st3r_canonical_view: null pointer");
  dim3 grid((W + 255) / 256, H);
  canonical_view_kernel<<<grid, 256, 0, stream>>>(ptmaps, confs, n_entries, H, W, subsample, canon, canon2, cconf);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_focal_weiszfeld(const float* canon, int n_img, int H, int W, float min_focal, float max_focal, float* focal_out,
                         cudaStream_t stream) {
  ST3R_CHECK_ARG(n_img >= 0 && H > 0 && W > 0, "st3r_focal_weiszfeld: bad sizes");
  if (n_img == 0) return ST3R_OK;
  ST3R_CHECK_ARG(canon && focal_out, "st3r_focal_weiszfeld: null pointer");
  if (align_variant() & 2)
    focal_weiszfeld_cluster_kernel<<<n_img * WZ_CLUSTER, 1024, 0, stream>>>(canon, H, W, min_focal, max_focal, focal_out);
  else
    focal_weiszfeld_kernel<<<n_img, 1024, 0, stream>>>(canon, H, W, min_focal, max_focal, focal_out);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_dense_points(const float* canon2, const float* core_depth, const float* h_cam2w, const float* h_K,
                      float base_focal, int H, int W, int subsample, float* pts3d, float* depth, cudaStream_t stream) {
  ST3R_CHECK_ARG(canon2 && core_depth && h_cam2w && h_K && pts3d && depth, "st3r_dense_points: null pointer");
  ST3R_CHECK_ARG(H > 0 && W > 0 && subsample > 0, "st3r_dense_points: bad sizes");
  DenseCam c;
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < 3; ++b) c.R[3 * a + b] = h_cam2w[4 * a + b];
    c.t[a] = h_cam2w[4 * a + 3];
  }
  c.f = h_K[0]; c.cx = h_K[2]; c.cy = h_K[5]; c.bf = base_focal;
  dense_points_kernel<<<(H * W + 255) / 256, 256, 0, stream>>>(canon2, core_depth, c, H, W, subsample, pts3d, depth);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_clean_cam_floats(void) { return (int)(sizeof(CleanCam) / sizeof(float)); }

int st3r_clean_pointcloud(const float* pts3d, float* confs, const float* depthmaps, const float* cams, int n_img, int H,
                          int W, float tol, float bad_conf, cudaStream_t stream) {
  ST3R_CHECK_ARG(n_img >= 0 && H > 0 && W > 0 && tol >= 0.f && tol < 1.f, "st3r_clean_pointcloud: bad args");
  if (n_img == 0) return ST3R_OK;
  ST3R_CHECK_ARG(pts3d && confs && depthmaps && cams, "st3r_clean_pointcloud: null pointer");
  const int HW = H * W;
  for (int i = 0; i < n_img; ++i) {
    clean_pointcloud_kernel<<<(HW + 255) / 256, 256, 0, stream>>>(pts3d + (size_t)i * HW * 3, confs, depthmaps,
                                                                 reinterpret_cast<const CleanCam*>(cams), n_img, i, H, W,
                                                                 tol, bad_conf);
    ST3R_CHECK_LAUNCH();
  }
  return ST3R_OK;
}
}
#endif  // ST3R_HOST_EMU
