// Fused photometric loss of starster/gs.py:126-131, forward and backward:
//   loss_view = (1 - f) * L1(truth, render) + f * (1 - SSIM(truth, render))
// SSIM restates torchmetrics StructuralSimilarityIndexMeasure(data_range=1):
// 11x11 Gaussian window (sigma 1.5), reflect-pad 5 then crop 5  ==  mean over the interior pixels
// [5, H-5) x [5, W-5) of the valid-window SSIM map, so the padding never contributes.
// Forward kernel: separable window statistics in shared memory -> SSIM sum, L1 sum and the three
// derivative maps (dS/dmu_x, dS/dsigma_x^2, dS/dsigma_xy, already scaled by dLoss/dS).
// Backward kernel: the same separable window over the derivative maps + the L1 sign term ->
// dLoss/d(render), which feeds the blend backward directly.  HBM-bound: reads 12 B render +
// 12 B truth, writes 12 B gradient per pixel (+ 36 B of derivative maps written and re-read once).
#include "common.cuh"
#include "gs.cuh"

namespace {

constexpr int LT = 16;         // output tile edge
constexpr int HALO = 5;
constexpr int LW = LT + 2 * HALO;  // 26
constexpr float C1 = 0.01f * 0.01f;
constexpr float C2 = 0.03f * 0.03f;

__constant__ float c_win[11];

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  }
  return s;  // valid in thread 0
}

// grid: (tiles_x, tiles_y, C); block: 256 threads.  Images are [C, H, W, 3] interleaved.
__global__ void __launch_bounds__(LT * LT)
ssim_l1_fwd_kernel(const float* __restrict__ render, const float* __restrict__ truth, int H, int W,
                   float coef_ssim /* dLoss/dS per interior sample */, float* __restrict__ dmaps /* [C,H,W,3,3] */,
                   float* __restrict__ sums /* per view: [ssim_sum, l1_sum] */) {
  __shared__ float tx[LW][LW + 1], ty[LW][LW + 1];
  __shared__ float hs[5][LW][LT + 1];
  __shared__ float red[8];
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  const int px = x0 + lx, py = y0 + ly;
  const bool in_img = px < W && py < H;
  const bool interior = px >= HALO && px < W - HALO && py >= HALO && py < H - HALO;
  const size_t img = (size_t)c * H * W;
  float ssim_acc = 0.f, l1_acc = 0.f;
  for (int ch = 0; ch < 3; ++ch) {
    __syncthreads();
    for (int e = threadIdx.x; e < LW * LW; e += LT * LT) {
      int r = e / LW, q = e - r * LW;
      int gy = y0 + r - HALO, gx = x0 + q - HALO;
      float a = 0.f, b = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        size_t p = (img + (size_t)gy * W + gx) * 3 + ch;
        a = render[p];
        b = truth[p];
      }
      tx[r][q] = a;
      ty[r][q] = b;
    }
    __syncthreads();
    // horizontal pass: rows 0..25, output columns 0..15
    for (int e = threadIdx.x; e < LW * LT; e += LT * LT) {
      int r = e / LT, q = e - r * LT;
      float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        float w = c_win[k], a = tx[r][q + k], b = ty[r][q + k];
        sx += w * a; sy += w * b; sxx += w * a * a; syy += w * b * b; sxy += w * a * b;
      }
      hs[0][r][q] = sx; hs[1][r][q] = sy; hs[2][r][q] = sxx; hs[3][r][q] = syy; hs[4][r][q] = sxy;
    }
    __syncthreads();
    float dm = 0.f, ds = 0.f, dc = 0.f;
    if (interior) {
      float mx = 0.f, my = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        float w = c_win[k];
        mx += w * hs[0][ly + k][lx]; my += w * hs[1][ly + k][lx]; exx += w * hs[2][ly + k][lx];
        eyy += w * hs[3][ly + k][lx]; exy += w * hs[4][ly + k][lx];
      }
      float sxx = exx - mx * mx, syy = eyy - my * my, sxy = exy - mx * my;
      float A1 = 2.f * mx * my + C1, A2 = 2.f * sxy + C2, B1 = mx * mx + my * my + C1, B2 = sxx + syy + C2;
      float inv = 1.f / (B1 * B2);
      float S = A1 * A2 * inv;
      ssim_acc += S;
      float S_mu = 2.f * my * A2 * inv - 2.f * mx * S / B1;
      float S_sx = -S / B2;
      float S_c = 2.f * A1 * inv;
      dm = coef_ssim * (S_mu - 2.f * mx * S_sx - my * S_c);
      ds = coef_ssim * S_sx;
      dc = coef_ssim * S_c;
    }
    if (in_img) {
      size_t p = ((img + (size_t)py * W + px) * 3 + ch) * 3;
      dmaps[p] = dm; dmaps[p + 1] = ds; dmaps[p + 2] = dc;
      l1_acc += fabsf(ty[ly + HALO][lx + HALO] - tx[ly + HALO][lx + HALO]);
    }
  }
  float s1 = block_sum(ssim_acc, red);
  float s2 = block_sum(l1_acc, red);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 2 * c, s1);
    atomicAdd(sums + 2 * c + 1, s2);
  }
}

__global__ void __launch_bounds__(LT * LT)
ssim_l1_bwd_kernel(const float* __restrict__ render, const float* __restrict__ truth, const float* __restrict__ dmaps,
                   int H, int W, float coef_l1 /* (1-f) / (3HW) */, float* __restrict__ v_render) {
  __shared__ float tm[3][LW][LW + 1];
  __shared__ float hs[3][LW][LT + 1];
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  const int px = x0 + lx, py = y0 + ly;
  const bool in_img = px < W && py < H;
  const size_t img = (size_t)c * H * W;
  for (int ch = 0; ch < 3; ++ch) {
    __syncthreads();
    for (int e = threadIdx.x; e < LW * LW; e += LT * LT) {
      int r = e / LW, q = e - r * LW;
      int gy = y0 + r - HALO, gx = x0 + q - HALO;
      float a = 0.f, b = 0.f, d = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        size_t p = ((img + (size_t)gy * W + gx) * 3 + ch) * 3;
        a = dmaps[p]; b = dmaps[p + 1]; d = dmaps[p + 2];
      }
      tm[0][r][q] = a; tm[1][r][q] = b; tm[2][r][q] = d;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < LW * LT; e += LT * LT) {
      int r = e / LT, q = e - r * LT;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        float w = c_win[k];
        s0 += w * tm[0][r][q + k]; s1 += w * tm[1][r][q + k]; s2 += w * tm[2][r][q + k];
      }
      hs[0][r][q] = s0; hs[1][r][q] = s1; hs[2][r][q] = s2;
    }
    __syncthreads();
    if (in_img) {
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        float w = c_win[k];
        g0 += w * hs[0][ly + k][lx]; g1 += w * hs[1][ly + k][lx]; g2 += w * hs[2][ly + k][lx];
      }
      size_t p = (img + (size_t)py * W + px) * 3 + ch;
      float x = render[p], y = truth[p];
      float d = y - x;
      float gl1 = d > 0.f ? -coef_l1 : (d < 0.f ? coef_l1 : 0.f);
      v_render[p] = g0 + 2.f * x * g1 + y * g2 + gl1;
    }
  }
}

// loss = sum_views [(1-f) * L1_v + f * (1 - SSIM_v)] + reg_o * sum sigmoid(opacity) + reg_s * sum exp(scale)
// (gs.py:126-136 summed over the views, gs.py:149-152), evaluated in the order the PyTorch expression uses.
__global__ void loss_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ reg, int C, float inv_l1,
                                     float inv_ssim, float f, float reg_o, float reg_s, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const float l1 = sums[2 * c + 1] * inv_l1, ssim = sums[2 * c] * inv_ssim;
    acc += l1 * (1.0f - f) + (1.0f - ssim) * f;
  }
  if (reg) acc = acc + reg[0] * reg_o + reg[1] * reg_s;
  *out = acc;
}

bool g_win_set = false;

int set_window() {
  if (g_win_set) return ST3R_OK;
  float g[11], s = 0.f;
  for (int k = 0; k < 11; ++k) {
    float d = (float)(k - 5);
    g[k] = expf(-(d / 1.5f) * (d / 1.5f) / 2.0f);
    s += g[k];
  }
  for (int k = 0; k < 11; ++k) g[k] /= s;
  ST3R_CHECK_CUDA(cudaMemcpyToSymbol(c_win, g, sizeof(g)));
  g_win_set = true;
  return ST3R_OK;
}

}  // namespace

extern "C" {

// sums [C,2] must be zeroed by the caller; dmaps [C,H,W,3,3] scratch.
int st3r_gs_loss_fwd(const float* render, const float* truth, int C, int height, int width, float ssim_fac,
                     float* dmaps, float* sums, cudaStream_t stream) {
  ST3R_CHECK_ARG(C >= 0 && height > 10 && width > 10, "st3r_gs_loss_fwd: images must be larger than the 11x11 SSIM window");
  if (C == 0) return ST3R_OK;
  ST3R_CHECK_ARG(render && truth && dmaps && sums, "st3r_gs_loss_fwd: null pointer");
  int rc = set_window();
  if (rc) return rc;
  // loss_view = ... + f * (1 - mean(S))  =>  dLoss/dS = -f / (3 (H-10) (W-10))
  float coef = -ssim_fac / (3.0f * (float)(height - 10) * (float)(width - 10));
  dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, C);
  ssim_l1_fwd_kernel<<<grid, LT * LT, 0, stream>>>(render, truth, height, width, coef, dmaps, sums);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_gs_loss_bwd(const float* render, const float* truth, const float* dmaps, int C, int height, int width,
                     float ssim_fac, float* v_render, cudaStream_t stream) {
  ST3R_CHECK_ARG(C >= 0 && height > 10 && width > 10, "st3r_gs_loss_bwd: bad sizes");
  if (C == 0) return ST3R_OK;
  ST3R_CHECK_ARG(render && truth && dmaps && v_render, "st3r_gs_loss_bwd: null pointer");
  int rc = set_window();
  if (rc) return rc;
  float coef_l1 = (1.0f - ssim_fac) / (3.0f * (float)height * (float)width);
  dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, C);
  ssim_l1_bwd_kernel<<<grid, LT * LT, 0, stream>>>(render, truth, dmaps, height, width, coef_l1, v_render);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_gs_loss_finalize(const float* sums, const float* reg_sums, int C, int height, int width, float ssim_fac,
                          float reg_opac, float reg_scale, float* loss_out, cudaStream_t stream) {
  ST3R_CHECK_ARG(C >= 0 && height > 10 && width > 10, "st3r_gs_loss_finalize: bad sizes");
  ST3R_CHECK_ARG(sums && loss_out, "st3r_gs_loss_finalize: null pointer");
  loss_finalize_kernel<<<1, 32, 0, stream>>>(sums, reg_sums, C, 1.0f / (3.0f * (float)height * (float)width),
                                             1.0f / (3.0f * (float)(height - 10) * (float)(width - 10)), ssim_fac,
                                             reg_opac, reg_scale, loss_out);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
}
