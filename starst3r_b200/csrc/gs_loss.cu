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

// Tiling: one CTA of 256 threads produces a 32 x 32 pixel tile of one view.  The 42 x 42 halo tile of all three
// channels is loaded once, coalesced, from the interleaved [H,W,3] images into planar shared arrays; the separable
// 11-tap window then runs per channel with register tiling: a thread of the horizontal pass produces 8 consecutive
// columns of one row from 18 loaded inputs, a thread of the vertical pass 4 consecutive rows of one column from 14
// (the first version loaded 11 inputs per output in both passes and three strided passes over global memory).
constexpr int LT = 32;             // output tile edge
constexpr int HALO = 5;
constexpr int LW = LT + 2 * HALO;  // 42
constexpr int LWP = LW + 1;        // padded row pitch of the input tiles
constexpr int HSP = LT + 1;        // padded row pitch of the horizontally filtered rows
constexpr int NT = 256;
constexpr float C1 = 0.01f * 0.01f;
constexpr float C2 = 0.03f * 0.03f;

#ifndef ST3R_HOST_EMU            // (CPU emulator builds, tests/host/, define it themselves)
#define ST3R_DYN_SMEM_F32(name) extern __shared__ float name[]
#endif
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

// Loads the halo tile of an interleaved [H,W,3] image into three planar [LW][LWP] arrays (zero outside the image).
__device__ __forceinline__ void load_rgb_tile(const float* __restrict__ img, int H, int W, int x0, int y0, float* dst) {
  for (int e = threadIdx.x; e < LW * LW * 3; e += NT) {
    const int r = e / (LW * 3), q3 = e - r * (LW * 3);
    const int q = q3 / 3, ch = q3 - q * 3;
    const int gy = y0 + r - HALO, gx = x0 + q - HALO;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = img[((size_t)gy * W + gx) * 3 + ch];
    dst[(ch * LW + r) * LWP + q] = v;
  }
}

// grid: (tiles_x, tiles_y, C); block: 256 threads.  dmaps is planar scratch: [C][channel 3][derivative 3][H][W].
__global__ void __launch_bounds__(NT)
ssim_l1_fwd_kernel(const float* __restrict__ render, const float* __restrict__ truth, int H, int W,
                   float coef_ssim /* dLoss/dS per interior sample */, float* __restrict__ dmaps,
                   float* __restrict__ sums /* per view: [ssim_sum, l1_sum] */) {
  ST3R_DYN_SMEM_F32(smem);
  float* tx = smem;                         // [3][LW][LWP]
  float* ty = tx + 3 * LW * LWP;            // [3][LW][LWP]
  float* hs = ty + 3 * LW * LWP;            // [5][LW][HSP]
  __shared__ float red[8];
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t plane = (size_t)H * W;
  load_rgb_tile(render + (size_t)c * plane * 3, H, W, x0, y0, tx);
  load_rgb_tile(truth + (size_t)c * plane * 3, H, W, x0, y0, ty);
  float w[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) w[k] = c_win[k];
  const int vx = threadIdx.x & 31, vr0 = (threadIdx.x >> 5) * 4;     // vertical pass: column vx, rows vr0 .. vr0 + 3
  float ssim_acc = 0.f, l1_acc = 0.f;
  for (int ch = 0; ch < 3; ++ch) {
    __syncthreads();
    // horizontal pass: item = (row r, group of 8 output columns)
    for (int it = threadIdx.x; it < LW * (LT / 8); it += NT) {
      const int r = it >> 2, q0 = (it & 3) * 8;
      const float* ax = tx + (ch * LW + r) * LWP + q0;
      const float* ay = ty + (ch * LW + r) * LWP + q0;
      float a[18], b[18];
#pragma unroll
      for (int k = 0; k < 18; ++k) { a[k] = ax[k]; b[k] = ay[k]; }
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const float wa = w[k] * a[o + k], wb = w[k] * b[o + k];
          sx += wa; sy += wb; sxx += wa * a[o + k]; syy += wb * b[o + k]; sxy += wa * b[o + k];
        }
        const int idx = r * HSP + q0 + o;
        hs[idx] = sx; hs[LW * HSP + idx] = sy; hs[2 * LW * HSP + idx] = sxx; hs[3 * LW * HSP + idx] = syy;
        hs[4 * LW * HSP + idx] = sxy;
      }
    }
    __syncthreads();
    // vertical pass: 4 output rows of one column from 14 filtered rows
    float col[5][14];
#pragma unroll
    for (int st = 0; st < 5; ++st)
#pragma unroll
      for (int k = 0; k < 14; ++k) col[st][k] = hs[st * LW * HSP + (vr0 + k) * HSP + vx];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int px = x0 + vx, py = y0 + vr0 + o;
      const bool in_img = px < W && py < H;
      const bool interior = px >= HALO && px < W - HALO && py >= HALO && py < H - HALO;
      float dm = 0.f, ds = 0.f, dc = 0.f;
      if (interior) {
        float mx = 0.f, my = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          mx += w[k] * col[0][o + k]; my += w[k] * col[1][o + k]; exx += w[k] * col[2][o + k];
          eyy += w[k] * col[3][o + k]; exy += w[k] * col[4][o + k];
        }
        const float sxx = exx - mx * mx, syy = eyy - my * my, sxy = exy - mx * my;
        const float A1 = 2.f * mx * my + C1, A2 = 2.f * sxy + C2, B1 = mx * mx + my * my + C1, B2 = sxx + syy + C2;
        const float inv = 1.f / (B1 * B2);
        const float S = A1 * A2 * inv;
        ssim_acc += S;
        const float S_mu = 2.f * my * A2 * inv - 2.f * mx * S / B1;
        const float S_sx = -S / B2;
        const float S_c = 2.f * A1 * inv;
        dm = coef_ssim * (S_mu - 2.f * mx * S_sx - my * S_c);
        ds = coef_ssim * S_sx;
        dc = coef_ssim * S_c;
      }
      if (in_img) {
        float* d = dmaps + (((size_t)c * 3 + ch) * 3) * plane + (size_t)py * W + px;
        d[0] = dm; d[plane] = ds; d[2 * plane] = dc;
        const int t = (ch * LW + vr0 + o + HALO) * LWP + vx + HALO;
        l1_acc += fabsf(ty[t] - tx[t]);
      }
    }
  }
  const float s1 = block_sum(ssim_acc, red);
  const float s2 = block_sum(l1_acc, red);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 2 * c, s1);
    atomicAdd(sums + 2 * c + 1, s2);
  }
}

__global__ void __launch_bounds__(NT)
ssim_l1_bwd_kernel(const float* __restrict__ render, const float* __restrict__ truth, const float* __restrict__ dmaps,
                   int H, int W, float coef_l1 /* (1-f) / (3HW) */, float* __restrict__ v_render) {
  ST3R_DYN_SMEM_F32(smem);
  float* tm = smem;                         // [3 derivatives][LW][LWP] of the current channel
  float* hs = tm + 3 * LW * LWP;            // [3][LW][HSP]
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t plane = (size_t)H * W;
  float w[11];
#pragma unroll
  for (int k = 0; k < 11; ++k) w[k] = c_win[k];
  const int vx = threadIdx.x & 31, vr0 = (threadIdx.x >> 5) * 4;
  float out[4][3];
  for (int ch = 0; ch < 3; ++ch) {
    __syncthreads();
    const float* src = dmaps + (((size_t)c * 3 + ch) * 3) * plane;
    for (int e = threadIdx.x; e < 3 * LW * LW; e += NT) {
      const int d = e / (LW * LW), rq = e - d * (LW * LW);
      const int r = rq / LW, q = rq - r * LW;
      const int gy = y0 + r - HALO, gx = x0 + q - HALO;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = src[(size_t)d * plane + (size_t)gy * W + gx];
      tm[(d * LW + r) * LWP + q] = v;
    }
    __syncthreads();
    for (int it = threadIdx.x; it < LW * (LT / 8); it += NT) {
      const int r = it >> 2, q0 = (it & 3) * 8;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float* a = tm + (d * LW + r) * LWP + q0;
        float v[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) v[k] = a[k];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 11; ++k) acc += w[k] * v[o + k];
          hs[d * LW * HSP + r * HSP + q0 + o] = acc;
        }
      }
    }
    __syncthreads();
    float g[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float col[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) col[k] = hs[d * LW * HSP + (vr0 + k) * HSP + vx];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) acc += w[k] * col[o + k];
        g[d][o] = acc;
      }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int px = x0 + vx, py = y0 + vr0 + o;
      float r = 0.f;
      if (px < W && py < H) {
        const size_t p = (((size_t)c * H + py) * W + px) * 3 + ch;
        const float x = render[p], y = truth[p];
        const float dd = y - x;
        const float gl1 = dd > 0.f ? -coef_l1 : (dd < 0.f ? coef_l1 : 0.f);
        r = g[0][o] + 2.f * x * g[1][o] + y * g[2][o] + gl1;
      }
      out[o][ch] = r;
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int px = x0 + vx, py = y0 + vr0 + o;
    if (px < W && py < H) {
      float* dst = v_render + (((size_t)c * H + py) * W + px) * 3;
      dst[0] = out[o][0]; dst[1] = out[o][1]; dst[2] = out[o][2];
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

PerDeviceOnce g_win_set;   // __constant__ memory is per device

int set_window() {
  if (g_win_set.done()) return ST3R_OK;
  float g[11], s = 0.f;
  for (int k = 0; k < 11; ++k) {
    float d = (float)(k - 5);
    g[k] = expf(-(d / 1.5f) * (d / 1.5f) / 2.0f);
    s += g[k];
  }
  for (int k = 0; k < 11; ++k) g[k] /= s;
  ST3R_CHECK_CUDA(cudaMemcpyToSymbol(c_win, g, sizeof(g)));
  g_win_set.mark();
  return ST3R_OK;
}

}  // namespace

extern "C" {

// sums [C,2] must be zeroed by the caller; dmaps: 9 * C * H * W floats of scratch (planar, opaque to the caller).
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
  constexpr size_t kSmemFwd = sizeof(float) * (6 * LW * LWP + 5 * LW * HSP);
  static PerDeviceOnce attr_fwd;
  if (!attr_fwd.done()) {
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(ssim_l1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemFwd));
    attr_fwd.mark();
  }
  ssim_l1_fwd_kernel<<<grid, NT, kSmemFwd, stream>>>(render, truth, height, width, coef, dmaps, sums);
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
  constexpr size_t kSmemBwd = sizeof(float) * (3 * LW * LWP + 3 * LW * HSP);
  ssim_l1_bwd_kernel<<<grid, NT, kSmemBwd, stream>>>(render, truth, dmaps, height, width, coef_l1, v_render);
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
