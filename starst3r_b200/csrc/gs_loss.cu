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
// channels is loaded once (a warp per row, 126 consecutive floats of the interleaved [H,W,3] image, de-interleaved
// into planar shared arrays with per-lane offsets computed once), then the separable 11-tap window runs per channel
// with register tiling and vector shared-memory accesses: a thread of the horizontal pass produces 8 consecutive
// columns of one row from 5 LDS.128 per operand (the squares / products are formed once per input, not per tap) and
// stores them with STS.128; a thread of the vertical pass produces 4 consecutive rows of one column from 14 loaded
// rows per statistic.  The row pitches (44 / 36 floats) keep every 128-bit access 16-byte aligned and bank-conflict
// free for the (row, column group) -> lane mapping used here.  The window weights are compile-time immediates of the
// FFMAs.  (Round-2 profile of the first version: 96 M warp instructions per launch at configs[1], 40 % of them
// integer address arithmetic; this version executes about half.)
constexpr int LT = 32;             // output tile edge
constexpr int HALO = 5;
constexpr int LW = LT + 2 * HALO;  // 42
constexpr int TP = 44;             // row pitch of the input tiles (floats)
constexpr int HP = 36;             // row pitch of the horizontally filtered rows
constexpr int NT = 256;
constexpr int HITEMS = LW * (LT / 8);   // (row, group of 8 output columns) items of the horizontal pass
constexpr float C1 = 0.01f * 0.01f;
constexpr float C2 = 0.03f * 0.03f;
static_assert(HITEMS <= NT && LT * (LT / 4) == NT * 1, "one horizontal item / one vertical item per thread");

#ifndef ST3R_HOST_EMU            // (CPU emulator builds, tests/host/, define it themselves)
#define ST3R_DYN_SMEM_F32(name) extern __shared__ float name[]
#endif

// 11-tap Gaussian window, sigma 1.5, normalised (torchmetrics _gaussian): exp(-((k-5)/1.5)^2 / 2) / sum, evaluated in
// fp32 (0.0010283802, 0.0075987582, 0.036000773, 0.10936069, 0.21300554, 0.26601174 and their mirror images).
#define ST3R_SSIM_W0 0x1.0d957p-10f
#define ST3R_SSIM_W1 0x1.f1fdfap-8f
#define ST3R_SSIM_W2 0x1.26eb18p-5f
#define ST3R_SSIM_W3 0x1.bff0fep-4f
#define ST3R_SSIM_W4 0x1.b43c4p-3f
#define ST3R_SSIM_W5 0x1.106562p-2f

// sum_k w[k] * v[k] for 11 consecutive values (k = 0 .. 10), in ascending-k order.
__device__ __forceinline__ float win11(const float* v) {
  float s = ST3R_SSIM_W0 * v[0];
  s = fmaf(ST3R_SSIM_W1, v[1], s); s = fmaf(ST3R_SSIM_W2, v[2], s); s = fmaf(ST3R_SSIM_W3, v[3], s);
  s = fmaf(ST3R_SSIM_W4, v[4], s); s = fmaf(ST3R_SSIM_W5, v[5], s); s = fmaf(ST3R_SSIM_W4, v[6], s);
  s = fmaf(ST3R_SSIM_W3, v[7], s); s = fmaf(ST3R_SSIM_W2, v[8], s); s = fmaf(ST3R_SSIM_W1, v[9], s);
  s = fmaf(ST3R_SSIM_W0, v[10], s);
  return s;
}

// 8 consecutive window sums of v[0 .. 17] -> dst[0 .. 7] (two 128-bit stores; dst is 16-byte aligned).
__device__ __forceinline__ void hconv8(const float* v, float* dst) {
  float4 lo, hi;
  lo.x = win11(v + 0); lo.y = win11(v + 1); lo.z = win11(v + 2); lo.w = win11(v + 3);
  hi.x = win11(v + 4); hi.y = win11(v + 5); hi.z = win11(v + 6); hi.w = win11(v + 7);
  reinterpret_cast<float4*>(dst)[0] = lo;
  reinterpret_cast<float4*>(dst)[1] = hi;
}

// 20 consecutive floats of a 16-byte aligned shared-memory row (the pass uses the first 18).
__device__ __forceinline__ void load20(const float* src, float* v) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float4 t = reinterpret_cast<const float4*>(src)[k];
    v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
  }
}

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

// 4-byte asynchronous global -> shared copies (LDGSTS): the halo tiles never pass through registers, and a whole tile
// is in flight behind one wait.  `valid` = false writes a zero (src-size 0; the address is not dereferenced).
#ifdef ST3R_HOST_EMU
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) { *smem_dst = valid ? *gsrc : 0.f; }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#endif

// Halo tiles of TWO interleaved [H,W,3] images -> planar [3][LW][TP] arrays (zero outside the image).  A warp takes a
// row: its 42 pixels are 126 consecutive floats in global memory, lane l reads floats l, l + 32, l + 64, l + 96.
__device__ __forceinline__ void load_rgb_tiles(const float* __restrict__ img_a, const float* __restrict__ img_b, int H, int W,
                                               int x0, int y0, float* dst_a, float* dst_b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int soff[4], gx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = lane + 32 * j, q = e / 3, ch = e - 3 * q;
    soff[j] = ch * LW * TP + q;
    gx[j] = x0 - HALO + q;
  }
  for (int r = warp; r < LW; r += NT / 32) {
    const int gy = y0 + r - HALO;
    const bool row_ok = gy >= 0 && gy < H;
    const long long gbase = row_ok ? ((long long)gy * W + (x0 - HALO)) * 3 : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = lane + 32 * j;
      if (e < LW * 3) {
        const bool ok = row_ok && gx[j] >= 0 && gx[j] < W;
        const long long gi = ok ? gbase + e : 0;
        cp_async4(dst_a + soff[j] + r * TP, img_a + gi, ok);
        cp_async4(dst_b + soff[j] + r * TP, img_b + gi, ok);
      }
    }
  }
  cp_async_commit();
}

// grid: (tiles_x, tiles_y, C); block: 256 threads.  dmaps is planar scratch: [C][channel 3][derivative 3][H][W].
__global__ void __launch_bounds__(NT, 3)
ssim_l1_fwd_kernel(const float* __restrict__ render, const float* __restrict__ truth, int H, int W,
                   float coef_ssim /* dLoss/dS per interior sample */, float* __restrict__ dmaps,
                   float* __restrict__ sums /* per view: [ssim_sum, l1_sum] */) {
  ST3R_DYN_SMEM_F32(smem);
  float* tx = smem;                         // [3][LW][TP]
  float* ty = tx + 3 * LW * TP;             // [3][LW][TP]
  float* hs = ty + 3 * LW * TP;             // [5][LW][HP]
  __shared__ float red[8];
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t plane = (size_t)H * W;
  load_rgb_tiles(render + (size_t)c * plane * 3, truth + (size_t)c * plane * 3, H, W, x0, y0, tx, ty);
  const int vx = threadIdx.x & 31, vr0 = (threadIdx.x >> 5) * 4;     // vertical pass: column vx, rows vr0 .. vr0 + 3
  float ssim_acc = 0.f, l1_acc = 0.f;
  cp_async_wait<0>();
  for (int ch = 0; ch < 3; ++ch) {
    __syncthreads();        // tiles loaded / the previous channel's vertical pass is done with hs
    if (threadIdx.x < HITEMS) {
      // horizontal pass: item = (row r, group of 8 output columns)
      const int r = threadIdx.x >> 2, q0 = (threadIdx.x & 3) * 8;
      float a[20], b[20], p[18];
      load20(tx + (ch * LW + r) * TP + q0, a);
      load20(ty + (ch * LW + r) * TP + q0, b);
      float* hrow = hs + r * HP + q0;
      hconv8(a, hrow);
      hconv8(b, hrow + LW * HP);
#pragma unroll
      for (int k = 0; k < 18; ++k) p[k] = a[k] * a[k];
      hconv8(p, hrow + 2 * LW * HP);
#pragma unroll
      for (int k = 0; k < 18; ++k) p[k] = b[k] * b[k];
      hconv8(p, hrow + 3 * LW * HP);
#pragma unroll
      for (int k = 0; k < 18; ++k) p[k] = a[k] * b[k];
      hconv8(p, hrow + 4 * LW * HP);
    }
    __syncthreads();
    // vertical pass: 4 output rows of one column from 14 filtered rows per statistic
    float st[5][4];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      float col[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) col[k] = hs[s * LW * HP + (vr0 + k) * HP + vx];
#pragma unroll
      for (int o = 0; o < 4; ++o) st[s][o] = win11(col + o);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int px = x0 + vx, py = y0 + vr0 + o;
      const bool in_img = px < W && py < H;
      const bool interior = px >= HALO && px < W - HALO && py >= HALO && py < H - HALO;
      float dm = 0.f, ds = 0.f, dc = 0.f;
      if (interior) {
        const float mx = st[0][o], my = st[1][o], exx = st[2][o], eyy = st[3][o], exy = st[4][o];
        const float sxx = exx - mx * mx, syy = eyy - my * my, sxy = exy - mx * my;
        const float A1 = 2.f * mx * my + C1, A2 = 2.f * sxy + C2, B1 = mx * mx + my * my + C1, B2 = sxx + syy + C2;
        const float inv = 1.f / (B1 * B2);
        const float S = A1 * A2 * inv;
        ssim_acc += S;
        const float S_mu = 2.f * my * A2 * inv - 2.f * mx * S * (B2 * inv);     // S / B1 = S * B2 / (B1 B2)
        const float S_sx = -S * (B1 * inv);                                     // S / B2
        const float S_c = 2.f * A1 * inv;
        dm = coef_ssim * (S_mu - 2.f * mx * S_sx - my * S_c);
        ds = coef_ssim * S_sx;
        dc = coef_ssim * S_c;
      }
      if (in_img) {
        float* d = dmaps + (((size_t)c * 3 + ch) * 3) * plane + (size_t)py * W + px;
        d[0] = dm; d[plane] = ds; d[2 * plane] = dc;
        const int t = (ch * LW + vr0 + o + HALO) * TP + vx + HALO;
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

// Backward: the three derivative planes of a channel are brought in with cp.async into one of two tile buffers while
// the previous channel is being filtered (the loads of channel ch + 1 are issued before channel ch waits for its own).
__global__ void __launch_bounds__(NT, 3)
ssim_l1_bwd_kernel(const float* __restrict__ render, const float* __restrict__ truth, const float* __restrict__ dmaps,
                   int H, int W, float coef_l1 /* (1-f) / (3HW) */, float* __restrict__ v_render) {
  ST3R_DYN_SMEM_F32(smem);
  float* tm0 = smem;                        // [2 buffers][3 derivatives][LW][TP]
  float* hs = tm0 + 2 * 3 * LW * TP;        // [3][LW][HP]
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
  const size_t plane = (size_t)H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vx = lane, vr0 = warp * 4;
  const int gx_a = x0 - HALO + lane, gx_b = gx_a + 32;        // the two columns of a halo row this lane loads
  const bool ok_a = gx_a >= 0 && gx_a < W, ok_b = lane < LW - 32 && gx_b < W;
  auto issue_tile = [&](int ch) {           // a warp per (derivative, row) of the halo tile
    const float* src = dmaps + (((size_t)c * 3 + ch) * 3) * plane;
    float* tm = tm0 + (ch & 1) * 3 * LW * TP;
    for (int dr = warp; dr < 3 * LW; dr += NT / 32) {
      const int d = dr / LW, r = dr - d * LW;
      const int gy = y0 + r - HALO;
      const bool row_ok = gy >= 0 && gy < H;
      const float* g = src + (size_t)d * plane + (row_ok ? (long long)gy * W : 0);
      float* t = tm + dr * TP;               // (d * LW + r) * TP
      cp_async4(t + lane, g + (row_ok && ok_a ? gx_a : 0), row_ok && ok_a);
      if (lane < LW - 32) cp_async4(t + lane + 32, g + (row_ok && ok_b ? gx_b : 0), row_ok && ok_b);
    }
    cp_async_commit();
  };
  issue_tile(0);
#pragma unroll 1
  for (int ch = 0; ch < 3; ++ch) {
    const float* tm = tm0 + (ch & 1) * 3 * LW * TP;
    // buffer (ch + 1) & 1 was last read by the horizontal pass of channel ch - 1, which every thread left before the
    // barrier that preceded that channel's vertical pass
    if (ch + 1 < 3) { issue_tile(ch + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    float xr[4], yt[4];                      // render / truth pixels of the vertical pass, in flight during the filters
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int px = x0 + vx, py = y0 + vr0 + o;
      xr[o] = yt[o] = 0.f;
      if (px < W && py < H) {
        const size_t p = (((size_t)c * H + py) * W + px) * 3 + ch;
        xr[o] = render[p]; yt[o] = truth[p];
      }
    }
    __syncthreads();        // this channel's tile has landed; the previous channel's vertical pass is done with hs
    if (threadIdx.x < HITEMS) {
      const int r = threadIdx.x >> 2, q0 = (threadIdx.x & 3) * 8;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float v[20];
        load20(tm + (d * LW + r) * TP + q0, v);
        hconv8(v, hs + d * LW * HP + r * HP + q0);
      }
    }
    __syncthreads();
    float g[3][4];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float col[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) col[k] = hs[d * LW * HP + (vr0 + k) * HP + vx];
#pragma unroll
      for (int o = 0; o < 4; ++o) g[d][o] = win11(col + o);
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int px = x0 + vx, py = y0 + vr0 + o;
      if (px < W && py < H) {
        const float x = xr[o], y = yt[o];
        const float dd = y - x;
        const float gl1 = dd > 0.f ? -coef_l1 : (dd < 0.f ? coef_l1 : 0.f);
        v_render[(((size_t)c * H + py) * W + px) * 3 + ch] = g[0][o] + 2.f * x * g[1][o] + y * g[2][o] + gl1;
      }
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

}  // namespace

extern "C" {

// sums [C,2] must be zeroed by the caller; dmaps: 9 * C * H * W floats of scratch (planar, opaque to the caller).
int st3r_gs_loss_fwd(const float* render, const float* truth, int C, int height, int width, float ssim_fac,
                     float* dmaps, float* sums, cudaStream_t stream) {
  ST3R_CHECK_ARG(C >= 0 && height > 10 && width > 10, "st3r_gs_loss_fwd: images must be larger than the 11x11 SSIM window");
  if (C == 0) return ST3R_OK;
  ST3R_CHECK_ARG(render && truth && dmaps && sums, "st3r_gs_loss_fwd: null pointer");
  // loss_view = ... + f * (1 - mean(S))  =>  dLoss/dS = -f / (3 (H-10) (W-10))
  float coef = -ssim_fac / (3.0f * (float)(height - 10) * (float)(width - 10));
  dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, C);
  constexpr size_t kSmemFwd = sizeof(float) * (6 * LW * TP + 5 * LW * HP);
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
  float coef_l1 = (1.0f - ssim_fac) / (3.0f * (float)height * (float)width);
  dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, C);
  constexpr size_t kSmemBwd = sizeof(float) * (6 * LW * TP + 3 * LW * HP);
  static PerDeviceOnce attr_bwd;
  if (!attr_bwd.done()) {
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(ssim_l1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBwd));
    attr_bwd.mark();
  }
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
