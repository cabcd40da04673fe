// Per-tile alpha blending, forward and backward (gsplat rasterize_to_pixels_fwd / _bwd,
// SURVEY.md Appendix A.6; called from starster/gs.py:76-87 and by loss.backward(), gs.py:153).
//
// One CTA of 16x16 threads per (camera, tile); the tile's depth-sorted Gaussians are staged
// through shared memory in batches of 256 as three float4 records (xy+opacity, conic, rgb:
// 48 B per intersection, read once per tile, coalesced 16-byte gathers), and every pixel
// walks the batch with broadcast LDS.128 reads.  Backward walks the same list back to front,
// re-derives the transmittance, reduces each Gaussian's nine partial derivatives across the
// warp with shuffles, accumulates the eight warps of the CTA in shared memory and issues one
// vector atomicAdd per (tile, Gaussian) record to HBM.
// Algorithmic bytes: fwd 40 B per intersection + 20 B per pixel;
// bwd 40 B per intersection + 24 B per pixel + 36 B of gradient per visible (Gaussian, view).
// ST3R_HOST_EMU: test builds that run this file on a CPU SIMT emulator (tests/host/): raster_emu_host.cpp includes the
// kernels only, build_emu_lib.py (ST3R_EMU_WHOLE) compiles the entry points too, with their launches rewritten.
#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
#include "common.cuh"
#include "gs.cuh"
#endif
#ifndef ST3R_EMU_COUNT
#define ST3R_EMU_COUNT(i)   // raster_emu_host.cpp counts how often a marked code path ran (path coverage of the emulated run)
#endif

namespace {

constexpr int TILE = 16;
constexpr int BLOCK = TILE * TILE;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float ALPHA_MAX = 0.999f;
constexpr float T_MIN = 1e-4f;

struct TileRange { int lo, hi; };

__device__ __forceinline__ TileRange tile_range(const int32_t* __restrict__ offsets, const int32_t* __restrict__ n_isect,
                                                int t, int total_tiles) {
  TileRange r;
  r.lo = offsets[t];
  r.hi = (t + 1 < total_tiles) ? offsets[t + 1] : *n_isect;
  return r;
}

// sigma = 0.5 (a dx^2 + c dy^2) + b dx dy in ONE fixed evaluation order for every blend kernel of this file (explicit
// fused multiply-adds; the compiler's own contraction choices could differ between kernels and flip a borderline
// alpha >= 1/255 test between the forward and the backward pass).
__device__ __forceinline__ float blend_sigma(float a, float b, float c, float dx, float dy) {
  return fmaf(b * dx, dy, 0.5f * fmaf(a * dx, dx, (c * dy) * dy));
}

// exp(-sigma) as one ex2.approx (the value __expf returns wherever alpha can pass the 1/255 test).
__device__ __forceinline__ float exp_neg(float sigma) {
#ifdef ST3R_HOST_EMU
  return expf(-sigma);
#else
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(sigma * -1.4426950408889634f));
  return r;
#endif
}

// Conservative set of tile rows (bit i = row i of the 16x16 tile) on which a Gaussian can pass the
// alpha >= 1/255 test: opac * exp(-sigma) >= 1/255  <=>  d^T Q d <= 2 ln(255 opac), whose bounding box has the
// vertical half-extent sqrt(2 L cov_yy), cov_yy = a / (a c - b^2).  A small margin keeps the cull strictly
// conservative w.r.t. fp32 rounding of sigma, so results are unchanged; a warp (two pixel rows) skips every
// Gaussian whose mask misses its rows without evaluating a single alpha.
__device__ __forceinline__ uint32_t row_mask(const float4 A, const float4 B, int tile_y0) {
  const float L = logf(255.0f * A.z);
  if (!(L > 0.f)) return 0u;
  const float det = B.x * B.z - B.y * B.y;
  if (!(det > 0.f)) return 0xffffu;
  const float hy = sqrtf(2.0f * L * B.x / det) * 1.0005f + 2e-3f;
  const float lo = ceilf(A.y - hy - 0.5f) - (float)tile_y0, hi = floorf(A.y + hy - 0.5f) - (float)tile_y0;
  if (hi < 0.f || lo > 15.f) return 0u;
  const int ilo = max(0, (int)lo), ihi = min(15, (int)hi);
  return ((2u << ihi) - 1u) & ~((1u << ilo) - 1u);
}

// Per-warp visit lists.  Every warp owns two tile rows; testing each staged Gaussian's row mask in every warp cost
// ~18 instructions per (warp, Gaussian) although 70 % of the tests fail (ncu source view, r01b: 29 % of all
// instructions of the backward kernel).  Instead the mask of the Gaussian held by thread `tr` is transposed once per
// batch with eight ballots per warp into list[row pair][word] (bit = Gaussian of the batch), and each warp walks
// only the set bits of its own 256-bit list.
__device__ __forceinline__ void publish_visit_lists(uint32_t m, uint32_t (*list)[8]) {
  uint32_t mine = 0;
#pragma unroll
  for (int rp = 0; rp < 8; ++rp) {
    const uint32_t b = __ballot_sync(0xffffffffu, (m >> (2 * rp)) & 3u);
    if (lane_id() == rp) mine = b;
  }
  if (lane_id() < 8) list[lane_id()][threadIdx.x >> 5] = mine;
}

__global__ void __launch_bounds__(BLOCK)
raster_fwd_kernel(const int32_t* __restrict__ offsets, const int32_t* __restrict__ n_isect,
                  const uint32_t* __restrict__ flatten, const float4* __restrict__ geomA,
                  const float4* __restrict__ geomB, const float4* __restrict__ rgb, int C, int W, int H, int tile_w,
                  int tile_h, float* __restrict__ render, float* __restrict__ alphas, int32_t* __restrict__ last_ids,
                  unsigned long long* __restrict__ n_blend) {
  __shared__ float4 sA[BLOCK], sB[BLOCK], sC[BLOCK];
  __shared__ uint32_t sL[8][8];
  const int c = blockIdx.y, tile = blockIdx.x;
  const int tyi = tile / tile_w, txi = tile - tyi * tile_w;
  const int tr = threadIdx.x;
  const int i = tyi * TILE + (tr >> 4), j = txi * TILE + (tr & 15);
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const bool inside = i < H && j < W;
  const int wrp = tr >> 5;                          // this warp owns tile rows 2 wrp, 2 wrp + 1
  bool done = !inside;
  const TileRange rg = tile_range(offsets, n_isect, c * tile_w * tile_h + tile, C * tile_w * tile_h);
  const int nb = (rg.hi - rg.lo + BLOCK - 1) / BLOCK;
  float T = 1.0f, pr = 0.f, pg = 0.f, pb = 0.f;
  int cur = 0, blends = 0;
  for (int b = 0; b < nb; ++b) {
    if (__syncthreads_count(done) == BLOCK) break;
    const int start = rg.lo + b * BLOCK;
    const int idx = start + tr;
    uint32_t m = 0;
    if (idx < rg.hi) {
      uint32_t e = flatten[idx];
      const float4 A = geomA[e], B = geomB[e];
      sA[tr] = A;
      sB[tr] = B;
      sC[tr] = rgb[e];
      m = row_mask(A, B, tyi * TILE);
    }
    publish_visit_lists(m, sL);
    __syncthreads();
    for (int wi = 0; wi < 8; ++wi) {
      uint32_t bits = sL[wrp][wi];        // warp-uniform: the Gaussians of this batch that can touch my two rows
      if (bits && __all_sync(0xffffffffu, done)) break;
      while (bits) {
        const int t = wi * 32 + __ffs(bits) - 1;
        bits &= bits - 1;
        if (done) continue;
        const float4 A = sA[t], B = sB[t];
        const float dx = A.x - px, dy = A.y - py;
        const float sigma = blend_sigma(B.x, B.y, B.z, dx, dy);
        const float alpha = fminf(ALPHA_MAX, A.z * exp_neg(sigma));
        if (sigma < 0.f || alpha < ALPHA_MIN) continue;
        const float nT = T * (1.0f - alpha);
        if (nT <= T_MIN) { done = true; continue; }
        const float w = alpha * T;
        const float4 col = sC[t];
        pr += col.x * w; pg += col.y * w; pb += col.z * w;
        cur = start + t;
        T = nT;
        ++blends;
      }
    }
  }
  if (inside) {
    const size_t p = ((size_t)c * H + i) * W + j;
    render[3 * p] = pr; render[3 * p + 1] = pg; render[3 * p + 2] = pb;
    alphas[p] = 1.0f - T;
    last_ids[p] = cur;
  }
  if (n_blend) {
    for (int off = 16; off; off >>= 1) blends += __shfl_xor_sync(0xffffffffu, blends, off);
    if (lane_id() == 0 && blends) atomicAdd(n_blend, (unsigned long long)blends);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// 1 - alpha lies in [1e-3, 1]: rcp.approx (MUFU.RCP, 1 ulp) instead of the 15-instruction IEEE division
__device__ __forceinline__ float fast_rcp(float x) {
#ifdef ST3R_HOST_EMU
  return 1.0f / x;
#else
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#endif
}

__global__ void __launch_bounds__(BLOCK)
raster_bwd_kernel(const int32_t* __restrict__ offsets, const int32_t* __restrict__ n_isect,
                  const uint32_t* __restrict__ flatten, const float4* __restrict__ geomA,
                  const float4* __restrict__ geomB, const float4* __restrict__ rgb, int C, int W, int H, int tile_w,
                  int tile_h, const float* __restrict__ alphas, const int32_t* __restrict__ last_ids,
                  const float* __restrict__ v_render, const float* __restrict__ v_alphas,
                  float4* __restrict__ v_geomA, float4* __restrict__ v_geomB, float4* __restrict__ v_rgb) {
  __shared__ float4 sA[BLOCK], sB[BLOCK], sC[BLOCK];
  __shared__ uint32_t sE[BLOCK];
  __shared__ uint32_t sL[8][8];
  __shared__ float acc[BLOCK][9];  // per-batch gradient accumulators (xy 2, opac 1, conic 3, rgb 3)
  const int c = blockIdx.y, tile = blockIdx.x;
  const int tyi = tile / tile_w, txi = tile - tyi * tile_w;
  const int tr = threadIdx.x;
  const int i = tyi * TILE + (tr >> 4), j = txi * TILE + (tr & 15);
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const bool inside = i < H && j < W;
  const int wrp = tr >> 5;
  const TileRange rg = tile_range(offsets, n_isect, c * tile_w * tile_h + tile, C * tile_w * tile_h);
  if (rg.hi <= rg.lo) return;
  const size_t p = ((size_t)c * H + min(i, H - 1)) * W + min(j, W - 1);
  const float T_final = 1.0f - alphas[p];
  float T = T_final;
  // The colour accumulated behind the current Gaussian only enters dL/dalpha through its dot product with this pixel's
  // upstream gradient: one running scalar instead of three colour sums,
  // dL/dalpha = T (c . v) + (T_final v_a - behind . v) / (1 - alpha).
  float behind_v = 0.f;
  const int bin_final = inside ? last_ids[p] : 0;
  float vr = 0.f, vg = 0.f, vb = 0.f, va = 0.f;
  if (inside) {
    vr = v_render[3 * p]; vg = v_render[3 * p + 1]; vb = v_render[3 * p + 2];
    va = v_alphas ? v_alphas[p] : 0.f;
  }
  const float tf_va = T_final * va;
  int warp_bin_final = bin_final;
  for (int off = 16; off; off >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, off));

  const int nb = (rg.hi - rg.lo + BLOCK - 1) / BLOCK;
  for (int b = 0; b < nb; ++b) {
    __syncthreads();
    const int batch_end = rg.hi - 1 - BLOCK * b;
    const int bs = min(BLOCK, batch_end + 1 - rg.lo);
    const int idx = batch_end - tr;
    uint32_t m = 0;
    if (idx >= rg.lo) {
      uint32_t e = flatten[idx];
      const float4 A = geomA[e], B = geomB[e];
      sE[tr] = e;
      sA[tr] = A;
      sB[tr] = B;
      sC[tr] = rgb[e];
      m = row_mask(A, B, tyi * TILE);
    }
    publish_visit_lists(m, sL);
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[tr][k] = 0.f;
    __syncthreads();
    // Batch slot t holds sorted position batch_end - t: ascending t walks back to front.  Slots before t0 lie behind
    // the last Gaussian any pixel of this warp blended in the forward pass.
    const int t0 = max(0, batch_end - warp_bin_final);
    for (int wi = t0 >> 5; wi < 8; ++wi) {
     uint32_t bits = sL[wrp][wi];
     if (wi == (t0 >> 5)) bits &= 0xffffffffu << (t0 & 31);
     while (bits) {
      const int t = wi * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      bool valid = inside && (batch_end - t <= bin_final);
      float alpha = 0.f, opac = 0.f, vis = 0.f, dx = 0.f, dy = 0.f;
      float4 B = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        const float4 A = sA[t];
        B = sB[t];
        opac = A.z;
        dx = A.x - px; dy = A.y - py;
        const float sigma = blend_sigma(B.x, B.y, B.z, dx, dy);
        vis = exp_neg(sigma);
        alpha = fminf(ALPHA_MAX, opac * vis);
        if (sigma < 0.f || alpha < ALPHA_MIN) valid = false;
      }
      if (!__any_sync(0xffffffffu, valid)) continue;
      float g[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const float ra = fast_rcp(1.0f - alpha);
        T *= ra;
        const float fac = alpha * T;
        const float4 col = sC[t];
        g[6] = fac * vr; g[7] = fac * vg; g[8] = fac * vb;
        const float cv = col.x * vr + col.y * vg + col.z * vb;
        const float v_alpha = T * cv + ra * (tf_va - behind_v);
        if (opac * vis <= ALPHA_MAX) {
          const float v_sigma = -opac * vis * v_alpha;
          g[3] = 0.5f * v_sigma * dx * dx;
          g[4] = v_sigma * dx * dy;
          g[5] = 0.5f * v_sigma * dy * dy;
          g[0] = v_sigma * (B.x * dx + B.y * dy);
          g[1] = v_sigma * (B.y * dx + B.z * dy);
          g[2] = vis * v_alpha;
        }
        behind_v += fac * cv;
      }
      // Transposed butterfly: 8 of the 9 sums are reduced with 4+2+1+2 shuffles (instead of 8 x 5) by halving the
      // number of live values at each exchange; afterwards lane l holds the warp total of value (l >> 2) and the 8
      // lanes with (l & 3) == 0 issue their shared-memory atomics in parallel.  The 9th value takes the plain tree.
      {
        const unsigned full = 0xffffffffu;
        const int lane = lane_id();
        float a[4], b2[2], c1;
        const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float send = h16 ? g[k] : g[k + 4];
          float keep = h16 ? g[k + 4] : g[k];
          a[k] = keep + __shfl_xor_sync(full, send, 16);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float send = h8 ? a[k] : a[k + 2];
          float keep = h8 ? a[k + 2] : a[k];
          b2[k] = keep + __shfl_xor_sync(full, send, 8);
        }
        {
          float send = h4 ? b2[0] : b2[1];
          float keep = h4 ? b2[1] : b2[0];
          c1 = keep + __shfl_xor_sync(full, send, 4);
        }
        c1 += __shfl_xor_sync(full, c1, 2);
        c1 += __shfl_xor_sync(full, c1, 1);
        const float g8 = warp_sum(g[8]);
        // value index held by this lane: bit 4 of the lane selects +4, bit 3 selects +2, bit 2 selects +1
        const int vi = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
        // One shared-memory update for all nine sums: lanes 0, 4, ..., 28 carry values 0..7 and lane 1 the ninth
        // (fp32 shared atomics are compare-and-swap loops, ATOMS.CAST.SPIN; nine distinct words, one pass).
        const bool ninth = lane == 1;
        const float sum = ninth ? g8 : c1;
        if (((lane & 3) == 0 || ninth) && sum != 0.f) atomicAdd(&acc[t][ninth ? 8 : vi], sum);
      }
     }
    }
    __syncthreads();
    if (tr < bs) {
      const uint32_t e = sE[tr];
      const float* a = acc[tr];
      if (a[0] != 0.f || a[1] != 0.f || a[2] != 0.f || a[3] != 0.f || a[4] != 0.f || a[5] != 0.f || a[6] != 0.f ||
          a[7] != 0.f || a[8] != 0.f) {
        atomicAdd(v_geomA + e, make_float4(a[0], a[1], a[2], 0.f));
        atomicAdd(v_geomB + e, make_float4(a[3], a[4], a[5], 0.f));
        atomicAdd(v_rgb + e, make_float4(a[6], a[7], a[8], 0.f));
      }
    }
  }
}


// The nine gradient terms of one (pixel, Gaussian) contribution from fac = alpha T and w = vis dL/dalpha (w = 0 when
// alpha was clamped to ALPHA_MAX): xy 2, opacity 1, conic 3, rgb 3 - same expressions as in raster_bwd_kernel.
__device__ __forceinline__ void blend_grad_terms(const float4 A, const float4 B, float dx, float dy, float fac, float w,
                                                 float vr, float vg, float vb, float* g) {
  const float v_sigma = -A.z * w;
  g[0] = v_sigma * (B.x * dx + B.y * dy);
  g[1] = v_sigma * (B.y * dx + B.z * dy);
  g[2] = w;
  g[3] = 0.5f * v_sigma * dx * dx;
  g[4] = v_sigma * dx * dy;
  g[5] = 0.5f * v_sigma * dy * dy;
  g[6] = fac * vr; g[7] = fac * vg; g[8] = fac * vb;
}

// Warp-collective: reduces g[0..8] over the warp and adds the totals to a9[0..8] (shared memory).
__device__ __forceinline__ void butterfly9_to_shared(const float* g, float* a9) {
  const unsigned full = 0xffffffffu;
  const int lane = lane_id();
  float a[4], b2[2], c1;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float send = h16 ? g[k] : g[k + 4];
    float keep = h16 ? g[k + 4] : g[k];
    a[k] = keep + __shfl_xor_sync(full, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float send = h8 ? a[k] : a[k + 2];
    float keep = h8 ? a[k + 2] : a[k];
    b2[k] = keep + __shfl_xor_sync(full, send, 8);
  }
  {
    float send = h4 ? b2[0] : b2[1];
    float keep = h4 ? b2[1] : b2[0];
    c1 = keep + __shfl_xor_sync(full, send, 4);
  }
  c1 += __shfl_xor_sync(full, c1, 2);
  c1 += __shfl_xor_sync(full, c1, 1);
  const float g8 = warp_sum(g[8]);
  const int vi = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
  const bool ninth = lane == 1;
  const float sum = ninth ? g8 : c1;
  if (((lane & 3) == 0 || ninth) && sum != 0.f) atomicAdd(a9 + (ninth ? 8 : vi), sum);
}

// ---- fragment-pool kernels (forward and backward) --------------------------------------------------------------------
// With the reference's 3e-3 initial scale a splat covers a handful of pixels, and the pixel-parallel walk above spends
// its instructions on alpha tests that fail (forward: 4.5 % hit rate) and on warp reductions in which ~6 of 32 lanes
// carry a value (ncu, round 1: both kernels issue-bound).  Here the work is split by what it is parallel in:
//   0. stage: thread t loads Gaussian t of the batch (up to PG per batch), derives its bounding box inside the tile
//      (conservative extents of the alpha >= 1/255 ellipse, clipped to the tile and the image) and gets a run of pool
//      slots, one per box pixel, from a block-wide exclusive scan of the box areas (small and tile-sized splats mix
//      freely; what does not fit into the pool is left to the next batch).  The same scan compacts the Gaussians that
//      have a box at all; each leaves a 48-byte record (geometry, box, slots) in compacted = depth order;
//   A. slot-parallel: the S slots of the batch are cut into 256 equal chunks.  A thread finds the Gaussian its chunk
//      starts in by binary search over the run offsets and walks its chunk in ONE flat loop (moving on to the next
//      record is a predicated block, so all lanes of all warps run the same trip count): alpha once per box pixel,
//      stored in the slot, and the Gaussian's bit set in the mask of every pixel that passes the alpha test;
//   B. pixel-parallel: every pixel walks the set bits of ITS OWN mask in depth order - no alpha test, no warp
//      collective, only the recurrence that is sequential per pixel (transmittance, colour; backward: colour behind,
//      dL/dalpha, results written back to the slot);
//   C. (backward) slot-parallel again, same chunks: a thread accumulates the moments of its slots in registers and
//      issues one vector atomic triple per Gaussian run it touches: no shared-memory atomics, no warp reductions.
// Batches of LARGE splats (boxes of PDENSE pixels and more on average: most pixels of the tile pass the alpha test
// anyway) skip the pool: the pixels walk the batch's records directly, a warp skipping the boxes that miss its two
// rows (forward: gsplat's loop; backward: the nine-value warp reduction of raster_bwd_kernel per contributing visit).
// Results: same expressions in the same per-pixel order as the kernels above (the forward is bit-identical); the
// backward's per-Gaussian sums are added in a different order.
// Batch geometry, per kernel (B200 sweeps at configs[1], `scripts/build_variants.sh` + `scripts/gpu_libvar.sh`,
// profiles/r02af_pool_sweeps.jsonl): the forward is fastest with 256 Gaussians / 4096 slots per batch; the backward, whose
// pool holds two values per slot, with 128 Gaussians / 2304 slots and 48 registers (5 CTAs per SM instead of 3:
// 0.753 -> 0.709 ms with the 3e-3 initial scales, 1.688 -> 1.507 ms with log-normal scales).  A slot count of (odd
// number) x BLOCK keeps every thread's chunk full when a batch fills the pool (see pool_chunk).
// -DST3R_POOL_PG / -DST3R_POOL_SLOTS set both kernels (A/B builds).
#ifdef ST3R_POOL_PG
#define ST3R_POOL_PG_FWD ST3R_POOL_PG
#define ST3R_POOL_PG_BWD ST3R_POOL_PG
#endif
#ifdef ST3R_POOL_SLOTS
#define ST3R_POOL_SLOTS_FWD ST3R_POOL_SLOTS
#define ST3R_POOL_SLOTS_BWD ST3R_POOL_SLOTS
#endif
#ifndef ST3R_POOL_PG_FWD
#define ST3R_POOL_PG_FWD 256
#endif
#ifndef ST3R_POOL_SLOTS_FWD
#define ST3R_POOL_SLOTS_FWD 4096
#endif
#ifndef ST3R_POOL_PG_BWD
#define ST3R_POOL_PG_BWD 128
#endif
#ifndef ST3R_POOL_SLOTS_BWD
#define ST3R_POOL_SLOTS_BWD 2304
#endif
#ifndef ST3R_POOL_DENSE
#define ST3R_POOL_DENSE 96
#endif
#ifndef ST3R_POOL_MINB_FWD
#define ST3R_POOL_MINB_FWD 1
#endif
#ifndef ST3R_POOL_MINB_BWD
#define ST3R_POOL_MINB_BWD 5
#endif
template <bool kBwd>
struct PoolCfg {
  static constexpr int PG = kBwd ? ST3R_POOL_PG_BWD : ST3R_POOL_PG_FWD;              // Gaussians per batch (one per thread of the first PG threads)
  static constexpr int PW = PG / 32;                                                 // mask words per pixel
  static constexpr int PSLOTS = kBwd ? ST3R_POOL_SLOTS_BWD : ST3R_POOL_SLOTS_FWD;    // pool slots per batch (a Gaussian takes at most 256)
  static_assert(PG == 64 || PG == 128 || PG == 256, "pool batch: 64, 128 or 256 Gaussians");
  static_assert(PSLOTS >= 9 * PG && PSLOTS >= 256 && PSLOTS < (1 << 20),
                "the pool doubles as the dense batches' [PG][9] accumulators and holds any one box");
};
constexpr int PDENSE = ST3R_POOL_DENSE;   // mean box area from which a batch is walked pixel-parallel

// Box of tile pixels (rows r0..r1, columns c0..c1, clipped to rmax / cmax) on which opac * exp(-sigma) >= 1/255 is
// possible; see row_mask for the bound (here with fast intrinsics and a wider safety margin: the box only has to
// contain the ellipse).  Returns the area (0: cannot contribute); geo = r0 | c0 << 4 | (wc - 1) << 8 | (nr - 1) << 12.
__device__ __forceinline__ int pool_box(const float4 A, const float4 B, float px0, float py0, int rmax, int cmax,
                                        uint32_t& geo) {
  geo = 0;
  if (!(A.z > ALPHA_MIN)) return 0;                      // opac * exp(-sigma) <= opac < 1/255 (also NaN)
  int r0 = 0, r1 = rmax, c0 = 0, c1 = cmax;
  const float det = B.x * B.z - B.y * B.y;
  if (det > 0.f) {
#ifdef ST3R_HOST_EMU
    const float L2 = 2.0f * logf(255.0f * A.z) / det;
#else
    const float L2 = __fdividef(2.0f * __logf(255.0f * A.z), det);
#endif
    const float hy = sqrtf(L2 * B.x) * 1.001f + 4e-3f;
    const float hx = sqrtf(L2 * B.z) * 1.001f + 4e-3f;
    // (NaN / inf extents fall through fmaxf / fminf to the whole tile)
    r0 = max(0, (int)fmaxf(ceilf(A.y - hy - py0), -1.0f));
    r1 = min(rmax, (int)fminf(floorf(A.y + hy - py0), (float)TILE));
    c0 = max(0, (int)fmaxf(ceilf(A.x - hx - px0), -1.0f));
    c1 = min(cmax, (int)fminf(floorf(A.x + hx - px0), (float)TILE));
  }
  if (r1 < r0 || c1 < c0) return 0;
  const int wc = c1 - c0 + 1, nr = r1 - r0 + 1;
  geo = (uint32_t)r0 | (uint32_t)c0 << 4 | (uint32_t)(wc - 1) << 8 | (uint32_t)(nr - 1) << 12;
  return wc * nr;
}

// One staged batch: Gaussians taken from the sorted list, how many of them have a box (= records), pool slots handed
// out (0 in a dense batch), whether the batch is walked pixel-parallel, barrier count of the caller's flag.
struct PoolBatch { int n, nc, S, n_done; bool dense; };

// Shared memory of one CTA.  kBwd: two pool values per slot and the tile's upstream gradients.
template <bool kBwd>
struct PoolSmem {
  static constexpr int PG = PoolCfg<kBwd>::PG, PW = PoolCfg<kBwd>::PW, PSLOTS = PoolCfg<kBwd>::PSLOTS;
  // per Gaussian WITH a box, in depth order (compacted): the record a walk loads when it enters the box
  float4 recA[PG];            // x, y, opacity, pixel-centre x of the box's first column
  float4 recB[PG];            // conic a, b, c, pixel-centre y of the box's first row
  int4 recC[PG];              // first slot behind the box | tile pixel of its first slot | wc, t << 5, area << 13,
                              // r0 << 23, (nr - 1) << 27 | entry id
  float4 col[PG];             // colour
  int cbase[BLOCK];           // first pool slot (ascending); INT_MAX behind the last record
  int2 boxB[PG];              // by batch slot t: pool slot of tile pixel (row, col) = x + row * (y & 31) + col; y >> 8: record
  uint32_t mask[PW][BLOCK];   // bit t of word w of pixel p: Gaussian 32 w + t contributes to p
  int scan[BLOCK / 32], take[BLOCK / 32];
  uint32_t vlist[BLOCK / 32][BLOCK / 32];   // dense batches: bit ci of vlist[w]: record ci touches the rows of warp w
  float poolA[PSLOTS];        // forward: alpha.  Backward, phase A: alpha; phase B: alpha T.  Dense backward: sums [PG][9]
  float poolV[kBwd ? PSLOTS : 1];    // backward, phase A: vis (0 when alpha was clamped); phase B: vis dL/dalpha
  float4 v[kBwd ? BLOCK : 1];        // backward: upstream colour gradient of every pixel of the tile
};

// Stage one batch (two barriers).  `have`: this thread's slot holds Gaussian e.
template <bool kBwd>
__device__ __forceinline__ PoolBatch pool_stage(bool have, uint32_t e, const float4* __restrict__ geomA,
                                                const float4* __restrict__ geomB, const float4* __restrict__ rgb,
                                                float px0, float py0, int rmax, int cmax, int done_pred,
                                                PoolSmem<kBwd>& sm) {
  constexpr int PW = PoolCfg<kBwd>::PW, PSLOTS = PoolCfg<kBwd>::PSLOTS;
  const unsigned full = 0xffffffffu;
  const int tr = threadIdx.x, lane = tr & 31, wrp = tr >> 5;
  int area = 0;
  uint32_t geo = 0;
  float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A, Cc = A;
  if (have) {
    A = geomA[e];
    B = geomB[e];
    Cc = rgb[e];
    area = pool_box(A, B, px0, py0, rmax, cmax, geo);
  }
  // one scan for both prefixes: slots in the low 20 bits, Gaussians with a box above
  const int mine = area | (area ? 1 << 20 : 0);
  int incl = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = (int)__shfl_up_sync(full, (unsigned)incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) sm.scan[wrp] = incl;
  sm.cbase[tr] = 0x7fffffff;
#pragma unroll
  for (int w = 0; w < PW; ++w) sm.mask[w][tr] = 0u;
  PoolBatch pb;
  pb.n_done = __syncthreads_count(done_pred);
  int before = 0, total = 0;
#pragma unroll
  for (int k = 0; k < BLOCK / 32; ++k) {
    const int v = sm.scan[k];
    before += (k < wrp) ? v : 0;
    total += v;
  }
  const int excl = before + incl - mine;
  const int base = excl & 0xfffff, ci = excl >> 20;
  // large splats: no pool, every candidate is taken
  pb.dense = (total & 0xfffff) >= PDENSE * (total >> 20);
  const bool fits = have && (pb.dense || base + area <= PSLOTS);   // monotone along t: a prefix of the batch is staged
  if (have && !fits) ST3R_EMU_COUNT(12);                  // Gaussians left to the next batch (pool full)
  if (!fits) area = 0;
  const uint32_t staged = __ballot_sync(full, fits);
  const uint32_t boxed = __ballot_sync(full, area != 0);
  const int top = __reduce_max_sync(full, fits ? base + area : 0);
  if (lane == 0) sm.take[wrp] = __popc(staged) | __popc(boxed) << 9 | top << 18;
  if (area) {
    const int r0 = (int)(geo & 15u), c0 = (int)((geo >> 4) & 15u), wc = (int)((geo >> 8) & 15u) + 1;
    sm.boxB[tr] = make_int2(base - r0 * wc - c0, wc | ci << 8);
    sm.recA[ci] = make_float4(A.x, A.y, A.z, px0 + (float)c0);
    sm.recB[ci] = make_float4(B.x, B.y, B.z, py0 + (float)r0);
    sm.recC[ci] = make_int4(base + area, r0 * TILE + c0,
                            wc | tr << 5 | area << 13 | r0 << 23 | (int)((geo >> 12) & 15u) << 27, (int)e);
    sm.col[ci] = Cc;
    sm.cbase[ci] = base;
  }
  __syncthreads();
  pb.n = 0;
  pb.nc = 0;
  pb.S = 0;
#pragma unroll
  for (int k = 0; k < BLOCK / 32; ++k) {
    const int v = sm.take[k];
    pb.n += v & 511;
    pb.nc += (v >> 9) & 511;
    pb.S = max(pb.S, v >> 18);
  }
  if (pb.dense) pb.S = 0;
  return pb;
}

// Candidates of the next batch: all PG while whole batches fit into the pool, else about twice what the pool took.
template <bool kBwd>
__device__ __forceinline__ int pool_next_cand(int cand, int taken) {
  constexpr int PG = PoolCfg<kBwd>::PG;
  return taken == cand ? min(PG, 2 * cand) : max(32, min(PG, 2 * taken));
}

// The chunk [s, s_end) of pool slots of this thread and the record of the Gaussian its first slot is in.
__device__ __forceinline__ bool pool_chunk(int S, const int* cbase, int& s, int& s_end, int& ci) {
  // odd: the lanes of a warp write their slots to distinct banks (even lengths, i.e. 2- / 4-way conflicts but up to
  // one slot less per thread, measured 1.3 % slower on the B200: profiles/r02af_pool_sweeps.jsonl)
  const int q = ((S + BLOCK - 1) / BLOCK) | 1;
  s = (int)threadIdx.x * q;
  s_end = min(S, s + q);
  if (s >= s_end) return false;
  ci = 0;                                                 // largest ci with cbase[ci] <= s (cbase[0] = 0)
#pragma unroll
  for (int step = BLOCK / 2; step; step >>= 1)
    if (cbase[ci + step] <= s) ci += step;                // (index <= BLOCK - 1)
  return true;
}

// Walk state of one thread inside the box of the Gaussian it is currently in.
struct PoolWalk {
  float x, y, opac, qa, qb, qc;
  float pxc0, pxc, pyr;      // pixel-centre coordinates of the box's first column, the current column, the current row
  int cc, wc, run_end, pix;  // column inside the box, box width, first slot behind the box, current tile pixel
  int t, e;                  // batch slot (depth order), entry id
  // Enter record ci at slot s (s > its first slot only for the first Gaussian of a chunk).
  template <bool kBwd>
  __device__ __forceinline__ void enter(const PoolSmem<kBwd>& sm, int ci, int s, bool first) {
    const float4 ra = sm.recA[ci], rb = sm.recB[ci];
    const int4 rc = sm.recC[ci];
    x = ra.x; y = ra.y; opac = ra.z; pxc0 = ra.w;
    qa = rb.x; qb = rb.y; qc = rb.z; pyr = rb.w;
    run_end = rc.x;
    pix = rc.y;
    wc = rc.z & 31;
    t = (rc.z >> 5) & 255;
    e = rc.w;
    cc = 0;
    pxc = pxc0;
    if (first) {
      const int o = s - (run_end - ((rc.z >> 13) & 1023));
      if (o) {
        const int rr = (int)(((float)o + 0.5f) / (float)wc);      // o / wc (o < 256, wc <= 16: far from the rounding edge)
        cc = o - rr * wc;
        pxc = pxc0 + (float)cc;
        pyr += (float)rr;
        pix += rr * TILE + cc;
      }
    }
  }
  // Next slot of the same box.
  __device__ __forceinline__ void step() {
    pxc += 1.0f;
    ++pix;
    if (++cc == wc) { cc = 0; pxc = pxc0; pyr += 1.0f; pix += TILE - wc; }
  }
};

// Dense batches: per-warp visit lists (warp w owns tile rows 2 w, 2 w + 1).  Thread ci holds the rows of record ci's box;
// eight ballots per warp transpose that into vlist[w][word] (bit = record), and each warp walks only its set bits.
template <bool kBwd>
__device__ __forceinline__ void pool_visit_lists(PoolSmem<kBwd>& sm, int nc) {
  const int tr = threadIdx.x;
  uint32_t m = 0;
  if (tr < nc) {
    const int rcz = sm.recC[tr].z;
    const int w0 = ((rcz >> 23) & 15) >> 1, w1 = (((rcz >> 23) & 15) + ((rcz >> 27) & 15)) >> 1;
    m = ((2u << w1) - 1u) & ~((1u << w0) - 1u);
  }
  uint32_t mine = 0;
#pragma unroll
  for (int w = 0; w < BLOCK / 32; ++w) {
    const uint32_t b = __ballot_sync(0xffffffffu, (m >> w) & 1u);
    if (lane_id() == w) mine = b;
  }
  if (lane_id() < BLOCK / 32) sm.vlist[lane_id()][tr >> 5] = mine;
  __syncthreads();
}

#ifdef ST3R_HOST_EMU
#define POOL_SMEM(kBwd) __shared__ PoolSmem<kBwd> sm
#else
#define POOL_SMEM(kBwd)                                 \
  extern __shared__ __align__(16) char pool_smem_raw[]; \
  PoolSmem<kBwd>& sm = *reinterpret_cast<PoolSmem<kBwd>*>(pool_smem_raw)
#endif

__global__ void __launch_bounds__(BLOCK, ST3R_POOL_MINB_FWD)
raster_fwd_pool_kernel(const int32_t* __restrict__ offsets, const int32_t* __restrict__ n_isect,
                       const uint32_t* __restrict__ flatten, const float4* __restrict__ geomA,
                       const float4* __restrict__ geomB, const float4* __restrict__ rgb, int C, int W, int H, int tile_w,
                       int tile_h, float* __restrict__ render, float* __restrict__ alphas,
                       int32_t* __restrict__ last_ids, unsigned long long* __restrict__ n_blend) {
  POOL_SMEM(false);
  const int c = blockIdx.y, tile = blockIdx.x;
  const int tyi = tile / tile_w, txi = tile - tyi * tile_w;
  const int tr = threadIdx.x, lane = tr & 31, wrp = tr >> 5;
  const int myrow = tr >> 4, mycol = tr & 15;
  const int i = tyi * TILE + myrow, j = txi * TILE + mycol;
  const bool inside = i < H && j < W;
  const float px0 = (float)(txi * TILE) + 0.5f, py0 = (float)(tyi * TILE) + 0.5f;   // centre of the tile's first pixel
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const int rmax = min(TILE, H - tyi * TILE) - 1, cmax = min(TILE, W - txi * TILE) - 1;
  bool done = !inside;
  const TileRange rg = tile_range(offsets, n_isect, c * tile_w * tile_h + tile, C * tile_w * tile_h);
  float T = 1.0f, pr = 0.f, pg = 0.f, pb_ = 0.f;
  int cur = 0, blends = 0;
  constexpr int PG = PoolCfg<false>::PG, PW = PoolCfg<false>::PW;
  int pos = rg.lo, cand = PG;
  while (pos < rg.hi) {
    const int idx = pos + tr;
    const bool have = tr < cand && idx < rg.hi;
    const uint32_t e = have ? flatten[idx] : 0u;
    const PoolBatch pb = pool_stage<false>(have, e, geomA, geomB, rgb, px0, py0, rmax, cmax, done, sm);
    if (pb.n_done == BLOCK) break;
    if (pb.dense) {
      // ---- large splats: every pixel walks the records that touch its warp's rows, front to back
      pool_visit_lists(sm, pb.nc);
      for (int wi = 0; wi < BLOCK / 32; ++wi) {
        uint32_t bits = sm.vlist[wrp][wi];
        if (bits && __all_sync(0xffffffffu, done)) break;
        while (bits) {
          const int ci = wi * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          ST3R_EMU_COUNT(13);                               // dense visits (x 32 lanes)
          if (done) continue;
          const float4 ra = sm.recA[ci], rb = sm.recB[ci];
          const float dx = ra.x - px, dy = ra.y - py;
          const float sigma = blend_sigma(rb.x, rb.y, rb.z, dx, dy);
          const float alpha = fminf(ALPHA_MAX, ra.z * exp_neg(sigma));
          if (sigma < 0.f || alpha < ALPHA_MIN) continue;
          const float nT = T * (1.0f - alpha);
          if (nT <= T_MIN) { done = true; continue; }
          const float wgt = alpha * T;
          const float4 col = sm.col[ci];
          pr += col.x * wgt; pg += col.y * wgt; pb_ += col.z * wgt;
          cur = pos + ((sm.recC[ci].z >> 5) & 255);
          T = nT;
          ++blends;
        }
      }
    } else {
      // ---- phase A: alpha of every box pixel, contribution bits
      int s, s_end, ci;
      if (pool_chunk(pb.S, sm.cbase, s, s_end, ci)) {
        PoolWalk wk;
        wk.enter(sm, ci, s, true);
        for (;;) {
          const float dx = wk.x - wk.pxc, dy = wk.y - wk.pyr;
          const float sigma = blend_sigma(wk.qa, wk.qb, wk.qc, dx, dy);
          const float alpha = fminf(ALPHA_MAX, wk.opac * exp_neg(sigma));
          ST3R_EMU_COUNT(10);                               // box pixels tested
          if (!(sigma < 0.f || alpha < ALPHA_MIN)) {
            sm.poolA[s] = alpha;
            atomicOr(&sm.mask[wk.t >> 5][wk.pix], 1u << (wk.t & 31));
          }
          if (++s == s_end) break;
          if (s == wk.run_end) wk.enter(sm, ++ci, s, false);
          else wk.step();
        }
      }
      __syncthreads();
      // ---- phase B: the pixel's own contributing Gaussians, front to back
      if (!done) {
        uint32_t occ = 0;                     // words of this pixel's mask that hold a bit
#pragma unroll
        for (int w = 0; w < PW; ++w) occ |= (sm.mask[w][tr] != 0u ? 1u : 0u) << w;
        uint32_t bits = 0;
        int w = 0;
        for (;;) {
          if (bits == 0u) {
            if (occ == 0u) break;
            w = __ffs(occ) - 1;
            occ &= occ - 1;
            bits = sm.mask[w][tr];
          }
          const int t = w * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          const int2 bx = sm.boxB[t];
          const float alpha = sm.poolA[bx.x + myrow * (bx.y & 31) + mycol];
          const float nT = T * (1.0f - alpha);
          if (nT <= T_MIN) { done = true; break; }
          const float wgt = alpha * T;
          const float4 col = sm.col[bx.y >> 8];
          pr += col.x * wgt; pg += col.y * wgt; pb_ += col.z * wgt;
          cur = pos + t;
          T = nT;
          ++blends;
        }
      }
    }
    pos += pb.n;
    cand = pool_next_cand<false>(cand, pb.n);
    __syncthreads();          // the next stage overwrites the records before its first barrier
  }
  if (inside) {
    const size_t p = ((size_t)c * H + i) * W + j;
    render[3 * p] = pr; render[3 * p + 1] = pg; render[3 * p + 2] = pb_;
    alphas[p] = 1.0f - T;
    last_ids[p] = cur;
  }
  if (n_blend) {
    for (int off = 16; off; off >>= 1) blends += __shfl_xor_sync(0xffffffffu, blends, off);
    if (lane == 0 && blends) atomicAdd(n_blend, (unsigned long long)blends);
  }
}

// Sums of one Gaussian over a set of pixels -> gradient contribution.  m: sum u, u dx, u dy, u dx^2, u dx dy, u dy^2
// (u = vis dL/dalpha) and sum alpha T v_rgb.  With v_sigma = -opac u:  d/dxy = v_sigma Q d,
// d/dconic = v_sigma (dx^2 / 2, dx dy, dy^2 / 2),  d/dopac = u.
__device__ __forceinline__ void pool_emit(const float* m, float opac, float qa, float qb, float qc, int e,
                                          float4* v_geomA, float4* v_geomB, float4* v_rgb) {
  const float no = -opac;
  atomicAdd(v_geomA + e, make_float4(no * (qa * m[1] + qb * m[2]), no * (qb * m[1] + qc * m[2]), m[0], 0.f));
  atomicAdd(v_geomB + e, make_float4(0.5f * no * m[3], no * m[4], 0.5f * no * m[5], 0.f));
  atomicAdd(v_rgb + e, make_float4(m[6], m[7], m[8], 0.f));
}

__global__ void __launch_bounds__(BLOCK, ST3R_POOL_MINB_BWD)
raster_bwd_pool_kernel(const int32_t* __restrict__ offsets, const int32_t* __restrict__ n_isect,
                       const uint32_t* __restrict__ flatten, const float4* __restrict__ geomA,
                       const float4* __restrict__ geomB, const float4* __restrict__ rgb, int C, int W, int H, int tile_w,
                       int tile_h, const float* __restrict__ alphas, const int32_t* __restrict__ last_ids,
                       const float* __restrict__ v_render, const float* __restrict__ v_alphas,
                       float4* __restrict__ v_geomA, float4* __restrict__ v_geomB, float4* __restrict__ v_rgb) {
  POOL_SMEM(true);
  const int c = blockIdx.y, tile = blockIdx.x;
  const int tyi = tile / tile_w, txi = tile - tyi * tile_w;
  const int tr = threadIdx.x, lane = tr & 31, wrp = tr >> 5;
  const int myrow = tr >> 4, mycol = tr & 15;
  const int i = tyi * TILE + myrow, j = txi * TILE + mycol;
  const bool inside = i < H && j < W;
  const float px0 = (float)(txi * TILE) + 0.5f, py0 = (float)(tyi * TILE) + 0.5f;
  const float px = (float)j + 0.5f, py = (float)i + 0.5f;
  const int rmax = min(TILE, H - tyi * TILE) - 1, cmax = min(TILE, W - txi * TILE) - 1;
  const TileRange rg = tile_range(offsets, n_isect, c * tile_w * tile_h + tile, C * tile_w * tile_h);
  if (rg.hi <= rg.lo) return;
  const size_t p = ((size_t)c * H + min(i, H - 1)) * W + min(j, W - 1);
  const float T_final = 1.0f - alphas[p];
  float T = T_final;
  float behind_v = 0.f;                      // (colour behind the current Gaussian) . (upstream colour gradient)
  const int bin_final = inside ? last_ids[p] : -1;
  float vr = 0.f, vg = 0.f, vb = 0.f, va = 0.f;
  if (inside) {
    vr = v_render[3 * p]; vg = v_render[3 * p + 1]; vb = v_render[3 * p + 2];
    va = v_alphas ? v_alphas[p] : 0.f;
  }
  sm.v[tr] = make_float4(vr, vg, vb, 0.f);
  const float tf_va = T_final * va;
  // nothing behind the last Gaussian any pixel of the tile blended takes part
  int tile_last = __reduce_max_sync(0xffffffffu, bin_final);
  if (lane == 0) sm.scan[wrp] = tile_last;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < BLOCK / 32; ++k) tile_last = max(tile_last, sm.scan[k]);
  __syncthreads();                           // scan[] is reused by the first pool_stage
  int hi = min(rg.hi - 1, tile_last);        // slot t of a batch holds sorted position hi - t: ascending t = back to front
  constexpr int PG = PoolCfg<true>::PG, PW = PoolCfg<true>::PW;
  int cand = PG;
  while (hi >= rg.lo) {
    const int idx = hi - tr;
    const bool have = tr < cand && idx >= rg.lo;
    const uint32_t e = have ? flatten[idx] : 0u;
    const PoolBatch pb = pool_stage<true>(have, e, geomA, geomB, rgb, px0, py0, rmax, cmax, 0, sm);
    if (pb.dense) {
      // ---- large splats: every pixel walks the records back to front; nine-value warp reduction per contributing
      // visit into shared-memory sums, one vector atomic triple per record at the end
      float (*acc)[9] = reinterpret_cast<float (*)[9]>(sm.poolA);
      if (tr < pb.nc) {
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[tr][k] = 0.f;
      }
      pool_visit_lists(sm, pb.nc);          // (its barrier also publishes the zeroed sums)
      for (int wi = 0; wi < BLOCK / 32; ++wi) {
        uint32_t bits = sm.vlist[wrp][wi];
        while (bits) {
          const int ci = wi * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          const float4 ra = sm.recA[ci], rb = sm.recB[ci];
          bool valid = hi - ((sm.recC[ci].z >> 5) & 255) <= bin_final;
          float alpha = 0.f, vis = 0.f;
          const float dx = ra.x - px, dy = ra.y - py;
          if (valid) {
            const float sigma = blend_sigma(rb.x, rb.y, rb.z, dx, dy);
            vis = exp_neg(sigma);
            alpha = fminf(ALPHA_MAX, ra.z * vis);
            if (sigma < 0.f || alpha < ALPHA_MIN) valid = false;
          }
          if (!__any_sync(0xffffffffu, valid)) continue;
          ST3R_EMU_COUNT(14);                 // dense backward visits with a contributing lane
          float fac = 0.f, u = 0.f;
          if (valid) {
            const float ra_ = fast_rcp(1.0f - alpha);
            T *= ra_;
            fac = alpha * T;
            const float4 col = sm.col[ci];
            const float cv = col.x * vr + col.y * vg + col.z * vb;
            const float v_alpha = T * cv + ra_ * (tf_va - behind_v);
            behind_v += fac * cv;
            if (ra.z * vis <= ALPHA_MAX) u = vis * v_alpha;
          }
          const float ux = u * dx, uy = u * dy;
          const float g[9] = {u, ux, uy, ux * dx, ux * dy, uy * dy, fac * vr, fac * vg, fac * vb};
          butterfly9_to_shared(g, acc[ci]);
        }
      }
      __syncthreads();
      if (tr < pb.nc) {
        const float* m = acc[tr];
        if (m[0] != 0.f || m[1] != 0.f || m[2] != 0.f || m[3] != 0.f || m[4] != 0.f || m[5] != 0.f || m[6] != 0.f ||
            m[7] != 0.f || m[8] != 0.f) {
          const float4 ra = sm.recA[tr], rb = sm.recB[tr];
          pool_emit(m, ra.z, rb.x, rb.y, rb.z, sm.recC[tr].w, v_geomA, v_geomB, v_rgb);
        }
      }
    } else {
      int s0, s_end, ci0;
      const bool work = pool_chunk(pb.S, sm.cbase, s0, s_end, ci0);
      // ---- phase A
      if (work) {
        int s = s0, ci = ci0;
        PoolWalk wk;
        wk.enter(sm, ci, s, true);
        for (;;) {
          const float dx = wk.x - wk.pxc, dy = wk.y - wk.pyr;
          const float sigma = blend_sigma(wk.qa, wk.qb, wk.qc, dx, dy);
          const float vis = exp_neg(sigma);
          const float alpha = fminf(ALPHA_MAX, wk.opac * vis);
          float pa = 0.f, pv = 0.f;           // non-contributing slots read as "nothing" in phase C
          if (!(sigma < 0.f || alpha < ALPHA_MIN)) {
            pa = alpha;
            pv = (wk.opac * vis <= ALPHA_MAX) ? vis : 0.f;
            atomicOr(&sm.mask[wk.t >> 5][wk.pix], 1u << (wk.t & 31));
          }
          sm.poolA[s] = pa;
          sm.poolV[s] = pv;
          if (++s == s_end) break;
          if (s == wk.run_end) wk.enter(sm, ++ci, s, false);
          else wk.step();
        }
      }
      __syncthreads();
      // ---- phase B: per-pixel recurrence over the pixel's own contributing Gaussians, back to front
      {
        uint32_t occ = 0;
#pragma unroll
        for (int w = 0; w < PW; ++w) occ |= (sm.mask[w][tr] != 0u ? 1u : 0u) << w;
        uint32_t bits = 0;
        int w = 0;
        for (;;) {
          if (bits == 0u) {
            if (occ == 0u) break;
            w = __ffs(occ) - 1;
            occ &= occ - 1;
            bits = sm.mask[w][tr];
          }
          const int t = w * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          const int2 bx = sm.boxB[t];
          const int slot = bx.x + myrow * (bx.y & 31) + mycol;
          float fac = 0.f, u = 0.f;           // a Gaussian behind this pixel's last blended one contributes nothing
          if (hi - t <= bin_final) {
            const float alpha = sm.poolA[slot];
            const float ra = fast_rcp(1.0f - alpha);
            T *= ra;
            fac = alpha * T;
            const float4 col = sm.col[bx.y >> 8];
            const float cv = col.x * vr + col.y * vg + col.z * vb;
            const float v_alpha = T * cv + ra * (tf_va - behind_v);
            behind_v += fac * cv;
            u = sm.poolV[slot] * v_alpha;
            ST3R_EMU_COUNT(11);               // contributing (pixel, Gaussian) pairs
          }
          sm.poolA[slot] = fac;
          sm.poolV[slot] = u;
        }
      }
      __syncthreads();
      // ---- phase C: the chunk's slots summed per Gaussian run
      if (work) {
        int s = s0, ci = ci0;
        PoolWalk wk;
        wk.enter(sm, ci, s, true);
        float m[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        bool any = false;
        for (;;) {
          const float fac = sm.poolA[s], u = sm.poolV[s];
          if (fac != 0.f || u != 0.f) {
            const float dx = wk.x - wk.pxc, dy = wk.y - wk.pyr;
            const float4 v = sm.v[wk.pix];
            const float ux = u * dx, uy = u * dy;
            m[0] += u; m[1] += ux; m[2] += uy;
            m[3] += ux * dx; m[4] += ux * dy; m[5] += uy * dy;
            m[6] += fac * v.x; m[7] += fac * v.y; m[8] += fac * v.z;
            any = true;
          }
          ++s;
          if (s == s_end || s == wk.run_end) {
            if (any) {
              pool_emit(m, wk.opac, wk.qa, wk.qb, wk.qc, wk.e, v_geomA, v_geomB, v_rgb);
#pragma unroll
              for (int k = 0; k < 9; ++k) m[k] = 0.f;
              any = false;
            }
            if (s == s_end) break;
            wk.enter(sm, ++ci, s, false);
          } else {
            wk.step();
          }
        }
      }
    }
    hi -= pb.n;
    cand = pool_next_cand<true>(cand, pb.n);
    __syncthreads();                         // the next stage overwrites the records before its first barrier
  }
}

}  // namespace

#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
static int g_raster_variant = 0;   // 0: fragment-pool kernels (default), 1: visit-list kernels (cross-check, A/B timing)

extern "C" {

int st3r_gs_set_raster_variant(int variant) {
  ST3R_CHECK_ARG(variant == 0 || variant == 1, "st3r_gs_set_raster_variant: unknown variant %d", variant);
  g_raster_variant = variant;
  return ST3R_OK;
}

int st3r_gs_raster_fwd(const int32_t* offsets, const int32_t* n_isect, const uint32_t* flatten_ids,
                       const float* geomA, const float* geomB, const float* rgb, int C, int width, int height,
                       int tile_size, float* render, float* alphas, int32_t* last_ids, uint64_t* n_blend,
                       cudaStream_t stream) {
  ST3R_CHECK_ARG(tile_size == TILE, "st3r_gs_raster_fwd: tile_size must be 16 (gsplat default used by Starst3r)");
  ST3R_CHECK_ARG(C >= 0 && width > 0 && height > 0, "st3r_gs_raster_fwd: bad sizes");
  if (C == 0) return ST3R_OK;
  ST3R_CHECK_ARG(offsets && n_isect && geomA && geomB && rgb && render && alphas && last_ids,
                 "st3r_gs_raster_fwd: null pointer");
  const int tile_w = (width + TILE - 1) / TILE, tile_h = (height + TILE - 1) / TILE;
  dim3 grid(tile_w * tile_h, C);
  const float4 *gA = reinterpret_cast<const float4*>(geomA), *gB = reinterpret_cast<const float4*>(geomB),
               *gC = reinterpret_cast<const float4*>(rgb);
  unsigned long long* nb = reinterpret_cast<unsigned long long*>(n_blend);
  if (g_raster_variant == 0) {
    static PerDeviceOnce once;
    if (!once.done()) {
      ST3R_CHECK_CUDA(cudaFuncSetAttribute(raster_fwd_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(PoolSmem<false>)));
      once.mark();
    }
    raster_fwd_pool_kernel<<<grid, BLOCK, sizeof(PoolSmem<false>), stream>>>(
        offsets, n_isect, flatten_ids, gA, gB, gC, C, width, height, tile_w, tile_h, render, alphas, last_ids, nb);
  } else {
    raster_fwd_kernel<<<grid, BLOCK, 0, stream>>>(offsets, n_isect, flatten_ids, gA, gB, gC, C, width, height, tile_w,
                                                  tile_h, render, alphas, last_ids, nb);
  }
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_gs_raster_bwd(const int32_t* offsets, const int32_t* n_isect, const uint32_t* flatten_ids,
                       const float* geomA, const float* geomB, const float* rgb, int C, int width, int height,
                       int tile_size, const float* alphas, const int32_t* last_ids, const float* v_render,
                       const float* v_alphas, float* v_geomA, float* v_geomB, float* v_rgb, cudaStream_t stream) {
  ST3R_CHECK_ARG(tile_size == TILE, "st3r_gs_raster_bwd: tile_size must be 16");
  ST3R_CHECK_ARG(C >= 0 && width > 0 && height > 0, "st3r_gs_raster_bwd: bad sizes");
  if (C == 0) return ST3R_OK;
  ST3R_CHECK_ARG(offsets && n_isect && geomA && geomB && rgb && alphas && last_ids && v_render && v_geomA &&
                     v_geomB && v_rgb,
                 "st3r_gs_raster_bwd: null pointer");
  const int tile_w = (width + TILE - 1) / TILE, tile_h = (height + TILE - 1) / TILE;
  dim3 grid(tile_w * tile_h, C);
  const float4 *gA = reinterpret_cast<const float4*>(geomA), *gB = reinterpret_cast<const float4*>(geomB),
               *gC = reinterpret_cast<const float4*>(rgb);
  float4 *vA = reinterpret_cast<float4*>(v_geomA), *vB = reinterpret_cast<float4*>(v_geomB),
         *vC = reinterpret_cast<float4*>(v_rgb);
  if (g_raster_variant == 0) {
    static PerDeviceOnce once;
    if (!once.done()) {
      ST3R_CHECK_CUDA(cudaFuncSetAttribute(raster_bwd_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(PoolSmem<true>)));
      once.mark();
    }
    raster_bwd_pool_kernel<<<grid, BLOCK, sizeof(PoolSmem<true>), stream>>>(
        offsets, n_isect, flatten_ids, gA, gB, gC, C, width, height, tile_w, tile_h, alphas, last_ids, v_render, v_alphas,
        vA, vB, vC);
  } else {
    raster_bwd_kernel<<<grid, BLOCK, 0, stream>>>(offsets, n_isect, flatten_ids, gA, gB, gC, C, width, height, tile_w,
                                                  tile_h, alphas, last_ids, v_render, v_alphas, vA, vB, vC);
  }
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
}
#endif  // ST3R_HOST_EMU
