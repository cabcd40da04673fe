// Tile binning of the projected Gaussians in one entry point: gsplat's isect_tiles + radix SortPairs +
// isect_offset_encode (called inside gsplat.rasterization, starster/gs.py:76-87).
//
// The generic path (st3r_gs_isect -> st3r_radix_sort_pairs -> st3r_gs_offsets) moves every 12-byte (key, value) pair
// through six 8-bit radix passes: 144 B per intersection, 0.47 ms of a 3.6 ms training step at 3.7 M intersections.
// The key is (camera | tile | depth): the high field has only C * tiles values, so this kernel chain does a counting
// sort on it and a shared-memory sort on the rest:
//   1. tile_hist    : per (Gaussian, view) entry, one atomicAdd per touched tile           -> counts[C * tiles]
//   2. exclusive scan of the counts = isect_offsets itself (and the intersection total)
//   3. tile_emit    : each entry claims slots in its tiles' segments (atomic cursor) and stores (depth bits << 32 | entry)
//   4. tile_sort    : one CTA per (camera, tile) sorts its segment by that 64-bit word: ascending depth, ties by entry
//                     id, which is exactly the order the stable LSD radix sort produces (entries are emitted in
//                     ascending id), so isect_ids / flatten_ids / isect_offsets stay bit-identical.
// HBM traffic: 8 B written + 8 B read + 12 B written per intersection (+ 20 B read per entry, twice).
// Segments up to 4096 pairs are sorted in shared memory; longer ones in place in global memory (L2) by the same
// comparator network.  The network is the "flip" form of bitonic sort, whose comparators all point the same way, so
// padding to a power of two is virtual (+inf never moves).
// ST3R_HOST_EMU: test builds that run this file on a CPU SIMT emulator (tests/host/): bin_emu_host.cpp includes the
// kernels only, build_emu_lib.py (ST3R_EMU_WHOLE) compiles the entry point too, with its launches rewritten.
#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
#include "common.cuh"
#include "gs.cuh"
#endif
#ifndef ST3R_HOST_EMU
#define ST3R_DYN_SMEM_I32(name) extern __shared__ int32_t name[]
#endif
#include "bitonic_reg.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SMEM_ELEMS = 4096;

struct TileRect { int x0, x1, y0, y1; };

// Same arithmetic as gs_isect_kernel (gs_project.cu): tile rectangle of a projected Gaussian.
__device__ __forceinline__ TileRect tile_rect(const float4 a, int r, int tile_size, int tile_w, int tile_h) {
  const float ts = (float)tile_size;
  const float txc = a.x / ts, tyc = a.y / ts, tr = (float)r / ts;
  TileRect q;
  q.x0 = min(max(0, (int)floorf(txc - tr)), tile_w); q.x1 = min(max(0, (int)ceilf(txc + tr)), tile_w);
  q.y0 = min(max(0, (int)floorf(tyc - tr)), tile_h); q.y1 = min(max(0, (int)ceilf(tyc + tr)), tile_h);
  return q;
}

// Both passes privatise the per-tile counters of ONE camera in shared memory (global atomics on ~10^3 hot
// addresses ran at 25 G/s: 150-190 us per pass at 3.7 M intersections): a CTA walks ENTRIES_PER_CTA consecutive
// entries of its camera, counts with shared-memory atomics and touches global memory once per non-empty tile.
constexpr int BIN_THREADS = 512;
constexpr int MIN_ENTRIES_PER_CTA = 1024, MAX_ENTRIES_PER_CTA = 16384;   // see st3r_gs_bin_tiles
constexpr int MAX_SMEM_TILES = 10240;      // 40 KB of counters; larger images fall back to global atomics

__global__ void __launch_bounds__(BIN_THREADS)
tile_hist_kernel(const int32_t* __restrict__ radii, const float4* __restrict__ geomA, int N, int tile_size,
                 int tile_w, int tile_h, int32_t* __restrict__ counts, int use_smem, int entries_per_cta) {
  ST3R_DYN_SMEM_I32(s_cnt);
  const int n_tiles = tile_w * tile_h;
  const int c = blockIdx.y;
  int32_t* row = counts + (size_t)c * n_tiles;
  if (use_smem) {
    for (int t = threadIdx.x; t < n_tiles; t += BIN_THREADS) s_cnt[t] = 0;
    __syncthreads();
  }
  int32_t* dst = use_smem ? s_cnt : row;
  const int g0 = blockIdx.x * entries_per_cta, g1 = min(N, g0 + entries_per_cta);
  for (int g = g0 + threadIdx.x; g < g1; g += BIN_THREADS) {
    const size_t e = (size_t)c * N + g;
    const int r = radii[e];
    if (r <= 0) continue;
    const TileRect q = tile_rect(geomA[e], r, tile_size, tile_w, tile_h);
    for (int y = q.y0; y < q.y1; ++y)
      for (int x = q.x0; x < q.x1; ++x) atomicAdd(dst + y * tile_w + x, 1);
  }
  if (use_smem) {
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += BIN_THREADS) {
      const int v = s_cnt[t];
      if (v) atomicAdd(row + t, v);
    }
  }
}

__global__ void __launch_bounds__(BIN_THREADS)
tile_emit_kernel(const int32_t* __restrict__ radii, const float4* __restrict__ geomA, int N, int tile_size,
                 int tile_w, int tile_h, const int32_t* __restrict__ offsets, int32_t* __restrict__ cursor,
                 uint64_t* __restrict__ pairs, int n_cap, int use_smem, int entries_per_cta) {
  ST3R_DYN_SMEM_I32(s_cnt);                // [n_tiles] local counts, then local cursors; [n_tiles] claimed bases
  const int n_tiles = tile_w * tile_h;
  const int c = blockIdx.y;
  const size_t row = (size_t)c * n_tiles;
  int32_t* s_base = s_cnt + n_tiles;
  const int g0 = blockIdx.x * entries_per_cta, g1 = min(N, g0 + entries_per_cta);
  if (use_smem) {
    for (int t = threadIdx.x; t < n_tiles; t += BIN_THREADS) s_cnt[t] = 0;
    __syncthreads();
    for (int g = g0 + threadIdx.x; g < g1; g += BIN_THREADS) {
      const size_t e = (size_t)c * N + g;
      const int r = radii[e];
      if (r <= 0) continue;
      const TileRect q = tile_rect(geomA[e], r, tile_size, tile_w, tile_h);
      for (int y = q.y0; y < q.y1; ++y)
        for (int x = q.x0; x < q.x1; ++x) atomicAdd(s_cnt + y * tile_w + x, 1);
    }
    __syncthreads();
    // claim a contiguous slice of every touched tile's segment, then restart the local counters as cursors
    for (int t = threadIdx.x; t < n_tiles; t += BIN_THREADS) {
      const int v = s_cnt[t];
      s_base[t] = v ? offsets[row + t] + atomicAdd(cursor + row + t, v) : 0;
      s_cnt[t] = 0;
    }
    __syncthreads();
  }
  for (int g = g0 + threadIdx.x; g < g1; g += BIN_THREADS) {
    const size_t e = (size_t)c * N + g;
    const int r = radii[e];
    if (r <= 0) continue;
    const float4 a = geomA[e];
    const TileRect q = tile_rect(a, r, tile_size, tile_w, tile_h);
    const uint64_t word = ((uint64_t)__float_as_uint(a.w) << 32) | (uint64_t)(uint32_t)e;
    for (int y = q.y0; y < q.y1; ++y)
      for (int x = q.x0; x < q.x1; ++x) {
        const int t = y * tile_w + x;
        const int pos = use_smem ? s_base[t] + atomicAdd(s_cnt + t, 1)
                                 : offsets[row + t] + atomicAdd(cursor + row + t, 1);
        if (pos < n_cap) pairs[pos] = word;
      }
  }
}

// Ascending comparator network over buf[0..n) (virtual +inf padding up to the next power of two): the "flip" form
// of bitonic sort.  All index arithmetic is shifts and masks (k = 1 << lk, j = 1 << lj).  Comparators with a span
// below 64 elements never leave a 64-element block, so a warp runs those steps of its blocks back to back with
// __syncwarp only; block-wide barriers remain for the long-span steps (12 instead of 45 for 512 elements).
__device__ __forceinline__ void compare_exchange(uint64_t* buf, int lo, int hi, int n) {
  if (hi < n) {
    const uint64_t a = buf[lo], b = buf[hi];
    if (a > b) { buf[lo] = b; buf[hi] = a; }
  }
}

// steps j = 2^lj_start .. 1 (lj_start <= 4) inside every 64-element block, one warp per block
__device__ __forceinline__ void warp_tail(uint64_t* buf, int n, int P, int lj_start, int nwarps) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  for (int base = warp * 64; base < P && base < n; base += nwarps * 64) {
    for (int lj = lj_start; lj >= 0; --lj) {
      const int lo = base + (((lane >> lj) << (lj + 1)) | (lane & ((1 << lj) - 1)));
      compare_exchange(buf, lo, lo + (1 << lj), n);
      __syncwarp();
    }
  }
}

__device__ __forceinline__ void bitonic_flip_sort(uint64_t* buf, int n, int nthreads) {
  int lp = 0;
  while ((1 << lp) < n) ++lp;
  const int P = 1 << lp, half = P >> 1, nwarps = nthreads >> 5;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  // stages k = 2 .. 64: entirely inside 64-element blocks
  for (int base = warp * 64; base < P && base < n; base += nwarps * 64) {
    for (int lk = 1; lk <= min(lp, 6); ++lk) {
      const int hk = lk - 1;
      const int lo = base + (((lane >> hk) << lk) | (lane & ((1 << hk) - 1)));
      compare_exchange(buf, lo, lo ^ ((1 << lk) - 1), n);
      __syncwarp();
      for (int lj = lk - 2; lj >= 0; --lj) {
        const int l2 = base + (((lane >> lj) << (lj + 1)) | (lane & ((1 << lj) - 1)));
        compare_exchange(buf, l2, l2 + (1 << lj), n);
        __syncwarp();
      }
    }
  }
  __syncthreads();
  for (int lk = 7; lk <= lp; ++lk) {
    const int hk = lk - 1;
    for (int i = threadIdx.x; i < half; i += nthreads) {
      const int lo = ((i >> hk) << lk) | (i & ((1 << hk) - 1));
      compare_exchange(buf, lo, lo ^ ((1 << lk) - 1), n);
    }
    __syncthreads();
    for (int lj = lk - 2; lj >= 5; --lj) {
      for (int i = threadIdx.x; i < half; i += nthreads) {
        const int lo = ((i >> lj) << (lj + 1)) | (i & ((1 << lj) - 1));
        compare_exchange(buf, lo, lo + (1 << lj), n);
      }
      __syncthreads();
    }
    warp_tail(buf, n, P, 4, nwarps);
    __syncthreads();
  }
}

// ---- register-resident sort for segments up to 2048 pairs: bitonic_reg.cuh (256 threads, two exchange buffers) -------
template <int EPT>
__device__ __forceinline__ bool reg_sort_segment(const uint64_t* __restrict__ seg, int n, uint64_t* sx, uint64_t key_hi,
                                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint64_t v[EPT];
  const int i0 = threadIdx.x * EPT;
#pragma unroll
  for (int e = 0; e < EPT; ++e) v[e] = i0 + e < n ? seg[i0 + e] : st3r_sort::SORT_PAD;    // padding sorts to the end
  st3r_sort::reg_bitonic_sort<SORT_THREADS, EPT, 2>(v, sx, n);
#pragma unroll
  for (int e = 0; e < EPT; ++e)
    if (i0 + e < n) {
      keys[i0 + e] = key_hi | (v[e] >> 32);
      vals[i0 + e] = (uint32_t)v[e];
    }
  return true;
}

// The same segments through 32-bit surrogate keys (default).  73 % of the 64-bit network's instructions are the compare +
// select pairs of its comparators (sm_100a has no 64-bit integer min / max); a 32-bit comparator is two instructions and
// a shuffle step moves one register.  Surrogate = the depth bits with their 11 lowest bits replaced by the element's
// position in the unsorted segment: unique, monotone in the depth, and it names the element.  Sorting the surrogates
// orders the segment up to the elements whose depths agree in the 21 leading bits (sign, exponent, 12 mantissa bits: a
// few pairs per tile); the full (depth, entry) words are then gathered in that order into shared memory and odd-even
// transposition passes finish the job - one pass fixes runs of two, the loop ends after the first pass without a swap, and
// a depth distribution that keeps it busy (thousands of equal depths) falls back to the shared-memory network.  The result
// is the ascending order of the full words, i.e. bit-identical to the 64-bit network and to the radix chain.
constexpr int SUR_IDX_BITS = 11;
constexpr int SUR_MAX_PASSES = 12;
static_assert(8 * SORT_THREADS <= (1 << SUR_IDX_BITS), "a segment position must fit into the surrogate's index bits");
static_assert(SMEM_ELEMS >= 2 * 8 * SORT_THREADS, "16 KB of exchange buffers + 16 KB for the repair passes");
template <int EPT>
__device__ __forceinline__ void reg_sort_segment32(const uint64_t* __restrict__ seg, int n, uint64_t* sbuf, uint64_t key_hi,
                                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  constexpr uint32_t IDX_MASK = (1u << SUR_IDX_BITS) - 1u;
  uint32_t k[EPT];
  const int i0 = threadIdx.x * EPT;
#pragma unroll
  for (int e = 0; e < EPT; ++e)
    k[e] = i0 + e < n ? (((uint32_t)(seg[i0 + e] >> 32) & ~IDX_MASK) | (uint32_t)(i0 + e)) : 0xffffffffu;
  st3r_sort::reg_bitonic_sort<SORT_THREADS, EPT, 2, uint32_t>(k, reinterpret_cast<uint32_t*>(sbuf), n);
  uint64_t* s = sbuf + 8 * SORT_THREADS;          // behind the exchange buffers (2 x 2048 x 4 bytes)
#pragma unroll
  for (int e = 0; e < EPT; ++e)
    if (i0 + e < n) s[i0 + e] = seg[k[e] & IDX_MASK];
  __syncthreads();
  for (int pass = 0;; ++pass) {
    int swapped = 0;
#pragma unroll
    for (int parity = 0; parity < 2; ++parity) {
      for (int p = 2 * (int)threadIdx.x + parity; p + 1 < n; p += 2 * SORT_THREADS) {
        const uint64_t a = s[p], b = s[p + 1];
        if (a > b) { s[p] = b; s[p + 1] = a; swapped = 1; }
      }
      __syncthreads();
    }
    if (__syncthreads_count(swapped) == 0) break;
    if (pass + 1 == SUR_MAX_PASSES) {
      bitonic_flip_sort(s, n, SORT_THREADS);
      __syncthreads();
      break;
    }
  }
#pragma unroll
  for (int e = 0; e < EPT; ++e)
    if (i0 + e < n) {
      const uint64_t w = s[i0 + e];
      keys[i0 + e] = key_hi | (w >> 32);
      vals[i0 + e] = (uint32_t)w;
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
tile_sort_kernel(int32_t* offsets, const int32_t* __restrict__ total_ptr, int n_cells, int n_tiles,
                 int tile_n_bits, uint64_t* __restrict__ pairs, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                 int n_cap, int reg_path) {
  __shared__ uint64_t sbuf[SMEM_ELEMS];
  const int cell = blockIdx.x;                       // camera * n_tiles + tile
  const int lo = min(offsets[cell], n_cap);
  const int hi = min(cell + 1 < n_cells ? offsets[cell + 1] : *total_ptr, n_cap);
  const int n = hi - lo;
  // The blend kernels index the lists through `offsets`: keep it inside the capacity (idempotent, so the CTA of the
  // previous cell may read either value).
  if (threadIdx.x == 0 && offsets[cell] > n_cap) offsets[cell] = n_cap;
  if (n <= 0) return;
  const int cam = cell / n_tiles, tile = cell - cam * n_tiles;
  const uint64_t key_hi = ((((uint64_t)cam << tile_n_bits) | (uint64_t)tile) << 32);
  uint64_t* seg = pairs + lo;
  if (reg_path && n <= 8 * SORT_THREADS) {           // CTA-uniform
    if (reg_path == 2) {
      if (n <= SORT_THREADS) reg_sort_segment32<1>(seg, n, sbuf, key_hi, keys + lo, vals + lo);
      else if (n <= 2 * SORT_THREADS) reg_sort_segment32<2>(seg, n, sbuf, key_hi, keys + lo, vals + lo);
      else if (n <= 4 * SORT_THREADS) reg_sort_segment32<4>(seg, n, sbuf, key_hi, keys + lo, vals + lo);
      else reg_sort_segment32<8>(seg, n, sbuf, key_hi, keys + lo, vals + lo);
      return;
    }
    bool done;
    if (n <= SORT_THREADS) done = reg_sort_segment<1>(seg, n, sbuf, key_hi, keys + lo, vals + lo);
    else if (n <= 2 * SORT_THREADS) done = reg_sort_segment<2>(seg, n, sbuf, key_hi, keys + lo, vals + lo);
    else if (n <= 4 * SORT_THREADS) done = reg_sort_segment<4>(seg, n, sbuf, key_hi, keys + lo, vals + lo);
    else done = reg_sort_segment<8>(seg, n, sbuf, key_hi, keys + lo, vals + lo);
    if (done) return;
  }
  uint64_t* buf = seg;
  if (n <= SMEM_ELEMS) {
    for (int i = threadIdx.x; i < n; i += SORT_THREADS) sbuf[i] = seg[i];
    __syncthreads();
    buf = sbuf;
  }
  if (n > 1) bitonic_flip_sort(buf, n, SORT_THREADS);
  for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
    const uint64_t w = buf[i];
    keys[lo + i] = key_hi | (w >> 32);
    vals[lo + i] = (uint32_t)w;
  }
}

#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
struct BinWs {
  int32_t* counts; int32_t* cursor; uint64_t* pairs; void* scan_ws; size_t scan_bytes;
};

size_t carve_bin(BinWs* w, void* ws, size_t ws_bytes, int n_cells, int n_cap, bool dry) {
  WsAlloc a(dry ? (void*)0 : ws, dry ? (size_t)-1 : ws_bytes);
  BinWs t;
  t.counts = a.take<int32_t>(2 * (size_t)n_cells);     // counts | cursor, zeroed with one memset
  t.cursor = t.counts + n_cells;
  t.pairs = a.take<uint64_t>((size_t)(n_cap > 0 ? n_cap : 1));
  t.scan_bytes = st3r_scan_ws_bytes((size_t)n_cells);
  t.scan_ws = a.take<char>(t.scan_bytes);
  if (w) *w = t;
  return a.off + 256;
}
#endif  // ST3R_HOST_EMU

}  // namespace

#if !defined(ST3R_HOST_EMU) || defined(ST3R_EMU_WHOLE)
// 2 (default): segments up to 2048 pairs are sorted in registers through 32-bit surrogate keys + repair passes
// (reg_sort_segment32); 1: in registers as 64-bit words (reg_bitonic_sort); 0: always the shared-memory / in-place network
// (the first implementation).  1 and 0 stay as cross-checks: tests/test_gs_gpu.py runs all three.
#ifndef ST3R_BIN_DEFAULT
#define ST3R_BIN_DEFAULT 2
#endif
static int g_bin_reg_sort = ST3R_BIN_DEFAULT;

extern "C" {

int st3r_gs_bin_set_variant(int variant) {
  ST3R_CHECK_ARG(variant >= 0 && variant <= 2, "st3r_gs_bin_set_variant: unknown variant %d", variant);
  g_bin_reg_sort = variant;
  return ST3R_OK;
}

size_t st3r_gs_bin_ws_bytes(int C, int width, int height, int tile_size, int n_cap) {
  if (C <= 0 || width <= 0 || height <= 0 || tile_size <= 0) return 256;
  const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
  return carve_bin(nullptr, nullptr, 0, C * tile_w * tile_h, n_cap, true);
}

int st3r_gs_bin_tiles(const int32_t* radii, const float* geomA, int N, int C, int width, int height, int tile_size,
                      int32_t* offsets, int32_t* n_isect_out, uint64_t* keys, uint32_t* vals, int n_cap, void* ws,
                      size_t ws_bytes, cudaStream_t stream) {
  ST3R_CHECK_ARG(N >= 0 && C >= 0 && width > 0 && height > 0 && tile_size > 0 && n_cap >= 0, "st3r_gs_bin_tiles: bad sizes");
  ST3R_CHECK_ARG(offsets && n_isect_out, "st3r_gs_bin_tiles: null pointer");
  const int tile_w = (width + tile_size - 1) / tile_size, tile_h = (height + tile_size - 1) / tile_size;
  const int n_tiles = tile_w * tile_h, n_cells = C * n_tiles;
  if (n_cells == 0 || N == 0) {
    ST3R_CHECK_CUDA(cudaMemsetAsync(n_isect_out, 0, sizeof(int32_t), stream));
    if (n_cells > 0) ST3R_CHECK_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * (size_t)n_cells, stream));
    return ST3R_OK;
  }
  ST3R_CHECK_ARG(radii && geomA && ws && (n_cap == 0 || (keys && vals)), "st3r_gs_bin_tiles: null pointer");
  BinWs w;
  const size_t need = carve_bin(&w, ws, ws_bytes, n_cells, n_cap, false);
  if (ws_bytes < need) {
    st3r_set_error("st3r_gs_bin_tiles: workspace too small (%zu < %zu)", ws_bytes, need);
    return ST3R_ERR_WORKSPACE;
  }
  ST3R_CHECK_CUDA(cudaMemsetAsync(w.counts, 0, sizeof(int32_t) * 2 * (size_t)n_cells, stream));
  const float4* gA = reinterpret_cast<const float4*>(geomA);
  const int use_smem = n_tiles <= MAX_SMEM_TILES ? 1 : 0;
  // Entries per CTA of the two counting passes: the smallest power of two that still amortises the CTA's fixed cost
  // (zeroing and flushing one shared-memory counter per tile).  Measured on the headline frame (8 x 200 k entries, 1024
  // tiles, B200): 16384 / 8192 / 4096 / 2048 / 1024 entries -> 0.333 / 0.271 / 0.237 / 0.223 / 0.211 ms for the whole
  // binning (a grid of 200 CTAs of 512 threads left the SMs mostly idle).
  int epc = MIN_ENTRIES_PER_CTA;
  while (epc < n_tiles && epc < MAX_ENTRIES_PER_CTA) epc *= 2;
  const dim3 grid((N + epc - 1) / epc, C);
  const size_t sm_hist = use_smem ? sizeof(int32_t) * (size_t)n_tiles : 0, sm_emit = 2 * sm_hist;
  static PerDeviceOnce attr_set;
  if (!attr_set.done()) {
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(tile_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(2 * sizeof(int32_t) * MAX_SMEM_TILES)));
    attr_set.mark();
  }
  tile_hist_kernel<<<grid, BIN_THREADS, sm_hist, stream>>>(radii, gA, N, tile_size, tile_w, tile_h, w.counts, use_smem, epc);
  ST3R_CHECK_LAUNCH();
  int rc = st3r_exclusive_scan_i32(w.counts, offsets, (size_t)n_cells, n_isect_out, w.scan_ws, w.scan_bytes, stream);
  if (rc != ST3R_OK) return rc;
  if (n_cap == 0) return ST3R_OK;
  tile_emit_kernel<<<grid, BIN_THREADS, sm_emit, stream>>>(radii, gA, N, tile_size, tile_w, tile_h, offsets, w.cursor,
                                                          w.pairs, n_cap, use_smem, epc);
  ST3R_CHECK_LAUNCH();
  tile_sort_kernel<<<n_cells, SORT_THREADS, 0, stream>>>(offsets, n_isect_out, n_cells, n_tiles, gs_tile_bits(n_tiles),
                                                         w.pairs, keys, vals, n_cap, g_bin_reg_sort);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
}
#endif  // ST3R_HOST_EMU
