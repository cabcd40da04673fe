// Stable LSD radix sort of (uint64 key, uint32 value) pairs, 8 bits per pass.
//
// Stands in for the CUB DeviceRadixSort::SortPairs call gsplat makes on
// (isect_id, flatten_id) and for np.unique's sort in
// mast3r/mast3r/fast_nn.py:92.  The element count lives on the device
// (n_ptr) so callers never synchronise; grids are sized for n_cap.
//
// Per pass: (1) upsweep: per-tile digit counts -> table[digit][tile];
// (2) rowscan: exclusive scan of each digit row + row totals;
// (3) downsweep: stable rank inside the tile (warp match + per-warp digit
// counters, warps own contiguous segments) and scatter.
#include "common.cuh"
#include "radix_sort.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                               // keys per lane
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;            // 4096 keys per CTA
constexpr int RS_SEG = 32 * RS_ITEMS;                     // keys per warp segment

__device__ __forceinline__ int rs_n(const int* n_ptr, int n_cap) {
  return n_ptr ? min(*n_ptr, n_cap) : n_cap;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_upsweep(const uint64_t* __restrict__ keys, const int* __restrict__ n_ptr, int n_cap, int shift,
           int num_tiles, uint32_t* __restrict__ table) {
  __shared__ uint32_t hist[RS_WARPS][256];
  const int n = rs_n(n_ptr, n_cap);
  const int tile = blockIdx.x;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&hist[0][0])[i] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const size_t base = (size_t)tile * RS_TILE;
  if (base < (size_t)n) {
    for (int it = 0; it < RS_ITEMS; ++it) {
      size_t i = base + (size_t)it * RS_THREADS + threadIdx.x;
      if (i < (size_t)n) {
        uint32_t d = (uint32_t)(keys[i] >> shift) & 0xffu;
        atomicAdd(&hist[warp][d], 1u);
      }
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < 256; d += RS_THREADS) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) s += hist[w][d];
    table[(size_t)d * num_tiles + tile] = s;
  }
}

// One CTA per digit: exclusive scan of table[d][0..num_tiles) in place; totals[d] = row sum.
__global__ void __launch_bounds__(1024)
rs_rowscan(uint32_t* __restrict__ table, int num_tiles, uint32_t* __restrict__ totals) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const int d = blockIdx.x;
  uint32_t* row = table + (size_t)d * num_tiles;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < num_tiles; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < num_tiles ? row[i] : 0;
    uint32_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane_id() >= off) x += y;
    }
    if (lane_id() == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = warp_sums[threadIdx.x];
      uint32_t xs = w;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, xs, off);
        if (lane_id() >= off) xs += y;
      }
      warp_sums[threadIdx.x] = xs - w;  // exclusive
    }
    __syncthreads();
    uint32_t carry = carry_s;
    uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + x - v;
    if (i < num_tiles) row[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[d] = carry_s;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_downsweep(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
             uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
             const int* __restrict__ n_ptr, int n_cap, int shift, int num_tiles,
             const uint32_t* __restrict__ table, const uint32_t* __restrict__ totals) {
  __shared__ uint32_t wcount[RS_WARPS][256];  // per-warp digit counters -> warp bases
  __shared__ uint32_t dbase[256];             // global base of this tile's run per digit
  const int n = rs_n(n_ptr, n_cap);
  const int tile = blockIdx.x;
  const size_t base = (size_t)tile * RS_TILE;
  if (base >= (size_t)n) return;
  const int warp = threadIdx.x >> 5, lane = lane_id();

  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcount[0][0])[i] = 0;
  // digit base = sum of totals of lower digits + this tile's scanned offset
  {
    // 256 threads: exclusive scan of totals over digits
    __shared__ uint32_t wsum[RS_WARPS];
    uint32_t v = totals[threadIdx.x];
    uint32_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t pre = 0;
    for (int w = 0; w < warp; ++w) pre += wsum[w];
    dbase[threadIdx.x] = pre + x - v + table[(size_t)threadIdx.x * num_tiles + tile];
  }
  __syncthreads();

  uint64_t k[RS_ITEMS];
  uint32_t rank[RS_ITEMS];
  const size_t seg = base + (size_t)warp * RS_SEG;
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    size_t i = seg + (size_t)it * 32 + lane;
    bool ok = i < (size_t)n;
    k[it] = ok ? keys_in[i] : 0;
    uint32_t d = ok ? ((uint32_t)(k[it] >> shift) & 0xffu) : 0x100u;  // 0x100: inactive group
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t r = __popc(peers & lt);
    uint32_t b = 0;
    int leader = __ffs(peers) - 1;
    if (ok && lane == leader) {
      b = wcount[warp][d];
      wcount[warp][d] = b + __popc(peers);
    }
    b = __shfl_sync(0xffffffffu, b, leader);
    rank[it] = b + r;
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan over warps for each digit
  {
    int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c = wcount[w][d];
      wcount[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    size_t i = seg + (size_t)it * 32 + lane;
    if (i < (size_t)n) {
      uint32_t d = (uint32_t)(k[it] >> shift) & 0xffu;
      size_t dst = (size_t)dbase[d] + wcount[warp][d] + rank[it];
      keys_out[dst] = k[it];
      if (vals_in) vals_out[dst] = vals_in[i];
    }
  }
}

__global__ void rs_copy(const uint64_t* __restrict__ ksrc, const uint32_t* __restrict__ vsrc,
                        uint64_t* __restrict__ kdst, uint32_t* __restrict__ vdst,
                        const int* __restrict__ n_ptr, int n_cap) {
  const int n = rs_n(n_ptr, n_cap);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n;
       i += (size_t)gridDim.x * blockDim.x) {
    kdst[i] = ksrc[i];
    if (vsrc) vdst[i] = vsrc[i];
  }
}

}  // namespace

size_t radix_sort_ws_bytes(int n_cap) {
  size_t tiles = ((size_t)n_cap + RS_TILE - 1) / RS_TILE;
  if (tiles == 0) tiles = 1;
  return st3r_align_up(tiles * 256 * sizeof(uint32_t), 256) + 256 * sizeof(uint32_t) + 512;
}

int radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                     const int* n_ptr, int n_cap, int begin_bit, int end_bit, void* ws,
                     size_t ws_bytes, cudaStream_t stream) {
  if (n_cap <= 0 || end_bit <= begin_bit) return ST3R_OK;
  ST3R_CHECK_ARG(ws_bytes >= radix_sort_ws_bytes(n_cap), "radix_sort: workspace too small");
  ST3R_CHECK_ARG((vals == nullptr) == (vals_alt == nullptr), "radix_sort: vals/vals_alt mismatch");
  const int num_tiles = (int)(((size_t)n_cap + RS_TILE - 1) / RS_TILE);
  WsAlloc a(ws, ws_bytes);
  uint32_t* table = a.take<uint32_t>((size_t)num_tiles * 256);
  uint32_t* totals = a.take<uint32_t>(256);

  uint64_t* kin = keys;
  uint32_t* vin = vals;
  uint64_t* kout = keys_alt;
  uint32_t* vout = vals_alt;
  for (int shift = begin_bit; shift < end_bit; shift += 8) {
    rs_upsweep<<<num_tiles, RS_THREADS, 0, stream>>>(kin, n_ptr, n_cap, shift, num_tiles, table);
    ST3R_CHECK_LAUNCH();
    rs_rowscan<<<256, 1024, 0, stream>>>(table, num_tiles, totals);
    ST3R_CHECK_LAUNCH();
    rs_downsweep<<<num_tiles, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, n_ptr, n_cap, shift,
                                                       num_tiles, table, totals);
    ST3R_CHECK_LAUNCH();
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  if (kin != keys) {
    int blocks = min(num_tiles * 4, 148 * 8);
    rs_copy<<<blocks, 256, 0, stream>>>(kin, vin, keys, vals, n_ptr, n_cap);
    ST3R_CHECK_LAUNCH();
  }
  return ST3R_OK;
}

#include "../../include/starst3r_b200.h"
extern "C" {
size_t st3r_radix_sort_ws_bytes(int n_cap) { return radix_sort_ws_bytes(n_cap); }
int st3r_radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                          const int32_t* n_ptr, int n_cap, int begin_bit, int end_bit, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
  ST3R_CHECK_ARG(n_cap >= 0 && begin_bit >= 0 && end_bit <= 64, "st3r_radix_sort_pairs: bad args");
  if (n_cap == 0) return ST3R_OK;
  ST3R_CHECK_ARG(keys && keys_alt && ws, "st3r_radix_sort_pairs: null pointer");
  return radix_sort_pairs(keys, vals, keys_alt, vals_alt, n_ptr, n_cap, begin_bit, end_bit, ws, ws_bytes, stream);
}
}
