// Device-wide exclusive prefix sum of int32 (tile counts -> intersection offsets; gsplat's
// torch.cumsum over tiles_per_gauss).  Three kernels: per-block reduce,
// single-block scan of the block sums, per-block scan + carry.  HBM-bound: 12 B / element.
#include "common.cuh"
#include "gs.cuh"

namespace {
constexpr int SC_THREADS = 512;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;  // 4096

__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
  int x = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, off);
    if (lane_id() >= off) x += y;
  }
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane_id() == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = lane_id() < nw ? warp_sums[lane_id()] : 0;
    int xs = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, xs, off);
      if (lane_id() >= off) xs += y;
    }
    if (lane_id() < nw) warp_sums[lane_id()] = xs - w;
    if (lane_id() == 31) *total = xs;
  }
  __syncthreads();
  int r = warp_sums[warp] + x - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SC_THREADS) scan_reduce(const int32_t* __restrict__ in, size_t n, int32_t* __restrict__ block_sums) {
  __shared__ int ws[32];
  size_t base = (size_t)blockIdx.x * SC_TILE;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    size_t i = base + (size_t)k * SC_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane_id() == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = threadIdx.x < (SC_THREADS >> 5) ? ws[threadIdx.x] : 0;
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
  }
}

// In-place exclusive scan of block_sums (single CTA), writes the grand total.
__global__ void __launch_bounds__(1024) scan_block_sums(int32_t* __restrict__ block_sums, int nb, int32_t* __restrict__ total_out) {
  __shared__ int ws[32];
  __shared__ int tot;
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nb ? block_sums[i] : 0;
    int ex = block_excl_scan(v, ws, &tot);
    int c = carry;
    if (i < nb) block_sums[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SC_THREADS) scan_final(const int32_t* __restrict__ in, size_t n, const int32_t* __restrict__ block_sums,
                                                        int32_t* __restrict__ out) {
  __shared__ int ws[32];
  __shared__ int tot;
  size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_ITEMS;  // blocked arrangement
  int v[SC_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    size_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  int ex = block_excl_scan(s, ws, &tot) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SC_ITEMS; ++k) {
    size_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
}
}  // namespace

extern "C" {

size_t st3r_scan_ws_bytes(size_t n) { return (((n + SC_TILE - 1) / SC_TILE) + 1) * sizeof(int32_t) + 512; }

int st3r_exclusive_scan_i32(const int32_t* in, int32_t* out, size_t n, int32_t* total_out, void* ws, size_t ws_bytes,
                            cudaStream_t stream) {
  if (n == 0) {
    if (total_out) ST3R_CHECK_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int32_t), stream));
    return ST3R_OK;
  }
  ST3R_CHECK_ARG(in && out && ws && ws_bytes >= st3r_scan_ws_bytes(n), "st3r_exclusive_scan_i32: bad args / workspace");
  int nb = (int)((n + SC_TILE - 1) / SC_TILE);
  int32_t* block_sums = (int32_t*)ws;
  scan_reduce<<<nb, SC_THREADS, 0, stream>>>(in, n, block_sums);
  ST3R_CHECK_LAUNCH();
  scan_block_sums<<<1, 1024, 0, stream>>>(block_sums, nb, total_out);
  ST3R_CHECK_LAUNCH();
  scan_final<<<nb, SC_THREADS, 0, stream>>>(in, n, block_sums, out);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
}
