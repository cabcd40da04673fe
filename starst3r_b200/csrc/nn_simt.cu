// Exact-fp32 brute-force nearest neighbour by dot product (SIMT path).
//
// Replaces mast3r/mast3r/fast_nn.py:16-70 (bruteforce_reciprocal_nns, dist='dot'):
// scores = A @ B.T in fp32, row arg-max, ties -> lowest DB index.  The score of a
// (query, db) pair is the sequential FMA chain  s = fma(a[k], b[k], s), k = 0..d-1,
// s0 = 0, which is bit-identical to the MKL sgemm the reference runs on CPU for
// K = 24 (checked in oracle/gen_golden.py), so results are bit-exact incl. ties.
// The score matrix is never written: each CTA keeps an 8x8 register tile per
// thread and reduces it into a per-row (score, index) running best; partial
// results of different DB splits meet in a 64-bit atomicMax on a packed key.
#include "common.cuh"
#include "nn.cuh"

namespace {

constexpr int TM = 128;       // query rows per CTA
constexpr int TN = 128;       // DB rows per smem tile
constexpr int NTHREADS = 256; // 16 x 16 threads, 8 x 8 outputs each

#ifdef ST3R_HOST_EMU   // CPU emulator build (tests/host/): the asynchronous copy completes on the spot
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  memset(smem_dst, 0, 16);
  memcpy(smem_dst, gsrc, (size_t)src_bytes);
}
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#define ST3R_DYN_SMEM_F32(name) extern __shared__ __align__(16) float name[]
#endif

// D = descriptor dim (multiple of 4).  LDB = padded smem row stride (floats), chosen
// so that 8 consecutive rows hit 8 distinct 16-byte bank groups (LDB/4 odd).
template <int D>
struct SimtCfg {
  static constexpr int LDB = (D % 8 == 0) ? D + 4 : D;  // D/4 odd -> already conflict-free
  static constexpr int CHUNKS = D / 4;                  // 16-byte chunks per row
  static constexpr size_t SMEM = sizeof(float) * (size_t)(D * TM + 2 * TN * LDB);
};

template <int D>
__global__ void __launch_bounds__(NTHREADS, 2)
nn_simt_kernel(const float* __restrict__ Qsrc, const int32_t* __restrict__ qidx,
               const int32_t* __restrict__ count_ptr, int Mmax,
               const float* __restrict__ DB, int N, int rows_per_split,
               unsigned long long* __restrict__ packed) {
  using Cfg = SimtCfg<D>;
  constexpr int LDB = Cfg::LDB;
  ST3R_DYN_SMEM_F32(smem);
  float* As = smem;                 // [D][TM]  (k-major)
  float* Bs = smem + D * TM;        // [2][TN][LDB]

  const int M = count_ptr ? min(*count_ptr, Mmax) : Mmax;
  const int m0 = blockIdx.y * TM;
  if (m0 >= M) return;
  const int n_begin = blockIdx.x * rows_per_split;
  if (n_begin >= N) return;
  const int n_end = min(N, n_begin + rows_per_split);
  const int ntiles = (n_end - n_begin + TN - 1) / TN;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  auto load_b_tile = [&](int t, int buf) {
    const int n0 = n_begin + t * TN;
    float* dst = Bs + buf * TN * LDB;
    for (int c = tid; c < TN * Cfg::CHUNKS; c += NTHREADS) {
      int r = c / Cfg::CHUNKS, ch = c - r * Cfg::CHUNKS;
      int gr = n0 + r;
      bool ok = gr < n_end;
      const float* src = DB + (size_t)(ok ? gr : n_begin) * D + ch * 4;
      cp_async16(dst + r * LDB + ch * 4, src, ok ? 16 : 0);
    }
  };

  load_b_tile(0, 0);
  cp_async_commit();

  // A tile: gather (optional) + transpose to k-major.
  for (int e = tid; e < TM * (D / 4); e += NTHREADS) {
    int r = e / (D / 4), ch = e - r * (D / 4);
    int gm = m0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gm < M) {
      size_t row = qidx ? (size_t)qidx[gm] : (size_t)gm;
      v = *reinterpret_cast<const float4*>(Qsrc + row * D + ch * 4);
    }
    As[(ch * 4 + 0) * TM + r] = v.x;
    As[(ch * 4 + 1) * TM + r] = v.y;
    As[(ch * 4 + 2) * TM + r] = v.z;
    As[(ch * 4 + 3) * TM + r] = v.w;
  }

  float best_s[8];
  int best_j[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { best_s[i] = -INFINITY; best_j[i] = 0x7fffffff; }

  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) load_b_tile(t + 1, (t + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const float* Bt = Bs + (t & 1) * TN * LDB;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int k0 = 0; k0 < D; k0 += 4) {
      float4 b4[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        b4[j] = *reinterpret_cast<const float4*>(Bt + (tx + 16 * j) * LDB + k0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 a0 = *reinterpret_cast<const float4*>(As + (k0 + kk) * TM + ty * 8);
        float4 a1 = *reinterpret_cast<const float4*>(As + (k0 + kk) * TM + ty * 8 + 4);
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float b = kk == 0 ? b4[j].x : kk == 1 ? b4[j].y : kk == 2 ? b4[j].z : b4[j].w;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(a[i], b, acc[i][j]);
        }
      }
    }

    const int n0 = n_begin + t * TN;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gj = n0 + tx + 16 * j;
      bool ok = gj < n_end;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float s = acc[i][j];
        // strictly greater -> earlier (lower) index wins ties inside this thread
        if (ok && s > best_s[i]) { best_s[i] = s; best_j[i] = gj; }
      }
    }
    __syncthreads();
  }

  // Reduce over the 16 tx lanes that share a query row (lanes differ in bits 0..3).
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s = best_s[i];
    int j = best_j[i];
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
      float so = __shfl_xor_sync(0xffffffffu, s, off);
      int jo = __shfl_xor_sync(0xffffffffu, j, off);
      if (so > s || (so == s && jo < j)) { s = so; j = jo; }
    }
    int gm = m0 + ty * 8 + i;
    if (tx == 0 && gm < M && j != 0x7fffffff) {
      unsigned long long key = nn_pack(s, j);
      atomicMax(packed + gm, key);
    }
  }
}

// Generic (any d) fallback: one warp per query row, lanes stride the DB.
__global__ void nn_simt_generic_kernel(const float* __restrict__ Qsrc, const int32_t* __restrict__ qidx,
                                       const int32_t* __restrict__ count_ptr, int Mmax,
                                       const float* __restrict__ DB, int N, int d,
                                       unsigned long long* __restrict__ packed) {
  const int M = count_ptr ? min(*count_ptr, Mmax) : Mmax;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= M) return;
  size_t row = qidx ? (size_t)qidx[warp] : (size_t)warp;
  const float* q = Qsrc + row * d;
  float bs = -INFINITY;
  int bj = 0x7fffffff;
  for (int j = lane_id(); j < N; j += 32) {
    const float* b = DB + (size_t)j * d;
    float s = 0.f;
    for (int k = 0; k < d; ++k) s = fmaf(q[k], b[k], s);
    if (s > bs) { bs = s; bj = j; }
  }
  for (int off = 1; off < 32; off <<= 1) {
    float so = __shfl_xor_sync(0xffffffffu, bs, off);
    int jo = __shfl_xor_sync(0xffffffffu, bj, off);
    if (so > bs || (so == bs && jo < bj)) { bs = so; bj = jo; }
  }
  if (lane_id() == 0 && bj != 0x7fffffff) atomicMax(packed + warp, nn_pack(bs, bj));
}

template <int D>
int launch_simt(const float* Q, const int32_t* qidx, const int32_t* count_ptr, int Mmax,
                const float* DB, int N, unsigned long long* packed, cudaStream_t stream) {
  using Cfg = SimtCfg<D>;
  static PerDeviceOnce attr_set;
  if (!attr_set.done()) {
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(nn_simt_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM));
    attr_set.mark();
  }
  const int sms = st3r_num_sms();
  const int mtiles = (Mmax + TM - 1) / TM;
  // Enough DB splits that a single query tile still fills every SM twice.
  int ntiles_total = (N + TN - 1) / TN;
  int nsplit = min(ntiles_total, max(1, (2 * sms + mtiles - 1) / mtiles));
  if (mtiles < 8) nsplit = min(ntiles_total, 2 * sms);
  int tiles_per_split = (ntiles_total + nsplit - 1) / nsplit;
  nsplit = (ntiles_total + tiles_per_split - 1) / tiles_per_split;
  dim3 grid(nsplit, mtiles);
  nn_simt_kernel<D><<<grid, NTHREADS, Cfg::SMEM, stream>>>(Q, qidx, count_ptr, Mmax, DB, N,
                                                           tiles_per_split * TN, packed);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

}  // namespace

int nn_simt_launch(const float* Q, const int32_t* qidx, const int32_t* count_ptr, int Mmax,
                   const float* DB, int N, int d, unsigned long long* packed, cudaStream_t stream) {
  if (Mmax <= 0 || N <= 0) return ST3R_OK;
  if (d == 24 && ((uintptr_t)DB % 16 == 0) && ((uintptr_t)Q % 16 == 0))
    return launch_simt<24>(Q, qidx, count_ptr, Mmax, DB, N, packed, stream);
  if (d == 32 && ((uintptr_t)DB % 16 == 0) && ((uintptr_t)Q % 16 == 0))
    return launch_simt<32>(Q, qidx, count_ptr, Mmax, DB, N, packed, stream);
  if (d == 16 && ((uintptr_t)DB % 16 == 0) && ((uintptr_t)Q % 16 == 0))
    return launch_simt<16>(Q, qidx, count_ptr, Mmax, DB, N, packed, stream);
  int warps_per_block = 8;
  int blocks = (Mmax + warps_per_block - 1) / warps_per_block;
  nn_simt_generic_kernel<<<blocks, warps_per_block * 32, 0, stream>>>(Q, qidx, count_ptr, Mmax, DB, N, d, packed);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
