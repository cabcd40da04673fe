// Device-resident seeded reciprocal nearest-neighbour search + correspondence merge.
//
// Replaces the host/numpy control flow of
//   mast3r/mast3r/fast_nn.py:109-188  (fast_reciprocal_NNs, pixel_tol=0, ret_basin=False)
//   mast3r/mast3r/fast_nn.py:87-106   (merge_corres)
//   mast3r/mast3r/cloud_opt/sparse_ga.py:595-630 (extract_correspondences)
// The reference does 2 D2H + 2 H2D round trips per iteration; here the active
// (not-yet-converged) seed list, its length and the convergence flags stay in
// HBM and every kernel of the fixed launch chain reads the live count.
#include "common.cuh"
#include "nn.cuh"
#include "radix_sort.cuh"
#include "recip.cuh"
#include "bitonic_reg.cuh"
#ifndef ST3R_HOST_EMU
#define ST3R_DYN_SMEM_U64(name) extern __shared__ __align__(16) uint64_t name[]
#endif

namespace {

constexpr int TPB = 256;
static inline int nblk(int n) { return (n + TPB - 1) / TPB; }

struct RecipState {
  int32_t* xy[2];      // xy[0] = xy1 (index into map 1), xy[1] = xy2
  int32_t* old[2];
  uint8_t* notyet;
  int32_t* active[2];  // ping-pong active slot lists
  int32_t* counts;     // counts[h] = active length before half-iteration h
  int32_t* qidx[2];    // query row indices of the current / next half-iteration (ping-pong)
  unsigned long long* packed[2];
};

constexpr int RB = NN_MAX_BATCH;
// Up to four seeded searches advanced in lock-step (blockIdx.y selects the search).
struct RecipBatch {
  RecipState st[RB];
  int nseed[RB], nx[RB], W1[RB], swap[RB], set[RB];
  const int32_t* seeds[RB];
  int S, n;
};

__global__ void recip_init(RecipBatch b, int ncounts) {
  const RecipState& st = b.st[blockIdx.y];
  const int n = b.nseed[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncounts) st.counts[i] = (i == 0) ? n : 0;
  if (i >= n) return;
  int32_t s;
  if (b.seeds[blockIdx.y]) {
    s = b.seeds[blockIdx.y][i];
  } else {
    const int nx = b.nx[blockIdx.y], S = b.S;
    int y = S / 2 + (i / nx) * S, x = S / 2 + (i % nx) * S;
    s = x + b.W1[blockIdx.y] * y;
  }
  st.xy[0][i] = s; st.old[0][i] = s;
  st.xy[1][i] = -1; st.old[1][i] = -1;
  st.notyet[i] = 1;
  st.active[0][i] = i;
  st.qidx[0][i] = s;          // first half-iteration queries P1[seed]
  st.packed[0][i] = 0ull;
}

// Scatter NN results, drop converged slots, build the next active list and - because the next half-iteration
// queries exactly the rows just found (P_dst[xy_dst[slot]]) - its query index list and cleared result slots.
__global__ void recip_update(RecipBatch b, int h, int dst) {
  const RecipState& st = b.st[blockIdx.y];
  int32_t* __restrict__ xy_dst = st.xy[dst];
  int32_t* __restrict__ old_dst = st.old[dst];
  const int32_t* __restrict__ active = st.active[h & 1];
  int32_t* __restrict__ active_next = st.active[(h + 1) & 1];
  const unsigned long long* __restrict__ packed = st.packed[h & 1];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool in = i < st.counts[h];
  bool keep = false;
  int s = 0, j = 0;
  if (in) {
    s = active[i];
    j = nn_unpack_idx(packed[i]);
    xy_dst[s] = j;
    keep = old_dst[s] != j;
    old_dst[s] = j;
    if (!keep) st.notyet[s] = 0;
  }
  unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m) {
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane_id() == leader) base = atomicAdd(st.counts + h + 1, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (keep) {
      int pos = base + __popc(m & ((1u << lane_id()) - 1u));
      active_next[pos] = s;
      st.qidx[(h + 1) & 1][pos] = j;
      st.packed[(h + 1) & 1][pos] = 0ull;
    }
  }
}

// Append converged (idx_a, idx_b) pairs as sort keys:  key = idx1<<33 | idx2<<1 | set.
__global__ void recip_collect(RecipBatch b, uint64_t* __restrict__ keys, int32_t* __restrict__ nkeys, int cap) {
  const RecipState& st = b.st[blockIdx.y];
  const int n = b.nseed[blockIdx.y];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool conv = i < n && st.notyet[i] == 0;
  unsigned m = __ballot_sync(0xffffffffu, conv);
  if (!m) return;
  int leader = __ffs(m) - 1;
  int base = 0;
  if (lane_id() == leader) base = atomicAdd(nkeys, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (conv) {
    int pos = base + __popc(m & ((1u << lane_id()) - 1u));
    uint32_t a = (uint32_t)st.xy[0][i], c = (uint32_t)st.xy[1][i];
    if (b.swap[blockIdx.y]) { uint32_t t = a; a = c; c = t; }
    if (pos < cap) keys[pos] = ((uint64_t)a << 33) | ((uint64_t)c << 1) | (uint64_t)b.set[blockIdx.y];
  }
}

__global__ void pack_pairs(const int32_t* __restrict__ idx1, const int32_t* __restrict__ idx2, int n,
                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int32_t* nkeys) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *nkeys = n;
  if (i >= n) return;
  keys[i] = ((uint64_t)(uint32_t)idx1[i] << 33) | ((uint64_t)(uint32_t)idx2[i] << 1);
  vals[i] = (uint32_t)i;
}

// Single-CTA ordered compaction of the first element of every run of equal (key>>1).
// Output modes: idx pairs (int32), first-occurrence positions, or xy (int64) + conf.
struct UniqueOut {
  int32_t* idx1; int32_t* idx2; int32_t* index;   // may be NULL
  int64_t* xy1; int64_t* xy2; float* conf;        // may be NULL
  const float* q1[2]; const float* q2[2];         // conf maps per descriptor set
  int W1, W2;
  int32_t* n_out;
};

__global__ void __launch_bounds__(1024)
unique_sorted(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
              const int32_t* __restrict__ nkeys, int cap, UniqueOut out) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int n = min(*nkeys, cap);
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int i = base + threadIdx.x;
    uint64_t k = 0;
    int flag = 0;
    if (i < n) {
      k = keys[i];
      flag = (i == 0) || ((keys[i - 1] >> 1) != (k >> 1));
    }
    int x = flag;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane_id() >= off) x += y;
    }
    if (lane_id() == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sums[threadIdx.x];
      int xs = w;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, xs, off);
        if (lane_id() >= off) xs += y;
      }
      warp_sums[threadIdx.x] = xs - w;
    }
    __syncthreads();
    int pos = carry_s + warp_sums[threadIdx.x >> 5] + x - flag;
    if (flag) {
      uint32_t a = (uint32_t)(k >> 33), b = (uint32_t)((k >> 1) & 0xffffffffu);
      int set = (int)(k & 1);
      if (out.idx1) { out.idx1[pos] = (int32_t)a; out.idx2[pos] = (int32_t)b; }
      if (out.index) out.index[pos] = (int32_t)vals[i];
      if (out.xy1) {
        out.xy1[2 * pos] = a % out.W1; out.xy1[2 * pos + 1] = a / out.W1;
        out.xy2[2 * pos] = b % out.W2; out.xy2[2 * pos + 1] = b / out.W2;
      }
      if (out.conf) out.conf[pos] = sqrtf(out.q1[set][a] * out.q2[set][b]);
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = pos + flag;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out.n_out = carry_s;
}

// One-CTA sort + unique for lists of up to 16384 keys - every 512 x 512 pair: 2 (n1 + n2) = 16384 candidates.  The
// generic chain (two LSD radix sorts = 19 launches of a handful of CTAs + the single-CTA compaction) cost ~250 us of a
// 2.45 ms pair; here the keys are sorted in registers (bitonic_reg.cuh: 1024 threads x EPT words, shuffles inside a warp,
// one shared-memory buffer for the longer spans) and compacted by the same CTA.  The sorted word is the key itself, or -
// when the caller wants the position of the first occurrence (merge_corres) - the key fields squeezed together with the
// element's position in the low bits, which reproduces the order of the stable radix sort exactly.
//   word = a << sh_a | b << sh_b | set << bv | position,  sh_b = 1 + bv,  (bv = 0, sh_a = 33: the raw key)
constexpr int SMALL_SORT_THREADS = 1024;
constexpr int SMALL_SORT_MAX = 16 * SMALL_SORT_THREADS;

// sur_bits > 0: the network runs on 32-bit surrogate words  a << sur_bits | position in the input  (one-instruction
// 32-bit min / max instead of two compares + four selects per 64-bit comparator, one register per shuffle), which orders
// the list by `a` and names every element; the full words are then gathered in that order into the (now free) exchange
// buffer and odd-even transposition passes restore the order inside the runs of equal `a` - a correspondence usually
// appears twice, from the two search directions, already in the right order, so the loop ends after a pass or two.  A
// list that keeps the passes busy (a degenerate map where thousands of seeds meet in one pixel) is handed to the 64-bit
// network.  Either way the result is the ascending order of the full words.
constexpr int SMALL_SORT_MAX_PASSES = 16;

template <int EPT>
__global__ void __launch_bounds__(SMALL_SORT_THREADS)
small_sort_unique(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const int32_t* __restrict__ nkeys,
                  int cap, int sh_a, int bv, int sur_bits, UniqueOut out) {
  ST3R_DYN_SMEM_U64(sx);                       // [SMALL_SORT_THREADS * EPT] exchange buffer
  __shared__ int warp_sums[32];
  __shared__ uint64_t warp_last[32];
  const int n = min(*nkeys, cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sh_b = 1 + bv;
  const uint64_t b_mask = (1ull << (sh_a - sh_b)) - 1ull;
  auto full_word = [&](int i) -> uint64_t {
    const uint64_t k = keys[i];
    return bv ? ((k >> 33) << sh_a) | (((k >> 1) & 0xffffffffull) << sh_b) | ((k & 1ull) << bv) | (uint64_t)vals[i] : k;
  };
  uint64_t v[EPT];
  const int i0 = threadIdx.x * EPT;
  bool sorted = false;                         // CTA-uniform
  if (sur_bits) {
    uint32_t sw[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e)
      sw[e] = i0 + e < n ? ((uint32_t)(keys[i0 + e] >> 33) << sur_bits) | (uint32_t)(i0 + e) : 0xffffffffu;
    st3r_sort::reg_bitonic_sort<SMALL_SORT_THREADS, EPT, 1, uint32_t>(sw, reinterpret_cast<uint32_t*>(sx), n);
    __syncthreads();
    const uint32_t pos_mask = (1u << sur_bits) - 1u;
#pragma unroll
    for (int e = 0; e < EPT; ++e)
      if (i0 + e < n) sx[i0 + e] = full_word((int)(sw[e] & pos_mask));
    __syncthreads();
    for (int pass = 0; pass < SMALL_SORT_MAX_PASSES && !sorted; ++pass) {
      int swapped = 0;
#pragma unroll
      for (int parity = 0; parity < 2; ++parity) {
        for (int p = 2 * (int)threadIdx.x + parity; p + 1 < n; p += 2 * SMALL_SORT_THREADS) {
          const uint64_t a = sx[p], b = sx[p + 1];
          if (a > b) { sx[p] = b; sx[p + 1] = a; swapped = 1; }
        }
        __syncthreads();
      }
      sorted = __syncthreads_count(swapped) == 0;
    }
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = i0 + e < n ? sx[i0 + e] : st3r_sort::SORT_PAD;
    __syncthreads();                           // everyone holds its words before a fallback sort reuses the buffer
  } else {
#pragma unroll
    for (int e = 0; e < EPT; ++e) v[e] = i0 + e < n ? full_word(i0 + e) : st3r_sort::SORT_PAD;
  }
  if (!sorted) st3r_sort::reg_bitonic_sort<SMALL_SORT_THREADS, EPT, 1>(v, sx, n);
  // first element of every run of equal (a, b): ordered compaction
  uint64_t prev = (uint64_t)__shfl_up_sync(0xffffffffu, (unsigned long long)v[EPT - 1], 1);
  if (lane == 31) warp_last[warp] = v[EPT - 1];
  __syncthreads();
  if (lane == 0 && warp > 0) prev = warp_last[warp - 1];
  int flags = 0, cnt = 0;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const uint64_t p = e == 0 ? prev : v[e - 1];
    const bool f = i0 + e < n && (i0 + e == 0 || (p >> sh_b) != (v[e] >> sh_b));
    flags |= (f ? 1 : 0) << e;
    cnt += f ? 1 : 0;
  }
  int x = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= off) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    const int w = warp_sums[lane];
    int xs = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, xs, off);
      if (lane >= off) xs += y;
    }
    warp_sums[lane] = xs - w;
    if (lane == 31) *out.n_out = xs;
  }
  __syncthreads();
  int pos = warp_sums[warp] + x - cnt;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    if (!((flags >> e) & 1)) continue;
    const uint64_t k = v[e];
    const uint32_t a = (uint32_t)(k >> sh_a), b = (uint32_t)((k >> sh_b) & b_mask);
    const int set = (int)((k >> bv) & 1ull);
    if (out.idx1) { out.idx1[pos] = (int32_t)a; out.idx2[pos] = (int32_t)b; }
    if (out.index) out.index[pos] = (int32_t)(k & ((1ull << bv) - 1ull));
    if (out.xy1) {
      out.xy1[2 * pos] = a % out.W1; out.xy1[2 * pos + 1] = a / out.W1;
      out.xy2[2 * pos] = b % out.W2; out.xy2[2 * pos + 1] = b / out.W2;
    }
    if (out.conf) out.conf[pos] = sqrtf(out.q1[set][a] * out.q2[set][b]);
    ++pos;
  }
}

__global__ void nn_decode(const unsigned long long* __restrict__ packed, int M, int32_t* idx, float* best) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  unsigned long long k = packed[i];
  if (idx) idx[i] = k ? nn_unpack_idx(k) : -1;
  if (best) best[i] = k ? nn_unpack_score(k) : -INFINITY;
}

int bits_for(long long n) {  // number of bits needed to represent values < n
  int b = 1;
  while ((1ll << b) < n) ++b;
  return b;
}

struct RecipWs {
  RecipState st[RB];
  uint64_t* keys; uint64_t* keys_alt;
  uint32_t* vals; uint32_t* vals_alt;
  int32_t* nkeys;
  float* norm_bound;  // [2]
  void* sort_ws; size_t sort_ws_bytes;
  float* split_hi[RB]; float* split_lo[RB];   // tf32 head / tail of the maps (split-precision matcher), else null
};

// `split_rows` (optional, n_split entries): rows of the descriptor maps whose tf32 head / tail arrays the
// split-precision matcher needs; they are carved last so that the layout of everything else does not depend on them.
size_t carve(RecipWs* w, void* ws, size_t ws_bytes, int nseed_max, int key_cap, int max_iter, bool dry, int nprob = 1,
             const int* split_rows = nullptr, int n_split = 0, int d = 0) {
  WsAlloc a(dry ? (void*)0 : ws, dry ? (size_t)-1 : ws_bytes);
  RecipWs t;
  for (int p = 0; p < RB; ++p) {
    if (p >= nprob) { t.st[p] = t.st[0]; continue; }
    for (int s = 0; s < 2; ++s) {
      t.st[p].xy[s] = a.take<int32_t>(nseed_max);
      t.st[p].old[s] = a.take<int32_t>(nseed_max);
      t.st[p].active[s] = a.take<int32_t>(nseed_max);
      t.st[p].qidx[s] = a.take<int32_t>(nseed_max);
      t.st[p].packed[s] = a.take<unsigned long long>(nseed_max);
    }
    t.st[p].notyet = a.take<uint8_t>(nseed_max);
    t.st[p].counts = a.take<int32_t>(2 * max_iter + 2);
  }
  t.keys = a.take<uint64_t>(key_cap);
  t.keys_alt = a.take<uint64_t>(key_cap);
  t.vals = a.take<uint32_t>(key_cap);
  t.vals_alt = a.take<uint32_t>(key_cap);
  t.nkeys = a.take<int32_t>(4);
  t.norm_bound = a.take<float>(8);
  t.sort_ws_bytes = radix_sort_ws_bytes(key_cap);
  t.sort_ws = a.take<char>(t.sort_ws_bytes);
  for (int k = 0; k < RB; ++k) {
    t.split_hi[k] = t.split_lo[k] = nullptr;
    if (split_rows && k < n_split) {
      t.split_hi[k] = a.take<float>((size_t)split_rows[k] * d);
      t.split_lo[k] = a.take<float>((size_t)split_rows[k] * d);
    }
  }
  if (w) *w = t;
  return a.off + 256;
}

int seed_count(int H, int W, int S, int* nx_out) {
  int off = S / 2;
  int ny = H > off ? (H - off + S - 1) / S : 0;
  int nx = W > off ? (W - off + S - 1) / S : 0;
  if (nx_out) *nx_out = nx;
  return ny * nx;
}

struct RecipProblem {
  const float* P1; int HW1, W1;      // seeds live in map 1
  const float* P2; int HW2;
  int nseed, nx; const int32_t* seeds;
  int swap, set;
  const float* norm1; const float* norm2;   // device max ||row||^2 of P1 / P2 (tcgen05 path)
  const float* hi1 = nullptr; const float* lo1 = nullptr;   // tf32 head / tail of P1 / P2 (split-precision variant)
  const float* hi2 = nullptr; const float* lo2 = nullptr;
};

// n seeded reciprocal searches P1 -> P2 -> P1 ... advanced in lock-step; converged pairs are appended to w.keys.
int run_recip_batch(const RecipWs& w, const RecipProblem* pr, int n, int d, int S, int max_iter, int key_cap, int impl,
                    cudaStream_t stream) {
  RecipBatch b;
  b.S = S; b.n = n;
  int max_seed = 0;
  for (int p = 0; p < RB; ++p) {
    const RecipProblem& q = pr[p < n ? p : 0];
    b.st[p] = w.st[p < n ? p : 0];
    b.nseed[p] = p < n ? q.nseed : 0;
    b.nx[p] = q.nx > 0 ? q.nx : 1; b.W1[p] = q.W1; b.swap[p] = q.swap; b.set[p] = q.set; b.seeds[p] = q.seeds;
    if (p < n) max_seed = max(max_seed, q.nseed);
  }
  if (max_seed <= 0) return ST3R_OK;
  const int ncounts = 2 * max_iter + 2;
  recip_init<<<dim3(nblk(max(max_seed, ncounts)), n), TPB, 0, stream>>>(b, ncounts);
  ST3R_CHECK_LAUNCH();
  const bool use_tc = (impl == ST3R_NN_TCGEN05) || (impl == ST3R_NN_AUTO && nn_tc_supported(d));
  int h = 0;
  for (int it = 0; it < max_iter; ++it) {
    for (int half = 0; half < 2; ++half, ++h) {
      // half 0: query rows P1[xy1[active]] against DB P2 -> xy2;  half 1: the converse.
      const int dst = 1 - half;
      NnBatchItem items[RB];
      for (int p = 0; p < n; ++p) {
        const RecipProblem& q = pr[p];
        const RecipState& st = w.st[p];
        items[p] = NnBatchItem{half == 0 ? q.P1 : q.P2, st.qidx[h & 1], st.counts + h, q.nseed,
                               half == 0 ? q.P2 : q.P1, half == 0 ? q.HW2 : q.HW1, half == 0 ? q.norm2 : q.norm1,
                               st.packed[h & 1], half == 0 ? q.hi2 : q.hi1, half == 0 ? q.lo2 : q.lo1};
      }
      int rc = ST3R_OK;
      if (use_tc) {
        rc = nn_tc_launch_batch(items, n, d, stream);
      } else {
        for (int p = 0; p < n && rc == ST3R_OK; ++p)
          rc = nn_simt_launch(items[p].Qsrc, items[p].qidx, items[p].count_ptr, items[p].Mmax, items[p].DB, items[p].N, d,
                              items[p].packed, stream);
      }
      if (rc) return rc;
      recip_update<<<dim3(nblk(max_seed), n), TPB, 0, stream>>>(b, h, dst);
      ST3R_CHECK_LAUNCH();
    }
  }
  recip_collect<<<dim3(nblk(max_seed), n), TPB, 0, stream>>>(b, w.keys, w.nkeys, key_cap);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

// 2 (default): lists of up to 16384 keys are sorted and compacted by one CTA (small_sort_unique) through 32-bit surrogate
// words where the fields fit; 1: the same CTA on the 64-bit words; 0: always the radix chain (the first implementation).
// 1 and 0 stay as cross-checks: st3r_recip_set_variant, tests/test_match_gpu.py runs all of them.
int g_small_sort = 2;

int sort_and_unique(const RecipWs& w, int key_cap, int HW1, int HW2, bool with_vals, const UniqueOut& out,
                    cudaStream_t stream) {
  // keys: idx1<<33 | idx2<<1 | set.  LSD: low field [0, 1+bits2) then high field [33, 33+bits1).
  int b2 = bits_for(HW2) + 1, b1 = bits_for(HW1);
  const int bv = with_vals ? bits_for(key_cap) : 0;
  if (g_small_sort && key_cap <= SMALL_SORT_MAX && b1 + b2 + bv < 64) {
    // (with positions the fields are squeezed together: a above b above set above the position)
    const int sh_a = with_vals ? b2 + bv : 33;
    static PerDeviceOnce attr_set;
    if (!attr_set.done()) {
      ST3R_CHECK_CUDA(cudaFuncSetAttribute(small_sort_unique<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(sizeof(uint64_t) * SMALL_SORT_MAX)));
      ST3R_CHECK_CUDA(cudaFuncSetAttribute(small_sort_unique<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(sizeof(uint64_t) * SMALL_SORT_MAX / 2)));
      attr_set.mark();
    }
    const uint32_t* vv = with_vals ? w.vals : nullptr;
    const int pos_bits = bits_for(key_cap);
    const int sur = (g_small_sort == 2 && b1 + pos_bits <= 32) ? pos_bits : 0;
    if (key_cap <= SMALL_SORT_THREADS)
      small_sort_unique<1><<<1, SMALL_SORT_THREADS, sizeof(uint64_t) * SMALL_SORT_THREADS, stream>>>(w.keys, vv, w.nkeys, key_cap, sh_a, bv, sur, out);
    else if (key_cap <= 2 * SMALL_SORT_THREADS)
      small_sort_unique<2><<<1, SMALL_SORT_THREADS, sizeof(uint64_t) * SMALL_SORT_THREADS * 2, stream>>>(w.keys, vv, w.nkeys, key_cap, sh_a, bv, sur, out);
    else if (key_cap <= 4 * SMALL_SORT_THREADS)
      small_sort_unique<4><<<1, SMALL_SORT_THREADS, sizeof(uint64_t) * SMALL_SORT_THREADS * 4, stream>>>(w.keys, vv, w.nkeys, key_cap, sh_a, bv, sur, out);
    else if (key_cap <= 8 * SMALL_SORT_THREADS)
      small_sort_unique<8><<<1, SMALL_SORT_THREADS, sizeof(uint64_t) * SMALL_SORT_THREADS * 8, stream>>>(w.keys, vv, w.nkeys, key_cap, sh_a, bv, sur, out);
    else
      small_sort_unique<16><<<1, SMALL_SORT_THREADS, sizeof(uint64_t) * SMALL_SORT_THREADS * 16, stream>>>(w.keys, vv, w.nkeys, key_cap, sh_a, bv, sur, out);
    ST3R_CHECK_LAUNCH();
    return ST3R_OK;
  }
  uint32_t* v = with_vals ? w.vals : nullptr;
  uint32_t* va = with_vals ? w.vals_alt : nullptr;
  int rc = radix_sort_pairs(w.keys, v, w.keys_alt, va, w.nkeys, key_cap, 0, b2, w.sort_ws,
                            w.sort_ws_bytes, stream);
  if (rc) return rc;
  rc = radix_sort_pairs(w.keys, v, w.keys_alt, va, w.nkeys, key_cap, 33, 33 + b1, w.sort_ws,
                        w.sort_ws_bytes, stream);
  if (rc) return rc;
  unique_sorted<<<1, 1024, 0, stream>>>(w.keys, v, w.nkeys, key_cap, out);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI (declared in include/starst3r_b200.h)
// ---------------------------------------------------------------------------------------------
extern "C" {

size_t st3r_nn_argmax_ws_bytes(int M, int N, int d) {
  size_t b = st3r_align_up((size_t)max(M, 1) * sizeof(unsigned long long), 256) +
             st3r_align_up((size_t)max(M, 1) * sizeof(int32_t), 256) + 1024;
  if (nn_tc_split_enabled() && nn_tc_supported(d))   // tf32 head / tail of the DB
    b += 2 * st3r_align_up((size_t)max(N, 1) * d * sizeof(float), 256);
  return b;
}

int st3r_nn_argmax(const float* Q, int M, const float* DB, int N, int d, int32_t* idx, float* best,
                   void* ws, size_t ws_bytes, int impl, cudaStream_t stream) {
  ST3R_CHECK_ARG(M >= 0 && N >= 0 && d > 0, "st3r_nn_argmax: bad sizes M=%d N=%d d=%d", M, N, d);
  if (M == 0) return ST3R_OK;
  ST3R_CHECK_ARG(Q && DB && idx, "st3r_nn_argmax: null pointer");
  ST3R_CHECK_ARG(ws && ws_bytes >= st3r_nn_argmax_ws_bytes(M, N, d), "st3r_nn_argmax: workspace too small");
  WsAlloc a(ws, ws_bytes);
  unsigned long long* packed = a.take<unsigned long long>(M);
  float* bound = a.take<float>(4);
  ST3R_CHECK_CUDA(cudaMemsetAsync(packed, 0, (size_t)M * sizeof(unsigned long long), stream));
  int rc;
  bool use_tc = (impl == ST3R_NN_TCGEN05) || (impl == ST3R_NN_AUTO && nn_tc_supported(d) && M >= 64);
  if (impl == ST3R_NN_TCGEN05) ST3R_CHECK_ARG(nn_tc_supported(d), "st3r_nn_argmax: tcgen05 path needs d=24");
  if (use_tc && N > 0) {
    rc = nn_db_norm_launch(DB, N, d, bound, stream);
    if (rc) return rc;
    float* hi = nullptr;
    float* lo = nullptr;
    if (nn_tc_split_enabled()) {
      hi = a.take<float>((size_t)N * d);
      lo = a.take<float>((size_t)N * d);
      rc = nn_tc_split_launch(DB, N, d, hi, lo, stream);
      if (rc) return rc;
    }
    rc = nn_tc_launch(Q, nullptr, nullptr, M, DB, N, d, bound, packed, stream, hi, lo);
  } else {
    rc = nn_simt_launch(Q, nullptr, nullptr, M, DB, N, d, packed, stream);
  }
  if (rc) return rc;
  nn_decode<<<nblk(M), TPB, 0, stream>>>(packed, M, idx, best);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_recip_seed_count(int H, int W, int subsample) { return seed_count(H, W, subsample, nullptr); }

size_t st3r_recip_nn_ws_bytes(int nseed_max, int key_cap, int max_iter) {
  return carve(nullptr, nullptr, 0, max(nseed_max, 1), max(key_cap, 1), max(max_iter, 1), true);
}

int st3r_recip_nn(const float* P1, int H1, int W1, const float* P2, int H2, int W2, int d, int subsample,
                  const int32_t* seeds, int nseeds, int max_iter, int32_t* out_idx1, int32_t* out_idx2,
                  int32_t* n_out, void* ws, size_t ws_bytes, int impl, cudaStream_t stream) {
  ST3R_CHECK_ARG(P1 && P2 && out_idx1 && out_idx2 && n_out, "st3r_recip_nn: null pointer");
  ST3R_CHECK_ARG(H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && d > 0 && max_iter > 0, "st3r_recip_nn: bad sizes");
  ST3R_CHECK_ARG((long long)H1 * W1 < (1ll << 30) && (long long)H2 * W2 < (1ll << 30), "st3r_recip_nn: map too large");
  int nx = 0;
  int nseed = seeds ? nseeds : seed_count(H1, W1, subsample, &nx);
  ST3R_CHECK_ARG(seeds || subsample > 0, "st3r_recip_nn: subsample must be > 0");
  RecipWs w;
  size_t need = carve(&w, ws, ws_bytes, max(nseed, 1), max(nseed, 1), max_iter, false);
  ST3R_CHECK_ARG(ws && ws_bytes >= need, "st3r_recip_nn: workspace too small (%zu < %zu)", ws_bytes, need);
  ST3R_CHECK_CUDA(cudaMemsetAsync(w.nkeys, 0, 4 * sizeof(int32_t), stream));
  if (nseed == 0) { ST3R_CHECK_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), stream)); return ST3R_OK; }
  const bool use_tc = (impl == ST3R_NN_TCGEN05) || (impl == ST3R_NN_AUTO && nn_tc_supported(d));
  if (use_tc) {
    int rc = nn_db_norm_launch(P1, H1 * W1, d, w.norm_bound + 0, stream);
    if (rc) return rc;
    rc = nn_db_norm_launch(P2, H2 * W2, d, w.norm_bound + 1, stream);
    if (rc) return rc;
  }
  RecipProblem pr{P1, H1 * W1, W1, P2, H2 * W2, nseed, nx, seeds, 0, 0, w.norm_bound + 0, w.norm_bound + 1};
  int rc = run_recip_batch(w, &pr, 1, d, subsample, max_iter, nseed, impl, stream);
  if (rc) return rc;
  UniqueOut out = {};
  out.idx1 = out_idx1; out.idx2 = out_idx2; out.n_out = n_out; out.W1 = W1; out.W2 = W2;
  return sort_and_unique(w, nseed, H1 * W1, H2 * W2, false, out, stream);
}

int st3r_recip_set_variant(int variant) {
  ST3R_CHECK_ARG(variant >= 0 && variant <= 2, "st3r_recip_set_variant: unknown variant %d", variant);
  g_small_sort = variant;
  return ST3R_OK;
}

size_t st3r_merge_corres_ws_bytes(int n) { return carve(nullptr, nullptr, 0, 1, max(n, 1), 1, true); }

int st3r_merge_corres(const int32_t* idx1, const int32_t* idx2, int n, int hw1, int hw2, int32_t* out_idx1,
                      int32_t* out_idx2, int32_t* out_index, int32_t* n_out, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
  ST3R_CHECK_ARG(n >= 0 && n_out, "st3r_merge_corres: bad args");
  if (n == 0) { ST3R_CHECK_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), stream)); return ST3R_OK; }
  ST3R_CHECK_ARG(idx1 && idx2 && out_idx1 && out_idx2, "st3r_merge_corres: null pointer");
  ST3R_CHECK_ARG(hw1 > 0 && hw2 > 0 && hw1 < (1 << 30) && hw2 < (1 << 30), "st3r_merge_corres: bad map sizes");
  RecipWs w;
  size_t need = carve(&w, ws, ws_bytes, 1, n, 1, false);
  ST3R_CHECK_ARG(ws && ws_bytes >= need, "st3r_merge_corres: workspace too small");
  pack_pairs<<<nblk(n), TPB, 0, stream>>>(idx1, idx2, n, w.keys, w.vals, w.nkeys);
  ST3R_CHECK_LAUNCH();
  UniqueOut out = {};
  out.idx1 = out_idx1; out.idx2 = out_idx2; out.index = out_index; out.n_out = n_out; out.W1 = 1; out.W2 = 1;
  return sort_and_unique(w, n, hw1, hw2, true, out, stream);
}

int st3r_extract_corres_cap(int H1, int W1, int H2, int W2, int subsample) {
  return 2 * (seed_count(H1, W1, subsample, nullptr) + seed_count(H2, W2, subsample, nullptr));
}

size_t st3r_extract_corres_ws_bytes(int H1, int W1, int H2, int W2, int subsample, int max_iter) {
  int n1 = seed_count(H1, W1, subsample, nullptr), n2 = seed_count(H2, W2, subsample, nullptr);
  const int rows[4] = {H1 * W1, H2 * W2, H1 * W1, H2 * W2};
  const bool split = nn_tc_split_enabled();     // the split-precision matcher only exists for d = 24
  return carve(nullptr, nullptr, 0, max(max(n1, n2), 1), max(2 * (n1 + n2), 1), max_iter, true, 4, split ? rows : nullptr,
               4, 24);
}

int st3r_extract_corres(const float* feat11, const float* feat21, const float* feat22, const float* feat12,
                        const float* qonf11, const float* qonf21, const float* qonf22, const float* qonf12,
                        int H1, int W1, int H2, int W2, int d, int subsample, int max_iter,
                        int64_t* out_xy1, int64_t* out_xy2, float* out_conf, int32_t* n_out,
                        void* ws, size_t ws_bytes, int impl, cudaStream_t stream) {
  ST3R_CHECK_ARG(feat11 && feat21 && feat22 && feat12 && qonf11 && qonf21 && qonf22 && qonf12,
                 "st3r_extract_corres: null input");
  ST3R_CHECK_ARG(out_xy1 && out_xy2 && out_conf && n_out, "st3r_extract_corres: null output");
  ST3R_CHECK_ARG(H1 > 0 && W1 > 0 && H2 > 0 && W2 > 0 && d > 0 && subsample > 0 && max_iter > 0,
                 "st3r_extract_corres: bad sizes");
  ST3R_CHECK_ARG((long long)H1 * W1 < (1ll << 30) && (long long)H2 * W2 < (1ll << 30), "st3r_extract_corres: map too large");
  int nx1, nx2;
  int n1 = seed_count(H1, W1, subsample, &nx1), n2 = seed_count(H2, W2, subsample, &nx2);
  int cap = 2 * (n1 + n2);
  RecipWs w;
  const bool use_tc = (impl == ST3R_NN_TCGEN05) || (impl == ST3R_NN_AUTO && nn_tc_supported(d));
  const bool split = use_tc && nn_tc_split_enabled() && nn_tc_supported(d);
  const int rows[4] = {H1 * W1, H2 * W2, H1 * W1, H2 * W2};
  size_t need = carve(&w, ws, ws_bytes, max(max(n1, n2), 1), max(cap, 1), max_iter, false, 4, split ? rows : nullptr, 4, d);
  ST3R_CHECK_ARG(ws && ws_bytes >= need, "st3r_extract_corres: workspace too small (%zu < %zu)", ws_bytes, need);
  ST3R_CHECK_CUDA(cudaMemsetAsync(w.nkeys, 0, 4 * sizeof(int32_t), stream));
  // sparse_ga.py:612-620 - two descriptor sets, each matched in both directions: four searches in lock-step.
  const float* maps[4] = {feat11, feat21, feat12, feat22};
  if (use_tc) {
    for (int k = 0; k < 4; ++k) {
      int rc = nn_db_norm_launch(maps[k], rows[k], d, w.norm_bound + k, stream);
      if (rc) return rc;
      if (split) {
        rc = nn_tc_split_launch(maps[k], rows[k], d, w.split_hi[k], w.split_lo[k], stream);
        if (rc) return rc;
      }
    }
  }
  const float* const* sh = w.split_hi;   // null unless `split`
  const float* const* sl = w.split_lo;
  RecipProblem pr[4] = {
      {feat11, H1 * W1, W1, feat21, H2 * W2, n1, nx1, nullptr, 0, 0, w.norm_bound + 0, w.norm_bound + 1, sh[0], sl[0], sh[1], sl[1]},
      {feat21, H2 * W2, W2, feat11, H1 * W1, n2, nx2, nullptr, 1, 0, w.norm_bound + 1, w.norm_bound + 0, sh[1], sl[1], sh[0], sl[0]},
      {feat12, H1 * W1, W1, feat22, H2 * W2, n1, nx1, nullptr, 0, 1, w.norm_bound + 2, w.norm_bound + 3, sh[2], sl[2], sh[3], sl[3]},
      {feat22, H2 * W2, W2, feat12, H1 * W1, n2, nx2, nullptr, 1, 1, w.norm_bound + 3, w.norm_bound + 2, sh[3], sl[3], sh[2], sl[2]}};
  {
    int rc = run_recip_batch(w, pr, 4, d, subsample, max_iter, cap, impl, stream);
    if (rc) return rc;
  }
  UniqueOut out = {};
  out.xy1 = out_xy1; out.xy2 = out_xy2; out.conf = out_conf; out.n_out = n_out;
  out.q1[0] = qonf11; out.q2[0] = qonf21; out.q1[1] = qonf12; out.q2[1] = qonf22;
  out.W1 = W1; out.W2 = W2;
  return sort_and_unique(w, cap, H1 * W1, H2 * W2, false, out, stream);
}

}  // extern "C"
