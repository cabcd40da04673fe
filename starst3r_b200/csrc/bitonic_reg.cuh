// Register-resident bitonic sort of 64-bit words by one CTA (the "flip" form of the network: all comparators point the
// same way, so padding to a power of two is virtual).  Shared by the per-tile depth sort of the binning (gs_bin.cu, 256
// threads, up to 2048 words) and the correspondence merge of the matcher (recip.cu, 1024 threads, up to 16384 words).
#pragma once
#include <stdint.h>

namespace st3r_sort {

// (Round-2 profile of the shared-memory network above on the headline frame, 458 pairs per tile on average: 12.8 k
// warp instructions per tile, 70 % of them index arithmetic, predicates and branches of the generic loops, the
// memory accesses generic LD / ST because `buf` may point to either space.)  Here a thread holds EPT = P / 256
// consecutive elements of the padded segment in registers and the same comparator network is unrolled at compile
// time: comparators that stay inside a thread are register compare-exchanges, those inside a warp exchange through
// __shfl_xor (the mirror step sends element EPT-1-e), and only the spans of 32 * EPT elements and more go through
// shared memory (two alternating buffers, one barrier per step).  Same comparators, same +inf padding => the same
// unique order (the 64-bit words are distinct).
// (FP64 min / max would order these words too - a positive finite fp32 depth in the high half makes the word a positive
// finite double - but sm_100a has no DMNMX: fmin(double) expands to DSETP + selects, slower than the integer compare.)
constexpr uint64_t SORT_PAD = ~0ull;                      // +inf padding: above every word, never moves
// The network is generic in the word type T: uint64_t (default) or uint32_t.  sm_100a has 32-bit integer min / max
// (one instruction, the predicated form selects min or max) but no 64-bit one: a 64-bit comparator is two compares and
// four selects, a 32-bit one two instructions, and a shuffle step moves one register instead of two.  gs_bin.cu sorts
// 32-bit surrogate keys where it can and repairs the order afterwards.
__device__ __forceinline__ uint64_t kmin(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t kmax(uint64_t a, uint64_t b) { return a < b ? b : a; }
__device__ __forceinline__ uint32_t kmin(uint32_t a, uint32_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t kmax(uint32_t a, uint32_t b) { return a < b ? b : a; }
__device__ __forceinline__ uint64_t shfl_xor_word(uint64_t v, int m) {
  return (uint64_t)__shfl_xor_sync(0xffffffffu, (unsigned long long)v, m);
}
__device__ __forceinline__ uint32_t shfl_xor_word(uint32_t v, int m) {
  return (uint32_t)__shfl_xor_sync(0xffffffffu, (int)v, m);
}
template <typename T>
__device__ __forceinline__ void cmpswap(T& a, T& b) {
  const T lo = kmin(a, b), hi = kmax(a, b);
  a = lo; b = hi;
}

// Warps whose elements are all padding (index >= n_act, the segment length rounded up to a warp's 32 * EPT elements)
// sit the network out: padding never moves (every comparator puts the minimum at the lower index and the padding
// holds the highest indices), so a comparator with such an element is a no-op for both sides.  They only keep the
// block barriers of the shared-memory steps company.  The work then scales with ceil(n / (32 EPT)) warps, not with P.
template <int THREADS, int EPT, int NBUF, typename T>
__device__ __forceinline__ void reg_exchange_smem(T (&v)[EPT], T* sx, int& buf, int xor_mask, int low_bit,
                                                  bool active, int n_act) {
  constexpr int P = THREADS * EPT;
  T* b = sx + buf * P;
  const int i0 = threadIdx.x * EPT;
  // element i = t * EPT + e is kept at b[e * THREADS + t]: the lanes of a warp touch consecutive words both when they
  // write their own elements and when they read their partners' (a partner differs in the bits of t above the lane, or
  // is the mirror image inside an aligned group of lanes), so the exchange is free of bank conflicts
  if (active) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) b[e * THREADS + threadIdx.x] = v[e];
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int i = i0 + e, partner = i ^ xor_mask;
      if (partner < n_act) {
        const T o = b[(partner % EPT) * THREADS + partner / EPT];
        v[e] = (i & low_bit) == 0 ? kmin(v[e], o) : kmax(v[e], o);
      }
    }
  }
  if constexpr (NBUF == 2) buf ^= 1;      // the next exchange writes the other buffer: no second barrier needed
  else __syncthreads();                   // one buffer (large P): everyone has read before the next exchange writes
}

// The network is unrolled through template recursion (K = block size of the stage, J = comparator span), so that every
// register index and every step kind is a compile-time constant.
template <int THREADS, int EPT, int NBUF, int J, typename T>
struct HalfSteps {       // element i against i + J, then J / 2, ... 1
  static __device__ __forceinline__ void run(T (&v)[EPT], T* sx, int& buf, int lane, bool active, int n_act) {
    if constexpr (J >= 1) {
      if constexpr (J < EPT) {
        if (active) {
#pragma unroll
          for (int e = 0; e < EPT; ++e)
            if ((e & J) == 0) cmpswap(v[e], v[e | J]);
        }
      } else if constexpr (J < 32 * EPT) {
        if (active) {
          constexpr int m = J / EPT;
          const bool keep_min = (lane & m) == 0;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const T o = shfl_xor_word(v[e], m);
            v[e] = keep_min ? kmin(v[e], o) : kmax(v[e], o);
          }
        }
      } else {
        reg_exchange_smem<THREADS, EPT, NBUF, T>(v, sx, buf, J, J, active, n_act);
      }
      HalfSteps<THREADS, EPT, NBUF, J / 2, T>::run(v, sx, buf, lane, active, n_act);
    }
  }
};

template <int THREADS, int EPT, int NBUF, int K, typename T>
struct Stages {          // stages K, 2K, ... P
  static __device__ __forceinline__ void run(T (&v)[EPT], T* sx, int& buf, int lane, bool active, int n_act) {
    if constexpr (K <= THREADS * EPT) {
      // first step of the stage: element i against i ^ (K - 1) (mirror inside every block of K)
      if constexpr (K <= EPT) {
        if (active) {
#pragma unroll
          for (int e = 0; e < EPT; ++e)
            if ((e & (K >> 1)) == 0) cmpswap(v[e], v[e ^ (K - 1)]);
        }
      } else if constexpr (K <= 32 * EPT) {
        if (active) {
          constexpr int m = K / EPT - 1;                   // lane mask of the partner thread
          const bool keep_min = (lane & ((m + 1) >> 1)) == 0;
          // element e meets the partner thread's element EPT - 1 - e: two at a time, so that only two words are in flight
          // (EPT = 16 at 1024 threads has 64 registers per thread)
#pragma unroll
          for (int e = 0; e < (EPT + 1) / 2; ++e) {
            const int f = EPT - 1 - e;
            const T of = shfl_xor_word(v[f], m);
            const T oe = shfl_xor_word(v[e], m);
            v[e] = keep_min ? kmin(v[e], of) : kmax(v[e], of);
            if (f != e) v[f] = keep_min ? kmin(v[f], oe) : kmax(v[f], oe);
          }
        }
      } else {
        reg_exchange_smem<THREADS, EPT, NBUF, T>(v, sx, buf, K - 1, K >> 1, active, n_act);
      }
      HalfSteps<THREADS, EPT, NBUF, K / 4, T>::run(v, sx, buf, lane, active, n_act);
      Stages<THREADS, EPT, NBUF, K * 2, T>::run(v, sx, buf, lane, active, n_act);
    }
  }
};

// v: this thread's EPT consecutive elements of the padded sequence (element index threadIdx.x * EPT + e; SORT_PAD beyond
// n); sx: NBUF * THREADS * EPT words of shared memory (unused when 32 * EPT >= THREADS * EPT); n: live elements.
template <int THREADS, int EPT, int NBUF, typename T = uint64_t>
__device__ __forceinline__ void reg_bitonic_sort(T (&v)[EPT], T* sx, int n) {
  int buf = 0;
  const int n_act = (n + 32 * EPT - 1) / (32 * EPT) * (32 * EPT);
  const bool active = (int)threadIdx.x * EPT < n_act;              // warp-uniform
  Stages<THREADS, EPT, NBUF, 2, T>::run(v, sx, buf, (int)(threadIdx.x & 31), active, n_act);
}


}  // namespace st3r_sort
