// Dot-product nearest neighbour on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Replaces the `A @ B.T` + torch.max of mast3r/mast3r/fast_nn.py:30-67.  The H*W x H*W x 24
// descriptor correlation is a dense contraction, so it runs as kind::tf32 UMMA:
//   TMA (cp.async.bulk.tensor, 128B swizzle, OOB zero-fill pads K 24 -> 32) -> smem ring
//   -> tcgen05.mma (M=128, N=128, K=8 x3) -> fp32 accumulators in TMEM (2 stages)
//   -> tcgen05.ld -> fused arg-max epilogue.  The score matrix never reaches HBM.
//
// Exactness.  The reference scores in true fp32 (sequential FMA chain).  TF32 drops 13
// mantissa bits of each operand, so |approx - exact| <= eps = 2^-9 * |q| * max_j |db_j|.
// The epilogue keeps a running approximate row maximum and appends every column within
// delta = 2 eps of it to a small per-row candidate list in shared memory (entries that fall out
// of the band as the maximum rises are compacted away).  At the end of a CTA's DB range the
// survivors (typically 1-3) are re-scored with the exact fp32 FMA chain; the true arg-max and all
// its exact ties are always among them, so the result is bit-identical to the SIMT kernel / the
// reference (ties -> lowest index).  When a list fills up - smooth descriptor fields such as real
// MASt3R maps or the synthetic scene put ~100 columns per row inside the band - it is resolved on
// the spot: its entries are re-scored exactly, the exact best (score, lowest index) is kept in a
// register and raises the running bound (a column can only win if its approximate score is
// >= best_exact - eps), and the list starts empty again.  (A first version flagged such rows for
// an exact warp-per-row kernel: 100 ms per pair on the synthetic scene instead of 2.5 ms.)
#ifndef ST3R_HOST_EMU   // (CPU emulator builds, tests/host/: a software model of TMA / mbarriers / tcgen05 stands in)
#include <cuda.h>
#endif
#include "common.cuh"
#include "nn.cuh"
#include "../../include/starst3r_b200.h"

// Default: one UMMA-issuing warp per query tile with per-(query tile, accumulator stage) barriers, so that the two
// accumulator pipelines of a CTA only share the DB tile in shared memory (B200, M = 262144: 7.73 -> 7.45 ms, split
// precision pairs 5.09 -> 4.72 ms against ONE issuing thread for both tiles).  -DNN_TC_ONE_ISSUER restores the single
// MMA thread with shared barriers (the instrumented timing experiments of scripts/build_dbg.sh are written against it);
// -DNN_TC_ONE_ISSUER -DNN_TC_DECOUPLE is that thread with the per-tile barriers (measured slower than both).
#if !defined(NN_TC_ONE_ISSUER)
#define NN_TC_TWO_ISSUERS
#define NN_TC_DECOUPLE
#endif

namespace {

// One CTA per SM scores MH = 2 query tiles of 128 rows against ONE stream of DB tiles: every DB tile that TMA brings
// in from L2 feeds two UMMA groups.  With one 128-row tile per CTA (two CTAs per SM, each with its own stream) the
// TMA stream alone - MMAs, TMEM reads and arithmetic compiled out - ran at 640 cycles per tile per CTA: 12 KB per
// 320 cycles per SM is 38 B/clk/SM, the L2 -> SM bandwidth of the chip (~42 B/clk/SM), i.e. the kernel was L2-bound at
// 64 FLOP per L2 byte.  Sharing the stream halves the L2 traffic per FLOP.
constexpr int UM = 128;            // rows of one UMMA (M)
constexpr int MH = 2;              // query tiles per CTA
constexpr int BM = UM * MH;        // query rows per CTA
#ifndef NN_TC_BN
#define NN_TC_BN 128
#endif
constexpr int BN = NN_TC_BN;       // DB rows per MMA tile (UMMA N)
constexpr int DK = 24;             // descriptor dim
constexpr int ROWB = 128;          // smem bytes per operand row (32 floats, 24 real + 8 zero)
// All 512 TMEM columns hold accumulators: the kernel's throughput is (columns in flight) / (round trip MMA issue ->
// commit -> epilogue read -> release), and narrower tiles shorten the round trip (its UMMA and TMEM-read parts scale
// with the tile width, its barrier hops do not).
constexpr int ACC_STAGES = 256 / BN;  // TMEM accumulator stages per query tile
// Warps 0..7: epilogue (warp % 4 = TMEM lane quarter, warp / 4 = query tile); warp 8: TMA producer; warp 9: MMA issuer.
constexpr int EPI_COLS = BN;
constexpr int EPI_CHUNKS = EPI_COLS / 32;
constexpr int EPI_THREADS = BM;                 // one thread per query row
constexpr int EPI_WARPS = EPI_THREADS / 32;
constexpr int WARP_TMA = EPI_WARPS, WARP_MMA = EPI_WARPS + 1;
#ifdef NN_TC_TWO_ISSUERS
constexpr int MMA_WARPS = MH;                   // one UMMA-issuing warp per query tile
#else
constexpr int MMA_WARPS = 1;
#endif
constexpr int NUM_THREADS = EPI_THREADS + 32 + 32 * MMA_WARPS;
constexpr int TMEM_COLS = MH * ACC_STAGES * BN;  // 512: the whole tensor memory of the SM
constexpr int MAX_TILES_PER_CHUNK = 65536 / BN;
constexpr int MAX_PROBE = 1024 / BN;  // max-only probe tiles per CTA (see the kernel)
constexpr float DELTA_COEF = 4.2e-3f;  // > 2 * 2^-9 (+ fp32 accumulation slack)

constexpr int CAND_CAP = 12;       // candidates kept per (row, epilogue thread) between resolutions
// Split-precision variant (kernel template parameter kSplit, st3r_nn_tc_set_split).  kind::tf32 keeps 10 mantissa
// bits per operand, so smooth descriptor fields (real MASt3R maps) leave ~100 columns per row inside the error band
// and the exact re-scores dominate (7.9 ms instead of 2.5 ms per 512 x 512 pair).  With x = hi + lo + r, hi and lo
// both exactly representable in tf32 (|lo| <= 2^-10 |x|, |r| <= 2^-20 |x|), the three products
// hi_q.hi_d + hi_q.lo_d + lo_q.hi_d reproduce the fp32 dot product to ~2^-18 |q| |d| plus the accumulation error of
// the tensor core: the band narrows ~40x for 3x the tensor work (which the hand-off bound pipeline largely hides).
// hi / lo of the DB are produced once per map by nn_tc_split_launch (so the tensor core never has to convert a value
// that is not already tf32: no dependence on its truncate-or-round behaviour); hi / lo of the query tile are
// produced while it is gathered into shared memory.  Shared memory: the query tile and every DB stage double.
template <bool kSplit>
struct Lay {
  static constexpr int PARTS = kSplit ? 2 : 1;
  static constexpr int A_BYTES = BM * ROWB * PARTS;
  static constexpr int STAGE_BYTES = BN * ROWB * PARTS;
  static constexpr int STAGES = (kSplit ? 512 : 768) / BN;   // smem ring depth for DB tiles (128 KB / 96 KB)
  static constexpr int SMEM_A = 0;
  static constexpr int SMEM_B = SMEM_A + A_BYTES;
  static constexpr int SMEM_CAND = SMEM_B + STAGES * STAGE_BYTES;
  static constexpr int SMEM_CTX = SMEM_CAND + BM * CAND_CAP * 8;    // EpiCtx per query row
  static constexpr int SMEM_DBP = SMEM_CTX + BM * 24;               // DB base pointer of this CTA's problem
  static constexpr int SMEM_BAR = SMEM_DBP + 16;
  static constexpr int SMEM_TOTAL = SMEM_BAR + 512;   // up to 48 mbarriers + the TMEM base address
  static constexpr int SMEM_DYN = SMEM_TOTAL + 1024;  // slack for 1024-byte alignment
  static_assert(SMEM_DYN <= 227 * 1024, "shared memory budget of one CTA");
};
constexpr float DELTA_COEF_SPLIT = 1.0e-4f;   // 2 x (3 x 2^-20 products dropped + <= ~2e-5 fp32 accumulation), 2x margin

// ------------------------------------------------------------------ PTX wrappers
#ifdef ST3R_HOST_EMU
#include "tcgen05_emu.h"
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a CUDA error (trap) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
#ifdef NN_TC_EXP_NOFENCE   // timing experiment only
__device__ __forceinline__ void tc_fence_before() {}
__device__ __forceinline__ void tc_fence_after() {}
#else
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#endif
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_mbarrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int kCols>
__device__ __forceinline__ void tmem_alloc_cols(uint32_t smem_result_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "n"(kCols) : "memory");
}
#define tmem_alloc(addr, cols) tmem_alloc_cols<cols>(addr)
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_cols(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kCols) : "memory");
}
#define tmem_dealloc(base, cols) tmem_dealloc_cols<cols>(base)
template <int kThreads>
__device__ __forceinline__ void named_bar_sync_n() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }
#define named_bar_sync(count) named_bar_sync_n<count>()
#endif  // ST3R_HOST_EMU
__device__ __forceinline__ float fmax3(float a, float b, float c) {
#ifdef NN_TC_FMNMX3
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
#else
  return fmaxf(fmaxf(a, b), c);
#endif
}

// x -> its tf32 head (returned in x: sign, exponent and the 10 leading mantissa bits, i.e. what kind::tf32 reads) and
// the tf32 head of the remainder (lo).  x - head is exact in fp32.
__host__ __device__ __forceinline__ void split_tf32(float& x, float& lo) {
#ifdef __CUDA_ARCH__
  const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u);
#else
  union { float f; uint32_t u; } a, b;
  a.f = x; a.u &= 0xffffe000u;
  const float hi = a.f;
  b.f = x - hi; b.u &= 0xffffe000u;
  lo = b.f;
#endif
  x = hi;
}

// Maximum of 32 consecutive accumulator columns: 15 three-input max instructions + 1.
__device__ __forceinline__ float chunk_max(const float* x) {
  float m01 = fmax3(x[0], x[1], x[2]), m02 = fmax3(x[3], x[4], x[5]), m03 = fmax3(x[6], x[7], x[8]);
  float m04 = fmax3(x[9], x[10], x[11]), m05 = fmax3(x[12], x[13], x[14]), m06 = fmax3(x[15], x[16], x[17]);
  float m07 = fmax3(x[18], x[19], x[20]), m08 = fmax3(x[21], x[22], x[23]), m09 = fmax3(x[24], x[25], x[26]);
  float m10 = fmax3(x[27], x[28], x[29]), m11 = fmaxf(x[30], x[31]);
  return fmax3(fmax3(m01, m02, m03), fmax3(m04, m05, m06), fmax3(fmax3(m07, m08, m09), m10, m11));
}

// K-major, 128-byte swizzle, dense 8-row groups (SBO = 1024 B), sm_100 descriptor version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);  // start address
  d |= (uint64_t)0 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;        // stride byte offset
  d |= (uint64_t)1 << 46;                   // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;                   // layout type SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128.
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(UM >> 4) << 24);

__device__ __forceinline__ float exact_score(const float* __restrict__ q, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int k4 = 0; k4 < DK / 4; ++k4) {
    float4 qa = reinterpret_cast<const float4*>(q)[k4];
    float4 bb = reinterpret_cast<const float4*>(b)[k4];
    s = fmaf(qa.x, bb.x, s); s = fmaf(qa.y, bb.y, s); s = fmaf(qa.z, bb.z, s); s = fmaf(qa.w, bb.w, s);
  }
  return s;
}

// Diagnostic cycle counters (epilogue warp of CTA 0): [0] tiles, [1] cycles waiting for the accumulator,
// [2] cycles in the arg-max epilogue proper, [3] total cycles of the tile loop.  Read by st3r_debug_nn_tc_cycles().
__device__ unsigned long long g_nn_tc_cycles[4];

struct Cand { int j; float s; };

// Per epilogue thread state that only the cold path touches lives in shared memory, so it costs the hot loop no
// registers and the out-of-line cold function needs few arguments.
struct EpiCtx {
  const float* q;              // this row's query descriptor (global)
  unsigned long long best;     // exact best so far, packed (score, ~index): 64-bit max == best score, lowest index
  float delta;
  int resolves;                // exact list resolutions of this row in this CTA (statistics)
};
static_assert(sizeof(EpiCtx) == 24, "EpiCtx layout");

#ifdef ST3R_HOST_EMU
#define smem_raw (reinterpret_cast<uint8_t*>(emu_dyn_smem))
#else
extern __shared__ uint8_t smem_raw[];
#endif
__device__ __forceinline__ uint8_t* smem_base() {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
}

// Exact fp32 re-score of the list entries still inside the band, merged into the row's exact best.
__device__ __forceinline__ float resolve_list(const Cand* list, int cnt, float thr, EpiCtx* ctx, const float* __restrict__ DB) {
  unsigned long long best = ctx->best;
  const float* q = ctx->q;
  for (int i = 0; i < cnt; ++i) {
    const Cand c = list[i];
    if (c.s < thr) continue;
    const unsigned long long k = nn_pack(exact_score(q, DB + (size_t)c.j * DK), c.j);
    best = k > best ? k : best;
  }
  ctx->best = best;
  return best ? nn_unpack_score(best) : -INFINITY;
}

// Cold path of the epilogue for one 32-column chunk (out of line, loop-based: the hot loop must stay small enough
// for the instruction cache and keep its registers).  Appends the columns flagged in `mask` (column = col0 + bit)
// to the row's candidate list; the stored score is an upper bound of the column's approximate score (the chunk
// maximum), which keeps the later `score >= thr` filter conservative.  A full list is first compacted against the
// current threshold; if it is still full of live candidates it is resolved exactly on the spot and restarted empty,
// and the exact best raises the running bound (a later column can only win if its approximate score is
// >= best - eps, and eps <= delta / 2).  Returns (bits of the new running maximum << 32) | new count.
template <bool kSplit>
__device__ __noinline__ unsigned long long absorb_candidates(int slot, int cnt, uint32_t mask, int col0, float cmax,
                                                             float run_max) {
  constexpr int SMEM_CAND = Lay<kSplit>::SMEM_CAND, SMEM_CTX = Lay<kSplit>::SMEM_CTX, SMEM_DBP = Lay<kSplit>::SMEM_DBP;
  uint8_t* smem = smem_base();
  Cand* list = reinterpret_cast<Cand*>(smem + SMEM_CAND) + slot * CAND_CAP;
  EpiCtx* ctx = reinterpret_cast<EpiCtx*>(smem + SMEM_CTX) + slot;
  const float delta = ctx->delta;
  while (mask) {
    if (cnt == CAND_CAP) {
      const float thr = run_max - delta;
      int w = 0;
      for (int k = 0; k < CAND_CAP; ++k) {
        Cand c = list[k];
        if (c.s >= thr) list[w++] = c;
      }
      cnt = w;
      if (cnt == CAND_CAP) {
        const float bs = resolve_list(list, cnt, thr, ctx, *reinterpret_cast<const float* const*>(smem + SMEM_DBP));
        ++ctx->resolves;
        cnt = 0;
        run_max = fmaxf(run_max, bs + 0.5f * delta);
      }
    }
    const int i = __ffs(mask) - 1;
    mask &= mask - 1;
    list[cnt].j = col0 + i;
    list[cnt].s = cmax;
    ++cnt;
  }
  return ((unsigned long long)__float_as_uint(run_max) << 32) | (unsigned long long)(uint32_t)cnt;
}

// ---- cooperative variant of the cold path (kernel template parameter kCoop) -----------------------------------
// Smooth descriptor fields (real MASt3R maps, the synthetic scene) keep ~100 columns per row inside the band, so
// lists fill up ~8 times per row; resolving a list in the thread that owns it means twelve dependent L2 round trips
// while the other 31 lanes of its warp wait.  Here the whole warp enters the cold path together, lanes append without
// resolving, and full lists are resolved by the warp: the 12 entries of one row are re-scored by 12 lanes in parallel.
// It costs the random-descriptor case ~14 % (warp-wide votes on every cold tile), so the host picks the variant from
// the resolution statistics of earlier calls (st3r_nn_tc_stats / st3r_nn_tc_set_cooperative).
template <bool kSplit>
__device__ __noinline__ unsigned long long absorb_no_resolve(int slot, int cnt, uint32_t mask, int col0, float cmax, float thr) {
  Cand* list = reinterpret_cast<Cand*>(smem_base() + Lay<kSplit>::SMEM_CAND) + slot * CAND_CAP;
  while (mask) {
    if (cnt == CAND_CAP) {
      int w = 0;
      for (int k = 0; k < CAND_CAP; ++k) {
        Cand c = list[k];
        if (c.s >= thr) list[w++] = c;
      }
      cnt = w;
      if (cnt == CAND_CAP) break;           // full of live candidates: the warp resolves it
    }
    const int i = __ffs(mask) - 1;
    mask &= mask - 1;
    list[cnt].j = col0 + i;
    list[cnt].s = cmax;
    ++cnt;
  }
  return ((unsigned long long)mask << 32) | (unsigned long long)(uint32_t)cnt;
}

// Called by ALL lanes of an epilogue warp; `need` marks the lanes whose list is full of live candidates.  Returns the
// (possibly raised) running maximum of the calling lane; the resolved lists restart empty (caller resets its count).
template <bool kSplit>
__device__ __noinline__ float resolve_full_lists(int slot, bool need, float run_max) {
  constexpr int SMEM_CAND = Lay<kSplit>::SMEM_CAND, SMEM_CTX = Lay<kSplit>::SMEM_CTX, SMEM_DBP = Lay<kSplit>::SMEM_DBP;
  uint8_t* smem = smem_base();
  const Cand* lists = reinterpret_cast<const Cand*>(smem + SMEM_CAND);
  EpiCtx* ctxs = reinterpret_cast<EpiCtx*>(smem + SMEM_CTX);
  const float* DB = *reinterpret_cast<const float* const*>(smem + SMEM_DBP);
  const float delta = ctxs[slot].delta;
  const int lane = lane_id();
  uint32_t victims = __ballot_sync(0xffffffffu, need);
  __syncwarp();                                // list writes are visible to the helping lanes
  while (victims) {
    const int v = __ffs(victims) - 1;
    victims &= victims - 1;
    const int vslot = __shfl_sync(0xffffffffu, slot, v);
    const float vthr = __shfl_sync(0xffffffffu, run_max - delta, v);
    unsigned long long k = 0ull;
    if (lane < CAND_CAP) {
      const Cand c = lists[vslot * CAND_CAP + lane];
      if (c.s >= vthr) k = nn_pack(exact_score(ctxs[vslot].q, DB + (size_t)c.j * DK), c.j);
    }
#pragma unroll
    for (int off = 8; off; off >>= 1) {         // entries live in lanes 0..11: a 16-lane tree covers them
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, k, off);
      k = o > k ? o : k;
    }
    k = __shfl_sync(0xffffffffu, k, 0);
    if (lane == v) {
      unsigned long long best = ctxs[slot].best;
      best = k > best ? k : best;
      ctxs[slot].best = best;
      ++ctxs[slot].resolves;
      if (best) run_max = fmaxf(run_max, nn_unpack_score(best) + 0.5f * delta);
    }
  }
  __syncwarp();
  return run_max;
}

// Arg-max arithmetic of NCH * 32 consecutive accumulator columns held in registers (columns col_base ...; `probe`: a
// max-only probe tile).  Hot path: three-input max instructions and ONE test, straight-line; the cold paths append
// near-maximum columns to the row's candidate list (see absorb_candidates / resolve_full_lists).
template <bool kCoop, bool kSplit, int NCH>
__device__ __forceinline__ void epi_cols(float* v, const bool probe, const int col_base, const int n_end, const int slot,
                                         const float delta, float& run_max, int& cnt) {
  float cm[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) cm[c] = chunk_max(v + c * 32);
  float tmax = cm[0];
#pragma unroll
  for (int c = 1; c < NCH; ++c) tmax = fmaxf(tmax, cm[c]);
  if (probe) {
    run_max = fmaxf(run_max, tmax);          // probe tiles are always full tiles
  } else {
    const bool ragged = col_base + (NCH * 32) > n_end;
    if (kCoop) {
      if (__builtin_expect(__any_sync(0xffffffffu, ragged || tmax >= run_max - delta), 0)) {
        // Cold path, taken by the whole warp together (see resolve_full_lists).
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          float* x = v + c * 32;
          float cmax = cm[c];
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col_base + c * 32 + i >= n_end) x[i] = -INFINITY;
            cmax = chunk_max(x);
          }
          uint32_t rem = 0;                    // flagged columns still to be listed
          if (cmax >= run_max - delta) {
            run_max = fmaxf(run_max, cmax);
            const float thr = run_max - delta;
#pragma unroll
            for (int i = 0; i < 32; ++i) rem |= (x[i] >= thr ? 1u : 0u) << i;
            const unsigned long long r = absorb_no_resolve<kSplit>(slot, cnt, rem, col_base + c * 32, cmax, thr);
            cnt = (int)(uint32_t)r;
            rem = (uint32_t)(r >> 32);
          }
          while (__any_sync(0xffffffffu, rem != 0)) {      // some lane's list is full of live candidates
            run_max = resolve_full_lists<kSplit>(slot, rem != 0, run_max);
            if (rem) {
              const unsigned long long r = absorb_no_resolve<kSplit>(slot, 0, rem, col_base + c * 32, cmax, run_max - delta);
              cnt = (int)(uint32_t)r;
              rem = (uint32_t)(r >> 32);
            }
          }
        }
      }
    } else if (__builtin_expect(ragged || tmax >= run_max - delta, 0)) {
      // Cold path (a few times per row, or the single ragged tile of the DB).
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float* x = v + c * 32;
        float cmax = cm[c];
        if (ragged) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col_base + c * 32 + i >= n_end) x[i] = -INFINITY;
          cmax = chunk_max(x);
        }
        if (cmax >= run_max - delta) {
          run_max = fmaxf(run_max, cmax);
          const float thr = run_max - delta;
          uint32_t mask = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) mask |= (x[i] >= thr ? 1u : 0u) << i;
          const unsigned long long r = absorb_candidates<kSplit>(slot, cnt, mask, col_base + c * 32, cmax, run_max);
          cnt = (int)(uint32_t)r;
          run_max = __uint_as_float((uint32_t)(r >> 32));
        }
      }
    }
  }
}

// Statistics for the host's choice of variant: [0] query rows scanned, [1] exact list resolutions.
__device__ unsigned long long g_nn_tc_stats[2];

struct NnTcParams {
  CUtensorMap tmap[NN_MAX_BATCH];      // the DB (kSplit: its tf32 head, NnBatchItem::DB_hi)
  CUtensorMap tmap_lo[NN_MAX_BATCH];   // kSplit only: the tf32 tail (NnBatchItem::DB_lo)
  NnBatchItem it[NN_MAX_BATCH];
  int tiles_per_chunk;
  int dynamic;
  float delta_coef;                    // width of the candidate band relative to |q| max|db| (DELTA_COEF / DELTA_COEF_SPLIT)
};

template <bool kCoop, bool kSplit>
__global__ void __launch_bounds__(NUM_THREADS, 1)
nn_tc_kernel(const __grid_constant__ NnTcParams prm) {
  using L = Lay<kSplit>;
  constexpr int STAGES = L::STAGES, SMEM_A = L::SMEM_A, SMEM_B = L::SMEM_B, SMEM_CAND = L::SMEM_CAND,
                SMEM_CTX = L::SMEM_CTX, SMEM_DBP = L::SMEM_DBP, SMEM_BAR = L::SMEM_BAR;
  const NnBatchItem& it = prm.it[blockIdx.z];
  const CUtensorMap* tmap_db = &prm.tmap[blockIdx.z];
  const CUtensorMap* tmap_db_lo = &prm.tmap_lo[blockIdx.z];
  const float* __restrict__ Qsrc = it.Qsrc;
  const int32_t* __restrict__ qidx = it.qidx;
  const int32_t* __restrict__ count_ptr = it.count_ptr;
  const int Mmax = it.Mmax;
  const float* __restrict__ DB = it.DB;
  const int N = it.N;
  const float* __restrict__ db_norm2_max = it.db_norm_bound;
  unsigned long long* __restrict__ packed = it.packed;
  int tiles_per_chunk = prm.tiles_per_chunk;
  const int dynamic = prm.dynamic;
  const int M = count_ptr ? min(*count_ptr, Mmax) : Mmax;
  int m_tile = blockIdx.y, chunk = blockIdx.x;
  if (dynamic) {
    // The live query count is only known on the device (reciprocal-search tail iterations shrink it from
    // thousands to a handful): re-derive the (query tile, DB chunk) decomposition from it so that the
    // whole grid stays busy instead of the few CTAs a host-side split for Mmax would leave.
    const int mtiles = (M + BM - 1) / BM;
    if (mtiles == 0) return;
    const int ntiles_total = (N + BN - 1) / BN;
    int nchunks = max(1, min(ntiles_total, (int)gridDim.x / mtiles));
    tiles_per_chunk = (ntiles_total + nchunks - 1) / nchunks;
    nchunks = (ntiles_total + tiles_per_chunk - 1) / tiles_per_chunk;
    m_tile = blockIdx.x / nchunks;
    chunk = blockIdx.x - m_tile * nchunks;
    if (m_tile >= mtiles) return;
  }
  const int m0 = m_tile * BM;
  if (m0 >= M) return;
  const int n_begin = chunk * tiles_per_chunk * BN;
  if (n_begin >= N) return;
  const int n_end = min(N, n_begin + tiles_per_chunk * BN);
  const int ntiles = (n_end - n_begin + BN - 1) / BN;
  // Probe tiles.  A streaming arg-max restarts its candidate bookkeeping at every new running maximum: ~ln(columns)
  // times per row, i.e. for a 32-row warp in a quarter of all 32-column chunks.  Each CTA therefore first runs a
  // few FULL tiles spread evenly over the WHOLE DB in max-only mode: their maximum is a valid lower bound of the
  // row's final approximate maximum, which makes later "near the running maximum" hits rare (a few per row) and
  // identical in all CTAs that share the query rows.  Probe columns are seen again by the CTA that owns them.
  const int ntiles_all = (N + BN - 1) / BN;
  const int n_probe = ntiles_all >= 2 ? min(MAX_PROBE, ntiles / 8) : 0;
  const int nseq = n_probe + ntiles;
  auto seq_col = [&](int t) {
    return t < n_probe ? (int)(((long long)t * (ntiles_all - 1)) / n_probe) * BN : n_begin + (t - n_probe) * BN;
  };

  uint8_t* smem = smem_base();
  const uint32_t sA = smem_u32(smem + SMEM_A);
  const uint32_t sB = smem_u32(smem + SMEM_B);
  const uint32_t bar0 = smem_u32(smem + SMEM_BAR);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
#ifdef NN_TC_DECOUPLE
  // Experiment (scripts/build_dbg.sh, not in the default build): one full / empty barrier pair per (query tile,
  // accumulator stage) instead of per stage.  The two query tiles of a CTA then form two pipelines that only share the
  // DB tile in shared memory: a tile's epilogue starts after ITS three UMMAs (192 instead of 384 tensor cycles into the
  // round trip issue -> commit -> epilogue -> release -> issue that bounds the kernel, profiles/r01b_nn_tc_experiments.md)
  // and its next UMMAs no longer wait for the other tile's epilogue warps.
#ifdef NN_TC_EXP_NOMMA
#error "NN_TC_DECOUPLE and NN_TC_EXP_NOMMA are separate experiments"
#endif
  constexpr int NACC = MH * ACC_STAGES;
  auto tfull_bar = [&](int h, int a) { return bar0 + 8u * (2 * STAGES + h * ACC_STAGES + a); };
  auto tempty_bar = [&](int h, int a) { return bar0 + 8u * (2 * STAGES + NACC + h * ACC_STAGES + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR + 8 * (2 * STAGES + 2 * NACC));
#else
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + ACC_STAGES + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR + 8 * (2 * STAGES + 2 * ACC_STAGES));
#endif

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
#if defined(NN_TC_TWO_ISSUERS)
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MH); }   // one commit per query tile
#else
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
#endif
#ifdef NN_TC_DECOUPLE
    for (int h = 0; h < MH; ++h)
      for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(tfull_bar(h, a), 1); mbar_init(tempty_bar(h, a), EPI_WARPS / MH); }
#else
    for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
#endif
    tc_fence_mbarrier_init();
  }
  if (warp == WARP_MMA) {
    tmem_alloc(smem_u32((const void*)tmem_ptr_smem), TMEM_COLS);
    tmem_relinquish();
  }
  // A tile: gather query rows, write the 128B-swizzled K-major layout (chunk ^= row & 7), zero pad.
  for (int e = threadIdx.x; e < BM * 8; e += NUM_THREADS) {
    int r = e >> 3, c = e & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int gm = m0 + r;
    if (gm < M && c < DK / 4) {
      size_t row = qidx ? (size_t)qidx[gm] : (size_t)gm;
      v = *reinterpret_cast<const float4*>(Qsrc + row * DK + c * 4);
    }
    if (kSplit) {     // tf32 head and tail of every element (see Lay): two tiles, same swizzled layout
      float4 lo;
      split_tf32(v.x, lo.x); split_tf32(v.y, lo.y); split_tf32(v.z, lo.z); split_tf32(v.w, lo.w);
      *reinterpret_cast<float4*>(smem + SMEM_A + BM * ROWB + r * ROWB + ((c ^ (r & 7)) << 4)) = lo;
    }
    *reinterpret_cast<float4*>(smem + SMEM_A + r * ROWB + ((c ^ (r & 7)) << 4)) = v;
  }
  tc_fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == WARP_TMA) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int t = 0; t < nseq; ++t) {
        int s = t % STAGES;
        uint32_t ph = (uint32_t)(t / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), L::STAGE_BYTES);
        tma_load_2d(sB + s * L::STAGE_BYTES, tmap_db, 0, seq_col(t), full_bar(s));
        if (kSplit) tma_load_2d(sB + s * L::STAGE_BYTES + BN * ROWB, tmap_db_lo, 0, seq_col(t), full_bar(s));
      }
    }
#ifdef NN_TC_TWO_ISSUERS
  } else if (warp >= WARP_MMA) {
    // ===== one UMMA issuer per query tile: the two accumulator pipelines only share the DB tile in shared memory =====
    if (lane == 0) {
      const int h = warp - WARP_MMA;
      for (int t = 0; t < nseq; ++t) {
        const int s = t % STAGES, a = t % ACC_STAGES;
        mbar_wait(tempty_bar(h, a), ((uint32_t)(t / ACC_STAGES) & 1u) ^ 1u);
        mbar_wait(full_bar(s), (uint32_t)(t / STAGES) & 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)((h * ACC_STAGES + a) * BN);
#pragma unroll
        for (int part = 0; part < (kSplit ? 3 : 1); ++part) {      // hi_q . hi_d (+ hi_q . lo_d + lo_q . hi_d, see Lay)
#pragma unroll
          for (int kk = 0; kk < DK / 8; ++kk) {
            const uint64_t da = make_smem_desc(sA + (part == 2 ? BM * ROWB : 0) + h * UM * ROWB + kk * 32);
            const uint64_t db = make_smem_desc(sB + s * L::STAGE_BYTES + (part == 1 ? BN * ROWB : 0) + kk * 32);
            tc_mma_tf32(d_tmem, da, db, IDESC, (part > 0 || kk > 0) ? 1u : 0u);
          }
        }
        tc_commit(empty_bar(s));            // (one arrival per query tile: the stage is free when both have retired)
        tc_commit(tfull_bar(h, a));
      }
    }
#endif
  } else if (warp == WARP_MMA) {
    // ===== MMA issuer =====
    if (lane == 0) {
      for (int t = 0; t < nseq; ++t) {
        int s = t % STAGES;
        uint32_t ph = (uint32_t)(t / STAGES) & 1u;
        int a = t % ACC_STAGES;
        uint32_t aph = (uint32_t)(t / ACC_STAGES) & 1u;
#ifdef NN_TC_DECOUPLE
        mbar_wait(full_bar(s), ph);
#else
        mbar_wait(tempty_bar(a), aph ^ 1u);
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
#endif
#ifdef NN_TC_EXP_NOMMA   // timing experiment only (wrong results): TMA stream + barrier hand-offs, no tensor work
        mbar_arrive(empty_bar(s));
        mbar_arrive(tfull_bar(a));
        continue;
#endif
#pragma unroll
        for (int h = 0; h < MH; ++h) {
#ifdef NN_TC_DECOUPLE
          mbar_wait(tempty_bar(h, a), aph ^ 1u);
          tc_fence_after();
#endif
          const uint32_t d_tmem = tmem_base + (uint32_t)((h * ACC_STAGES + a) * BN);
#pragma unroll
          for (int kk = 0; kk < DK / 8; ++kk) {
            uint64_t da = make_smem_desc(sA + h * UM * ROWB + kk * 32);
            uint64_t db = make_smem_desc(sB + s * L::STAGE_BYTES + kk * 32);
            tc_mma_tf32(d_tmem, da, db, IDESC, kk > 0 ? 1u : 0u);
          }
          if (kSplit) {   // + hi_q . lo_d + lo_q . hi_d into the same accumulator
#pragma unroll
            for (int kk = 0; kk < DK / 8; ++kk) {
              uint64_t da = make_smem_desc(sA + h * UM * ROWB + kk * 32);
              uint64_t db = make_smem_desc(sB + s * L::STAGE_BYTES + BN * ROWB + kk * 32);
              tc_mma_tf32(d_tmem, da, db, IDESC, 1u);
            }
#pragma unroll
            for (int kk = 0; kk < DK / 8; ++kk) {
              uint64_t da = make_smem_desc(sA + BM * ROWB + h * UM * ROWB + kk * 32);
              uint64_t db = make_smem_desc(sB + s * L::STAGE_BYTES + kk * 32);
              tc_mma_tf32(d_tmem, da, db, IDESC, 1u);
            }
          }
#ifdef NN_TC_DECOUPLE
          tc_commit(tfull_bar(h, a));   // this query tile's accumulator is ready for its epilogue warps
#endif
        }
        tc_commit(empty_bar(s));   // smem stage free once these MMAs retire
#ifndef NN_TC_DECOUPLE
        tc_commit(tfull_bar(a));   // accumulator ready for the epilogue
#endif
      }
    }
  } else {
    // ===== epilogue: fused arg-max over TMEM accumulators =====
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = warp >> 2;              // which of the CTA's query tiles
    const int row = half * UM + quarter * 32 + lane;   // query row inside the CTA; its accumulator is TMEM lane row % 128
    const int gm = m0 + row;
    const bool row_ok = gm < M;
    const int slot = row;                    // candidate list / EpiCtx of this thread
    int cnt = 0;
    const float* qrow = Qsrc;
    float delta = 0.f;
    if (row_ok) {
      qrow = Qsrc + (qidx ? (size_t)qidx[gm] : (size_t)gm) * DK;
      float n2 = 0.f;
#pragma unroll
      for (int k = 0; k < DK; ++k) n2 = fmaf(qrow[k], qrow[k], n2);
      delta = prm.delta_coef * sqrtf(n2) * sqrtf(*db_norm2_max) + 1e-30f;
    }
    float run_max = row_ok ? -INFINITY : INFINITY;  // padded rows never trigger
    {
      EpiCtx* ctx = reinterpret_cast<EpiCtx*>(smem + SMEM_CTX) + slot;
      ctx->q = qrow; ctx->best = 0ull; ctx->delta = delta; ctx->resolves = 0;
      if (threadIdx.x == 0) *reinterpret_cast<const float**>(smem + SMEM_DBP) = DB;
      named_bar_sync(EPI_THREADS);                                                       // DB pointer visible to all
    }

#ifdef NN_TC_DEBUG_CYCLES
    const bool dbg = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 0 && lane == 0);
    unsigned long long c_wait = 0, c_epi = 0;
    const long long c_start = clock64();
#endif
    for (int t = 0; t < nseq; ++t) {
      int a = t % ACC_STAGES;
      uint32_t aph = (uint32_t)(t / ACC_STAGES) & 1u;
#ifdef NN_TC_DEBUG_CYCLES
      const long long c0 = clock64();
#endif
#ifdef NN_TC_DECOUPLE
      mbar_wait(tfull_bar(half, a), aph);
#else
      mbar_wait(tfull_bar(a), aph);
#endif
      tc_fence_after();
#ifdef NN_TC_DEBUG_CYCLES
      const long long c1 = clock64();
#endif
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((half * ACC_STAGES + a) * BN);
      // Pull this thread's accumulator columns into registers with back-to-back tcgen05.ld (their latencies
      // overlap), then hand the TMEM stage back to the MMA warp BEFORE the arg-max arithmetic so tile t+2 can
      // start immediately.
      float v[EPI_COLS];
#ifdef NN_TC_EXP_NOLD   // timing experiment only (wrong results): no TMEM read
#pragma unroll
      for (int c = 0; c < EPI_COLS; ++c) v[c] = __int_as_float(taddr + c + t);
#else
#pragma unroll
      for (int c = 0; c < EPI_CHUNKS; ++c) tc_ld32(taddr + c * 32, v + c * 32);
      tc_wait_ld();
#endif
#ifdef NN_TC_DEBUG_CYCLES
#ifdef NN_TC_DEBUG_LDONLY   // (tcgen05.wait::ld compiles to scoreboard waits at the consumers: touch the last register of every load)
      {
        const float sink = (v[31] + v[EPI_COLS / 2 - 1]) + (v[EPI_COLS - 33] + v[EPI_COLS - 1]);
        *reinterpret_cast<volatile float*>(smem + SMEM_DBP + 8) = sink;      // (in-order issue: the store waits for the data)
      }
#endif
      const long long c2 = clock64();      // accumulator columns in registers
#endif
      tc_fence_before();
#if defined(NN_TC_DECOUPLE)
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(half, a));
#else
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(a));      // one arrival per warp
#endif
#ifdef NN_TC_EXP_NOALU  // timing experiment only (wrong results): no arg-max arithmetic
      if (v[t & (EPI_COLS - 1)] == 123.456f) run_max = v[5];
      continue;
#endif
      epi_cols<kCoop, kSplit, EPI_CHUNKS>(v, t < n_probe, n_begin + (t - n_probe) * BN, n_end, slot, delta, run_max, cnt);
#ifdef NN_TC_DEBUG_CYCLES
      c_wait += (unsigned long long)(c1 - c0);
#ifdef NN_TC_DEBUG_LDONLY     // slot [2] = the TMEM read alone (issue of the four tcgen05.ld .. wait::ld)
      c_epi += (unsigned long long)(c2 - c1);
#else
      c_epi += (unsigned long long)(clock64() - c1);
#endif
#endif
    }
#ifdef NN_TC_DEBUG_CYCLES
    if (dbg) {
      atomicAdd(&g_nn_tc_cycles[0], (unsigned long long)nseq);
      atomicAdd(&g_nn_tc_cycles[1], c_wait);
      atomicAdd(&g_nn_tc_cycles[2], c_epi);
      atomicAdd(&g_nn_tc_cycles[3], (unsigned long long)(clock64() - c_start));
    }
#endif

    if (row_ok) {
      EpiCtx* ctx = reinterpret_cast<EpiCtx*>(smem + SMEM_CTX) + slot;
      resolve_list(reinterpret_cast<Cand*>(smem + SMEM_CAND) + slot * CAND_CAP, cnt, run_max - delta, ctx, DB);
      if (ctx->best != 0ull) atomicMax(packed + gm, ctx->best);
    }
    {   // statistics: one pair of atomics per warp
      EpiCtx* ctx = reinterpret_cast<EpiCtx*>(smem + SMEM_CTX) + slot;
      int nres = row_ok ? ctx->resolves : 0, nrow = (row_ok && n_begin == 0) ? 1 : 0;   // each query row counted once
      for (int off = 16; off; off >>= 1) {
        nres += __shfl_xor_sync(0xffffffffu, nres, off);
        nrow += __shfl_xor_sync(0xffffffffu, nrow, off);
      }
      if (lane == 0) {
        if (nrow) atomicAdd(&g_nn_tc_stats[0], (unsigned long long)nrow);
        if (nres) atomicAdd(&g_nn_tc_stats[1], (unsigned long long)nres);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__global__ void db_norm_kernel(const float* __restrict__ DB, int N, int d, uint32_t* __restrict__ out_bits) {
  float m = 0.f;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    const float* b = DB + (size_t)j * d;
    float n2 = 0.f;
    for (int k = 0; k < d; ++k) n2 = fmaf(b[k], b[k], n2);
    m = fmaxf(m, n2);
  }
  for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if (lane_id() == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order like uints
}

// DB [N, d] -> its tf32 head and tail (same shape), see Lay.  One pass per map and API call, next to db_norm_kernel.
__global__ void db_split_kernel(const float4* __restrict__ DB, size_t n4, float4* __restrict__ hi, float4* __restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = DB[i], l;
    split_tf32(v.x, l.x); split_tf32(v.y, l.y); split_tf32(v.z, l.z); split_tf32(v.w, l.w);
    hi[i] = v;
    lo[i] = l;
  }
}

#ifndef ST3R_HOST_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

#endif

}  // namespace

bool nn_tc_supported(int d) { return d == DK; }

static int g_nn_tc_coop = 0;
static int g_nn_tc_split = 0;
static float g_nn_tc_delta_coef[2] = {0.f, 0.f};   // diagnostic overrides (0: DELTA_COEF / DELTA_COEF_SPLIT)

// Diagnostic (scripts/nn_split_margin.py): overrides the width of the candidate band of the plain / split-precision
// kernel for subsequent launches, 0 restores the built-in value.  With a band narrower than the tensor core's real
// error the result is no longer guaranteed exact - that is what the script measures.
extern "C" int st3r_debug_nn_tc_set_delta_coef(float plain, float split) {
  ST3R_CHECK_ARG(plain >= 0.f && split >= 0.f, "st3r_debug_nn_tc_set_delta_coef: negative coefficient");
  g_nn_tc_delta_coef[0] = plain;
  g_nn_tc_delta_coef[1] = split;
  return ST3R_OK;
}

bool nn_tc_split_enabled() { return g_nn_tc_split != 0; }

extern "C" int st3r_nn_tc_set_split(int on) {
  g_nn_tc_split = on ? 1 : 0;
  return ST3R_OK;
}

int nn_tc_split_launch(const float* DB, int N, int d, float* hi, float* lo, cudaStream_t stream) {
  ST3R_CHECK_ARG(d == DK, "nn_tc: the split-precision variant needs d == 24");
  ST3R_CHECK_ARG(DB && hi && lo && ((uintptr_t)DB % 16) == 0 && ((uintptr_t)hi % 16) == 0 && ((uintptr_t)lo % 16) == 0,
                 "nn_tc_split: operands must be non-null and 16-byte aligned");
  if (N <= 0) return ST3R_OK;
  const size_t n4 = (size_t)N * DK / 4;
  const int blocks = (int)min((n4 + 255) / 256, (size_t)st3r_num_sms() * 8);
  db_split_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(DB), n4, reinterpret_cast<float4*>(hi),
                                              reinterpret_cast<float4*>(lo));
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

extern "C" int st3r_nn_tc_set_cooperative(int on) {
  g_nn_tc_coop = on ? 1 : 0;
  return ST3R_OK;
}

extern "C" int st3r_nn_tc_stats(unsigned long long* h_out2, int reset) {
  if (h_out2) ST3R_CHECK_CUDA(cudaMemcpyFromSymbol(h_out2, g_nn_tc_stats, sizeof(g_nn_tc_stats)));
  if (reset) {
    unsigned long long z[2] = {0, 0};
    ST3R_CHECK_CUDA(cudaMemcpyToSymbol(g_nn_tc_stats, z, sizeof(z)));
  }
  return ST3R_OK;
}

extern "C" int st3r_debug_nn_tc_cycles(unsigned long long* h_out4, int reset) {
  if (h_out4) ST3R_CHECK_CUDA(cudaMemcpyFromSymbol(h_out4, g_nn_tc_cycles, sizeof(g_nn_tc_cycles)));
  if (reset) {
    unsigned long long z[4] = {0, 0, 0, 0};
    ST3R_CHECK_CUDA(cudaMemcpyToSymbol(g_nn_tc_cycles, z, sizeof(z)));
  }
  return ST3R_OK;
}

int nn_db_norm_launch(const float* DB, int N, int d, float* out_bound, cudaStream_t stream) {
  ST3R_CHECK_CUDA(cudaMemsetAsync(out_bound, 0, sizeof(float), stream));
  if (N <= 0) return ST3R_OK;
  int blocks = min((N + 255) / 256, st3r_num_sms() * 4);
  db_norm_kernel<<<blocks, 256, 0, stream>>>(DB, N, d, reinterpret_cast<uint32_t*>(out_bound));
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

#ifdef ST3R_HOST_EMU
static int encode_db_tmap(CUtensorMap* tmap, const float* DB, int N) {     // the software model's tensor map
  tmap->base = DB; tmap->rows = N; tmap->cols = DK; tmap->box_cols = 32; tmap->box_rows = BN;
  return ST3R_OK;
}
#else
static int encode_db_tmap(CUtensorMap* tmap, const float* DB, int N) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) {
    st3r_set_error("nn_tc: cuTensorMapEncodeTiled is unavailable in this driver");
    return ST3R_ERR_CUDA;
  }
  // DB viewed as a [N, 24] fp32 tensor; the box is [128 rows, 32 floats]: the 8 out-of-bounds
  // floats per row are zero-filled by TMA, which pads K to the 128-byte swizzle span for free.
  cuuint64_t gdim[2] = {(cuuint64_t)DK, (cuuint64_t)N};
  cuuint64_t gstride[1] = {(cuuint64_t)DK * sizeof(float)};
  cuuint32_t box[2] = {32, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(DB), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    st3r_set_error("nn_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return ST3R_ERR_CUDA;
  }
  return ST3R_OK;
}

#endif

int nn_tc_launch_batch(const NnBatchItem* items, int n, int d, cudaStream_t stream) {
  ST3R_CHECK_ARG(n >= 1 && n <= NN_MAX_BATCH, "nn_tc: batch size must be 1..%d", NN_MAX_BATCH);
  ST3R_CHECK_ARG(d == DK, "nn_tc: only d == 24 is supported (got %d)", d);
  NnTcParams prm;
  int m = 0, max_M = 0;
  bool all_counts = true;
  // the split-precision variant runs when it is switched on AND the caller prepared the tf32 head / tail of every DB
  bool split = g_nn_tc_split != 0;
  for (int i = 0; i < n; ++i)
    if (items[i].Mmax > 0 && items[i].N > 0 && !(items[i].DB_hi && items[i].DB_lo)) split = false;
  for (int i = 0; i < n; ++i) {
    const NnBatchItem& it = items[i];
    if (it.Mmax <= 0 || it.N <= 0) continue;
    ST3R_CHECK_ARG(((uintptr_t)it.DB % 16) == 0 && ((uintptr_t)it.Qsrc % 16) == 0, "nn_tc: operands must be 16-byte aligned");
    ST3R_CHECK_ARG(it.db_norm_bound && it.packed, "nn_tc: missing scratch");
    int rc;
    if (split) {
      ST3R_CHECK_ARG(((uintptr_t)it.DB_hi % 16) == 0 && ((uintptr_t)it.DB_lo % 16) == 0, "nn_tc: split arrays must be 16-byte aligned");
      rc = encode_db_tmap(&prm.tmap[m], it.DB_hi, it.N);
      if (rc) return rc;
      rc = encode_db_tmap(&prm.tmap_lo[m], it.DB_lo, it.N);
    } else {
      rc = encode_db_tmap(&prm.tmap[m], it.DB, it.N);
      prm.tmap_lo[m] = prm.tmap[m];
    }
    if (rc) return rc;
    prm.it[m] = it;
    max_M = max(max_M, it.Mmax);
    all_counts = all_counts && it.count_ptr != nullptr;
    ++m;
  }
  if (m == 0) return ST3R_OK;
  prm.delta_coef = split ? (g_nn_tc_delta_coef[1] > 0.f ? g_nn_tc_delta_coef[1] : DELTA_COEF_SPLIT)
                         : (g_nn_tc_delta_coef[0] > 0.f ? g_nn_tc_delta_coef[0] : DELTA_COEF);
  for (int i = m; i < NN_MAX_BATCH; ++i) { prm.tmap[i] = prm.tmap[0]; prm.tmap_lo[i] = prm.tmap_lo[0]; prm.it[i] = prm.it[0]; }
  static PerDeviceOnce attr_set;
  if (!attr_set.done()) {
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(nn_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<false>::SMEM_DYN));
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(nn_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<false>::SMEM_DYN));
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(nn_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<true>::SMEM_DYN));
    ST3R_CHECK_CUDA(cudaFuncSetAttribute(nn_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<true>::SMEM_DYN));
    attr_set.mark();
  }
  auto launch = [&](dim3 grid, const NnTcParams& q) {
    if (split) {
      if (g_nn_tc_coop) nn_tc_kernel<true, true><<<grid, NUM_THREADS, Lay<true>::SMEM_DYN, stream>>>(q);
      else nn_tc_kernel<false, true><<<grid, NUM_THREADS, Lay<true>::SMEM_DYN, stream>>>(q);
    } else {
      if (g_nn_tc_coop) nn_tc_kernel<true, false><<<grid, NUM_THREADS, Lay<false>::SMEM_DYN, stream>>>(q);
      else nn_tc_kernel<false, false><<<grid, NUM_THREADS, Lay<false>::SMEM_DYN, stream>>>(q);
    }
  };
  const int slots = st3r_num_sms();      // one CTA per SM (it owns the whole tensor memory)
  const int mtiles = (max_M + BM - 1) / BM;
  if (all_counts && mtiles <= slots) {
    // device-side decomposition (see the kernel): one wave of `slots` CTAs per problem
    prm.tiles_per_chunk = 0;
    prm.dynamic = 1;
    launch(dim3(slots, 1, m), prm);
    ST3R_CHECK_LAUNCH();
  } else {
    // host-side decomposition, one problem per launch.  Chunk the DB so that the grid is (close to) a whole number
    // of waves of one CTA per SM: long chunks keep the per-row candidate restarts rare, and a partial last wave
    // would idle most of the chip.
    for (int i = 0; i < m; ++i) {
      NnTcParams one = prm;
      one.tmap[0] = prm.tmap[i];
      one.tmap_lo[0] = prm.tmap_lo[i];
      one.it[0] = prm.it[i];
      const int mt = (prm.it[i].Mmax + BM - 1) / BM;
      const int ntiles_total = (prm.it[i].N + BN - 1) / BN;
      int waves = max(1, (mt * ((ntiles_total + MAX_TILES_PER_CHUNK - 1) / MAX_TILES_PER_CHUNK) + slots - 1) / slots);
      int nchunks = max(1, min(ntiles_total, (waves * slots) / mt));
      int tiles_per_chunk = (ntiles_total + nchunks - 1) / nchunks;
      nchunks = (ntiles_total + tiles_per_chunk - 1) / tiles_per_chunk;
      one.tiles_per_chunk = tiles_per_chunk;
      one.dynamic = 0;
      launch(dim3(nchunks, mt, 1), one);
      ST3R_CHECK_LAUNCH();
    }
  }
  return ST3R_OK;
}

int nn_tc_launch(const float* Qsrc, const int32_t* qidx, const int32_t* count_ptr, int Mmax, const float* DB,
                 int N, int d, const float* db_norm_bound, unsigned long long* packed, cudaStream_t stream,
                 const float* DB_hi, const float* DB_lo) {
  if (Mmax <= 0 || N <= 0) return ST3R_OK;
  NnBatchItem it{Qsrc, qidx, count_ptr, Mmax, DB, N, db_norm_bound, packed, DB_hi, DB_lo};
  return nn_tc_launch_batch(&it, 1, d, stream);
}
