// TEST-ONLY: a software model of the Blackwell pieces nn_tc.cu programs through PTX - mbarriers with transaction
// counts, TMA 2-D tile loads with the 128-byte swizzle and out-of-bounds zero fill, tcgen05.mma kind::tf32 reading
// swizzled K-major shared-memory operands through 64-bit matrix descriptors into tensor memory, tcgen05.ld /
// commit / alloc - so that the kernel's OWN SOURCE (roles, barrier protocol with its phases and parities, tile and
// column indexing, descriptor arithmetic, the candidate / exact re-score epilogue) runs on the SIMT emulator.
// Arithmetic model: operands truncated to tf32 (sign, exponent, 10 mantissa bits), products and sums in fp32 in
// ascending k.  The real tensor core's summation order / rounding is NOT modelled: exactness of the kernel's results
// rests on its candidate band, which tests/test_match_gpu.py checks on the hardware.  Asynchrony is not modelled
// either: a TMA load or an MMA completes on the spot.
#pragma once
// (<map> comes from emu_cuda_shim.h: this header is included inside nn_tc.cu's anonymous namespace)

struct CUtensorMap { const float* base; int rows, cols; int box_cols, box_rows; char pad[128 - 24]; };
static_assert(sizeof(CUtensorMap) == 128, "same size as the driver's opaque struct");

namespace tc_emu {
struct MBar { uint32_t count = 0; int pending = 0; long tx = 0; uint32_t phase = 0; };
static std::map<uint32_t, MBar> g_bars;          // keyed by shared-memory address
static float g_tmem[128][512];
static inline char* smem_ptr(uint32_t a) { return emu_dyn_smem + a; }
static inline uint32_t swz(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }      // SWIZZLE_128B on address bits
static inline void progress() { ++emu::g_cta->progress; }
static inline void maybe_flip(MBar& b) {
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = (int)b.count; }
}
static inline float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
}  // namespace tc_emu

static inline uint32_t smem_u32(const void* p) { return (uint32_t)((const char*)p - emu_dyn_smem); }
static inline void mbar_init(uint32_t bar, uint32_t count) {
  tc_emu::MBar& b = tc_emu::g_bars[bar];
  b = tc_emu::MBar();
  b.count = count; b.pending = (int)count;
}
static inline void mbar_arrive(uint32_t bar) {
  tc_emu::MBar& b = tc_emu::g_bars.at(bar);
  --b.pending;
  tc_emu::maybe_flip(b);
  tc_emu::progress();
}
static inline void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  tc_emu::MBar& b = tc_emu::g_bars.at(bar);
  b.tx += bytes;
  --b.pending;
  tc_emu::maybe_flip(b);
  tc_emu::progress();
}
static inline bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  const bool done = tc_emu::g_bars.at(bar).phase != parity;     // the phase with this parity has completed
  if (!done) emu::yield();
  return done;
}
static inline void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
static inline void tma_load_2d(uint32_t smem_dst, const CUtensorMap* t, int c0, int c1, uint32_t bar) {
  for (int r = 0; r < t->box_rows; ++r)
    for (int c = 0; c < t->box_cols; ++c) {
      const int row = c1 + r, col = c0 + c;
      const float v = (row >= 0 && row < t->rows && col >= 0 && col < t->cols) ? t->base[(size_t)row * t->cols + col] : 0.f;
      memcpy(tc_emu::smem_ptr(tc_emu::swz(smem_dst + (uint32_t)(r * t->box_cols + c) * 4u)), &v, 4);
    }
  tc_emu::MBar& b = tc_emu::g_bars.at(bar);
  b.tx -= (long)t->box_rows * t->box_cols * 4;
  tc_emu::maybe_flip(b);
  tc_emu::progress();
}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void tc_commit(uint32_t bar) { mbar_arrive(bar); }      // the MMAs issued so far have already completed
// D[128 x N] (+)= A[128 x 8] B[N x 8]^T; A / B: K-major rows of 128 bytes, 8-row groups 1024 bytes apart, swizzled.
static inline void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const uint32_t a0 = (uint32_t)(desc_a & 0x3fffu) << 4, b0 = (uint32_t)(desc_b & 0x3fffu) << 4;
  const int N = (int)((idesc >> 17) & 0x3fu) << 3, M = (int)((idesc >> 24) & 0x1fu) << 4;
  if (M != 128 || ((desc_a >> 61) & 7) != 2 || ((desc_b >> 61) & 7) != 2 || ((desc_a >> 32) & 0x3fff) != (1024 >> 4)) abort();
  const int col0 = (int)(tmem_d & 0xffffu), lane0 = (int)(tmem_d >> 16);
  auto elem = [](uint32_t base, int r, int k) {
    float v;
    memcpy(&v, tc_emu::smem_ptr(tc_emu::swz(base + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)k * 4u)), 4);
    return tc_emu::tf32(v);
  };
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < N; ++c) {
      float s = accumulate ? tc_emu::g_tmem[lane0 + r][col0 + c] : 0.f;
      for (int k = 0; k < 8; ++k) s += elem(a0, r, k) * elem(b0, c, k);
      tc_emu::g_tmem[lane0 + r][col0 + c] = s;
    }
}
static inline void tc_ld32(uint32_t taddr, float* v) {      // 32x32b.x32: lane l of the warp reads TMEM lane base + l
  const int lane = (int)(taddr >> 16) + lane_id(), col = (int)(taddr & 0xffffu);
  for (int c = 0; c < 32; ++c) v[c] = tc_emu::g_tmem[lane][col + c];
}
static inline void tc_wait_ld() {}
static inline void tc_fence_mbarrier_init() {}
static inline void tc_fence_proxy_async() {}
static inline void tmem_alloc(uint32_t smem_result_addr, int /*cols*/) { const uint32_t base = 0; memcpy(tc_emu::smem_ptr(smem_result_addr), &base, 4); }
static inline void tmem_relinquish() {}
static inline void tmem_dealloc(uint32_t, int) {}
// named barrier 1 among `count` threads (the epilogue warps)
namespace tc_emu { static emu::BlockBarrier g_named; }
static inline void named_bar_sync(int count) { emu::barrier_wait(tc_emu::g_named, count, 0); }
static inline void __trap() { abort(); }
