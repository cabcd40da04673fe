// TEST-ONLY: a small SIMT emulator that runs the SOURCE of a CUDA kernel on the host, one CTA at a time, every CUDA
// thread as a cooperatively scheduled fiber of ONE OS thread.  Block barriers and warp collectives are
// rendezvous points: a fiber that reaches one yields until all its peers have arrived, so shuffles, votes, redux and
// __syncthreads have their CUDA meaning, a missing peer shows up as a reported deadlock instead of a hang, and the
// kernel's indexing / queueing / reduction logic is exercised exactly as written.  What it does NOT model: memory
// consistency (one OS thread: every store is immediately visible), timing, bank conflicts, divergent collectives
// (all 32 lanes of a warp must execute the same sequence of collectives - true of the kernels tested with it).
// Used by tests/host/raster_emu_host.cpp; never part of the product library.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <functional>
#include <vector>

struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
// Shared variables are function-local statics in a dedicated section: one CTA runs at a time (nothing is re-zeroed:
// kernels must initialise what they read); a CLUSTER run keeps one copy of the section per CTA and swaps it in and out
// when the scheduler moves from one CTA of the cluster to the next, which is also what map_shared_rank() resolves into.
#define __shared__ static __attribute__((section("emu_shared")))
extern "C" char __start_emu_shared[], __stop_emu_shared[];

namespace emu {

constexpr int WARP = 32;
constexpr size_t STACK = 256 * 1024;

// Context switch.  glibc's swapcontext saves and restores the signal mask with a system call per switch, and a launch
// sequence makes millions of switches; on x86-64 a fiber context is just a saved stack pointer (the callee-saved
// registers are pushed on the fiber's own stack), elsewhere ucontext is used.
#if defined(__x86_64__)
struct Context { void* sp = nullptr; };
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.weak emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
static inline void context_switch(Context& from, Context& to) { emu_switch(&from.sp, to.sp); }
static inline void context_init(Context& c, char* stack, size_t bytes, void (*entry)()) {
  // initial frame: six zeroed callee-saved registers, then the entry point as return address; the stack pointer is
  // 16-byte aligned + 8 at the entry's first instruction, as after a call
  uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
  void** sp = (void**)(top - 8);
  *--sp = (void*)entry;
  for (int i = 0; i < 6; ++i) *--sp = nullptr;
  c.sp = sp;
}
#else
struct Context { ucontext_t uc; };
static inline void context_switch(Context& from, Context& to) { swapcontext(&from.uc, &to.uc); }
static inline void context_init(Context& c, char* stack, size_t bytes, void (*entry)()) {
  getcontext(&c.uc);
  c.uc.uc_stack.ss_sp = stack;
  c.uc.uc_stack.ss_size = bytes;
  c.uc.uc_link = nullptr;
  makecontext(&c.uc, entry, 0);
}
#endif

struct Fiber {
  Context ctx;
  char* stack = nullptr;   // malloc'ed, untouched until used
  uint3 tid;
  int cta = 0;             // rank of the fiber's CTA inside its cluster
  bool done = false;
};

struct WarpState {
  uint64_t slot[WARP];     // values published by the lanes for the collective in flight
  int arrived = 0;         // lanes that published
  int departed = 0;        // lanes that consumed
  unsigned generation = 0;
};

struct BlockBarrier {
  int arrived = 0;
  unsigned generation = 0;
  int count_true = 0, count_result = 0;
};

struct Cta {               // one CTA, or the CTAs of one cluster (threads_per_cta fibers each)
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  std::vector<BlockBarrier> bars;      // one per CTA
  BlockBarrier cluster_bar;
  int threads_per_cta = 0, n_ctas = 1, live_cta = 0;
  std::vector<std::vector<char>> shared_copy;   // cluster runs: the shared-memory section of every CTA
  Context sched;
  int cur = -1;
  long progress = 0;       // arrivals, departures and exits: a scheduler round without any is a deadlock
};

static Cta* g_cta = nullptr;
static uint3 g_blockIdx;          // of CTA 0 of the running cluster; blockIdx.x adds the fiber's rank
static dim3 g_blockDim, g_gridDim;
static std::function<void()> g_body;

static inline Fiber& self() { return g_cta->fibers[g_cta->cur]; }
static inline void yield() { context_switch(self().ctx, g_cta->sched); }

static void trampoline() {
  g_body();
  self().done = true;
  ++g_cta->progress;
  context_switch(self().ctx, g_cta->sched);
  abort();     // a finished fiber is never resumed
}

// fiber stacks are recycled between launches (a launch sequence starts thousands of CTAs)
static std::vector<char*> g_stack_pool;
static size_t g_stack_pool_bytes = 0;
static char* take_stack(size_t bytes) {
  if (bytes != g_stack_pool_bytes) {
    for (char* p : g_stack_pool) free(p);
    g_stack_pool.clear();
    g_stack_pool_bytes = bytes;
  }
  if (g_stack_pool.empty()) return (char*)malloc(bytes);
  char* p = g_stack_pool.back();
  g_stack_pool.pop_back();
  return p;
}

static inline size_t shared_bytes() { return (size_t)(__stop_emu_shared - __start_emu_shared); }
static void switch_cta(Cta& c, int next) {      // cluster runs: park the live CTA's shared memory, bring in the next one's
  if (c.n_ctas == 1 || next == c.live_cta) return;
  memcpy(c.shared_copy[c.live_cta].data(), __start_emu_shared, shared_bytes());
  memcpy(__start_emu_shared, c.shared_copy[next].data(), shared_bytes());
  c.live_cta = next;
}

// Runs `body` once per thread of a cluster of `n_ctas` CTAs of `nthreads` threads (n_ctas = 1: a plain CTA).
// Returns false on deadlock (no fiber can progress).
static bool run_cluster(int n_ctas, int nthreads, const std::function<void()>& body, size_t stack_bytes = STACK) {
  Cta cta;
  const int total = n_ctas * nthreads;
  cta.fibers.resize(total);
  cta.warps.resize((size_t)n_ctas * ((nthreads + WARP - 1) / WARP));
  cta.bars.resize(n_ctas);
  cta.threads_per_cta = nthreads;
  cta.n_ctas = n_ctas;
  if (n_ctas > 1) cta.shared_copy.assign(n_ctas, std::vector<char>(shared_bytes(), 0));
  g_cta = &cta;
  g_body = body;
  for (int t = 0; t < total; ++t) {
    Fiber& f = cta.fibers[t];
    f.stack = take_stack(stack_bytes);
    f.tid = uint3{(unsigned)(t % nthreads), 0, 0};
    f.cta = t / nthreads;
    context_init(f.ctx, f.stack, stack_bytes, trampoline);
  }
  auto release = [&]() { for (Fiber& f : cta.fibers) g_stack_pool.push_back(f.stack); g_cta = nullptr; };
  int alive = total;
  while (alive > 0) {
    alive = 0;
    const long before = cta.progress;
    for (int t = 0; t < total; ++t) {
      if (cta.fibers[t].done) continue;
      ++alive;
      switch_cta(cta, cta.fibers[t].cta);
      cta.cur = t;
      context_switch(cta.sched, cta.fibers[t].ctx);
    }
    // fibers only block at rendezvous points: a round in which nobody arrived, departed or finished cannot be followed
    // by a better one
    if (alive > 0 && cta.progress == before) {
      fprintf(stderr, "simt_emu: deadlock (%d fibers blocked at a barrier / warp collective)\n", alive);
      release();
      return false;
    }
  }
  release();
  return true;
}
static bool run_cta(int nthreads, const std::function<void()>& body) { return run_cluster(1, nthreads, body); }

// ---- block barrier (of the calling fiber's CTA) and cluster barrier
static inline int barrier_wait(BlockBarrier& b, int parties, int pred) {
  Cta& c = *g_cta;
  const unsigned gen = b.generation;
  b.count_true += pred ? 1 : 0;
  ++c.progress;
  if (++b.arrived == parties) {
    b.arrived = 0;
    b.count_result = b.count_true;
    b.count_true = 0;
    ++b.generation;
  } else {
    while (b.generation == gen) yield();
  }
  return b.count_result;
}
static inline int syncthreads_count(int pred) {
  Cta& c = *g_cta;
  return barrier_wait(c.bars[self().cta], c.threads_per_cta, pred);
}
static inline void cluster_sync() {
  Cta& c = *g_cta;
  barrier_wait(c.cluster_bar, c.threads_per_cta * c.n_ctas, 0);
}
// Address of a shared variable in CTA `rank` of the cluster (distributed shared memory).
template <typename T>
static inline T* map_shared_rank(T* p, int rank) {
  Cta& c = *g_cta;
  if (c.n_ctas == 1 || rank == c.live_cta) return p;
  return reinterpret_cast<T*>(c.shared_copy[rank].data() + (reinterpret_cast<char*>(p) - __start_emu_shared));
}

// ---- warp collectives: publish, wait for all 32 lanes, read, wait for all lanes to have read
template <typename F>
static inline uint64_t warp_collective(uint64_t v, F&& combine) {
  Cta& c = *g_cta;
  const int lane = self().tid.x % WARP;
  WarpState& w = c.warps[(size_t)self().cta * ((c.threads_per_cta + WARP - 1) / WARP) + self().tid.x / WARP];
  const unsigned gen = w.generation;
  w.slot[lane] = v;
  ++w.arrived;
  ++c.progress;
  while (w.arrived < WARP && w.generation == gen) yield();
  const uint64_t r = combine(w.slot, lane);
  ++c.progress;
  if (++w.departed == WARP) {
    w.arrived = 0;
    w.departed = 0;
    ++w.generation;
  } else {
    while (w.generation == gen) yield();
  }
  return r;
}

}  // namespace emu

#define threadIdx (emu::self().tid)
#define blockIdx (uint3{emu::g_blockIdx.x + (unsigned)emu::self().cta, emu::g_blockIdx.y, emu::g_blockIdx.z})
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline uint64_t emu_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float emu_float(uint64_t b) { uint32_t u = (uint32_t)b; float f; memcpy(&f, &u, 4); return f; }

static inline void __syncthreads() { emu::syncthreads_count(0); }
static inline int __syncthreads_count(int p) { return emu::syncthreads_count(p); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_collective(0, [](const uint64_t*, int) { return (uint64_t)0; }); }

static inline unsigned __ballot_sync(unsigned, int pred) {
  return (unsigned)emu::warp_collective(pred ? 1 : 0, [](const uint64_t* s, int) {
    uint64_t m = 0;
    for (int l = 0; l < 32; ++l) m |= (s[l] & 1) << l;
    return m;
  });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

static inline uint64_t emu_shfl(uint64_t v, int src) {
  return emu::warp_collective(v, [src](const uint64_t* s, int lane) { return (src >= 0 && src < 32) ? s[src] : s[lane]; });
}
static inline uint64_t emu_shfl_lane(uint64_t v, int mode, int arg) {   // per-lane source: 0 xor, 1 down, 2 up
  return emu::warp_collective(v, [mode, arg](const uint64_t* s, int lane) {
    int src = mode == 0 ? (lane ^ arg) : mode == 1 ? lane + arg : lane - arg;
    return (src >= 0 && src < 32) ? s[src] : s[lane];
  });
}
static inline float __shfl_sync(unsigned, float v, int src) { return emu_float(emu_shfl(emu_bits(v), src & 31)); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)(uint32_t)emu_shfl((uint32_t)v, src & 31); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return (unsigned)emu_shfl(v, src & 31); }
static inline unsigned long long __shfl_sync(unsigned, unsigned long long v, int src) { return emu_shfl(v, src & 31); }
static inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long v, int m) { return emu_shfl_lane(v, 0, m); }
static inline float __shfl_xor_sync(unsigned, float v, int m) { return emu_float(emu_shfl_lane(emu_bits(v), 0, m)); }
static inline int __shfl_xor_sync(unsigned, int v, int m) { return (int)(uint32_t)emu_shfl_lane((uint32_t)v, 0, m); }
static inline float __shfl_down_sync(unsigned, float v, int d) { return emu_float(emu_shfl_lane(emu_bits(v), 1, d)); }
static inline unsigned __shfl_down_sync(unsigned, unsigned v, int d) { return (unsigned)emu_shfl_lane(v, 1, d); }
static inline unsigned __shfl_up_sync(unsigned, unsigned v, int d) { return (unsigned)emu_shfl_lane(v, 2, d); }
static inline unsigned long long __shfl_up_sync(unsigned, unsigned long long v, int d) { return emu_shfl_lane(v, 2, d); }
static inline int __shfl_up_sync(unsigned, int v, int d) { return (int)(uint32_t)emu_shfl_lane((uint32_t)v, 2, d); }

static inline unsigned __match_any_sync(unsigned, int v) {
  return (unsigned)emu::warp_collective((uint64_t)(uint32_t)v, [](const uint64_t* s, int lane) {
    uint64_t m = 0;
    for (int l = 0; l < 32; ++l) m |= (uint64_t)(s[l] == s[lane]) << l;
    return m;
  });
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  return (unsigned)emu::warp_collective(v, [](const uint64_t* s, int) { uint64_t r = 0; for (int l = 0; l < 32; ++l) r |= s[l]; return r; });
}
static inline int __reduce_max_sync(unsigned, int v) {
  return (int)(int64_t)emu::warp_collective((uint64_t)(int64_t)v, [](const uint64_t* s, int) {
    int64_t r = (int64_t)s[0];
    for (int l = 1; l < 32; ++l) r = (int64_t)s[l] > r ? (int64_t)s[l] : r;
    return (uint64_t)r;
  });
}

static inline int __reduce_add_sync(unsigned, int v) {
  return (int)(int64_t)emu::warp_collective((uint64_t)(int64_t)v, [](const uint64_t* s, int) {
    int64_t r = 0;
    for (int l = 0; l < 32; ++l) r += (int64_t)s[l];
    return (uint64_t)r;
  });
}

static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
#define __expf(x) expf(x)   // glibc declares __expf itself
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }

// one OS thread: plain read-modify-write
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline int atomicExch(int* p, int v) { int o = *p; *p = v; return o; }
static inline uint32_t __float_as_uint(float f) { return (uint32_t)emu_bits(f); }
static inline float __uint_as_float(uint32_t u) { return emu_float(u); }
static inline int __float_as_int(float f) { return (int)(uint32_t)emu_bits(f); }
static inline float __int_as_float(int i) { return emu_float((uint32_t)i); }
static inline float4 atomicAdd(float4* p, float4 v) {
  float4 o = *p;
  p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w;
  return o;
}

static inline int lane_id() { return (int)(threadIdx.x & 31); }

// ---- the part of cooperative_groups the kernels use (thread-block clusters)
namespace cooperative_groups {
struct cluster_group {
  unsigned block_rank() const { return (unsigned)emu::self().cta; }
  void sync() const { emu::cluster_sync(); }
  template <typename T> T* map_shared_rank(T* p, int rank) const { return emu::map_shared_rank(p, rank); }
};
static inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups
#define __cluster_dims__(...)
