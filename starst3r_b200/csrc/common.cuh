// Shared helpers for the starst3r_b200 CUDA kernels (sm_100a only).
#pragma once
#ifdef ST3R_HOST_EMU            // test builds that run the library on a CPU SIMT emulator (tests/host/)
#include "emu_cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>

#define ST3R_OK 0
#define ST3R_ERR_BAD_ARG (-1)
#define ST3R_ERR_CUDA (-2)
#define ST3R_ERR_WORKSPACE (-3)
#define ST3R_ERR_UNSUPPORTED (-4)

// Records a message retrievable through st3r_last_error() (thread-local).
void st3r_set_error(const char* fmt, ...);

#define ST3R_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      st3r_set_error(__VA_ARGS__);           \
      return ST3R_ERR_BAD_ARG;               \
    }                                        \
  } while (0)

#define ST3R_CHECK_CUDA(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      st3r_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),         \
                     __FILE__, __LINE__);                                            \
      return ST3R_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

// Counts every kernel this library launches (bench.py reports it as gpu_launches).
extern unsigned long long g_st3r_launches;
#define ST3R_CHECK_LAUNCH()                  \
  do {                                       \
    ++g_st3r_launches;                       \
    ST3R_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

static inline size_t st3r_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace (no hidden allocation).
struct WsAlloc {
  char* base;
  size_t off;
  size_t cap;
  __host__ WsAlloc(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes) {}
  template <typename T>
  __host__ T* take(size_t n) {
    off = st3r_align_up(off, 256);
    T* r = (T*)(base + off);
    off += n * sizeof(T);
    return r;
  }
  __host__ bool ok() const { return off <= cap; }
};

#ifndef ST3R_HOST_EMU            // (the emulator header brings its own)
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
#endif

// Monotone map float -> uint32 (larger float -> larger uint).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

int st3r_num_sms();

// "Has this call site already configured its kernels on the CURRENT device?"  Function attributes
// (cudaFuncSetAttribute) belong to a device's context, so a process that works on two GPUs has to set them on both.
struct PerDeviceOnce {
  unsigned long long seen = 0;
  static unsigned long long bit() {
    int d = 0;
    cudaGetDevice(&d);
    return 1ull << (d & 63);
  }
  bool done() const { return (seen & bit()) != 0; }
  void mark() { seen |= bit(); }
};
