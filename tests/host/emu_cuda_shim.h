// TEST-ONLY: the sliver of the CUDA runtime API the library's HOST code uses, mapped onto the SIMT emulator
// (simt_emu.h), so that whole entry points of the C ABI - their launch sequences, workspace carving, device-side
// counters - run unmodified on the CPU.  "Device" memory is host memory, streams are ignored (everything is
// synchronous), kernel launches (rewritten from <<< >>> by tests/host/build_emu_lib.py) run CTA after CTA.
#pragma once
#define __CUDA_RUNTIME_H__      // include/starst3r_b200.h: cudaStream_t comes from here
#include <map>
#include "simt_emu.h"

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef struct CUstream_st* cudaStream_t;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };

static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// dynamic shared memory of the kernel being launched (one CTA at a time)
alignas(1024) static char emu_dyn_smem[256 * 1024];
#define ST3R_DYN_SMEM_F32(name) float* name = reinterpret_cast<float*>(emu_dyn_smem)
#define ST3R_DYN_SMEM(name) float* name = reinterpret_cast<float*>(emu_dyn_smem)
#define ST3R_DYN_SMEM_I32(name) int32_t* name = reinterpret_cast<int32_t*>(emu_dyn_smem)
#define ST3R_DYN_SMEM_U64(name) uint64_t* name = reinterpret_cast<uint64_t*>(emu_dyn_smem)
#define __constant__ static
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void*)&(sym), (src), (n)), cudaSuccess)
#define cudaMemcpyFromSymbol(dst, sym, n) (memcpy((dst), (const void*)&(sym), (n)), cudaSuccess)
#define __grid_constant__
template <typename T> static inline T __ldcv(const T* p) { return *p; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }

static bool g_emu_launch_failed = false;
template <typename F>
static inline void emu_launch(dim3 grid, dim3 block, size_t smem, F&& body) {
  if (smem > sizeof(emu_dyn_smem) || block.y != 1 || block.z != 1) { g_emu_launch_failed = true; return; }
  emu::g_gridDim = grid;
  emu::g_blockDim = block;
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        emu::g_blockIdx = uint3{x, y, z};
        if (!emu::run_cta((int)block.x, body)) { g_emu_launch_failed = true; return; }
      }
}
#define EMU_LAUNCH(kernel, grid, block, smem, ...) emu_launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
// kernels declared with __cluster_dims__(n, 1, 1): the grid is walked cluster by cluster
template <typename F>
static inline void emu_launch_cluster(int cluster, dim3 grid, dim3 block, F&& body) {
  emu::g_gridDim = grid;
  emu::g_blockDim = block;
  for (unsigned x = 0; x < grid.x; x += cluster) {
    emu::g_blockIdx = uint3{x, 0, 0};
    if (!emu::run_cluster(cluster, (int)block.x, body, 64 * 1024)) { g_emu_launch_failed = true; return; }
  }
}
