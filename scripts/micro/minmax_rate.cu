// Issue rate of the max instructions the matcher's arg-max epilogue can be built from (B200, one CTA per SM), measured
// with asm volatile so that nothing is hoisted or folded: FMNMX (2-input fp32), FMNMX3 (3-input fp32), VIMNMX (2-input
// s32), VIMNMX3 (3-input s32: two dependent max.s32, fused by ptxas), FADD as the FMA-pipe yardstick.  Every thread keeps
// 96 values in registers and issues 32 independent instructions per iteration; prints cycles per warp instruction per SM
// sub-partition with 1, 2 and 4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o minmax_rate minmax_rate.cu && ./minmax_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op3(float a, float b, float c) {
  float d;
  if (OP == 0) asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  if (OP == 1) asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  if (OP == 2) { int t, r; asm volatile("max.s32 %0, %1, %2;" : "=r"(t) : "r"(__float_as_int(a)), "r"(__float_as_int(b)));
                 asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(t), "r"(__float_as_int(c))); d = __int_as_float(r); }
  if (OP == 3) { int r; asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(__float_as_int(a)), "r"(__float_as_int(b))); d = __int_as_float(r); }
  if (OP == 4) asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

template <int OP>
__global__ void k(const float* in, float* out, int iters, long long* cyc) {
  float v[96];
#pragma unroll
  for (int i = 0; i < 96; ++i) v[i] = in[(threadIdx.x + i * 32) & 1023];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[3 * i] = op3<OP>(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 96; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int warps, const float* in, float* out, long long* cyc) {
  const int iters = 4000;
  k<OP><<<148, warps * 32>>>(in, out, 10, cyc);
  k<OP><<<148, warps * 32>>>(in, out, iters, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("%-34s %d warps / sub-partition: %.2f cycles per warp instruction per sub-partition\n", name, warps / 4,
         (double)c / iters / 32.0 / (warps / 4.0));
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (float)((i * 7919) % 1000) / 1000.f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int warps : {4, 8, 16}) {
    run<0>("FMNMX3  (max.f32 3-input)", warps, in, out, cyc);
    run<1>("FMNMX   (max.f32 2-input)", warps, in, out, cyc);
    run<2>("VIMNMX3 (2 dependent max.s32)", warps, in, out, cyc);
    run<3>("VIMNMX  (max.s32 2-input)", warps, in, out, cyc);
    run<4>("FADD", warps, in, out, cyc);
  }
  return 0;
}
