// Issue rate of the packed fp32 instructions of sm_100a (FFMA2 / FMUL2 / FADD2: fma.rn.f32x2 etc., two fp32 lanes per
// 64-bit register pair) against the scalar FFMA, measured with asm volatile on independent accumulators: prints cycles
// per warp instruction per SM sub-partition and the fp32 FMA lanes per clock per SM this amounts to.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(const float* in, float* out, int iters, long long* cyc) {
  float a[32], b[32];
  unsigned long long p[16], q[16];
#pragma unroll
  for (int i = 0; i < 32; ++i) { a[i] = in[(threadIdx.x + i * 32) & 1023]; b[i] = in[(threadIdx.x + i * 32 + 7) & 1023]; }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    q[i] = ((unsigned long long)__float_as_uint(b[2 * i + 1]) << 32) | __float_as_uint(b[2 * i]);
  }
  const float w = in[5];
  const unsigned long long w2 = ((unsigned long long)__float_as_uint(w) << 32) | __float_as_uint(w);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(w));
    } else if (OP == 1) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(q[i]), "l"(w2));
    } else if (OP == 2) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(w2));
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int lanes_per_instr, int warps, const float* in, float* out, long long* cyc) {
  const int iters = 4000;
  k<OP><<<148, warps * 32>>>(in, out, 10, cyc);
  k<OP><<<148, warps * 32>>>(in, out, iters, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double cpi = (double)c / iters / 32.0 / (warps / 4.0);
  printf("{\"op\": \"%s\", \"warps_per_subpartition\": %d, \"cycles_per_warp_instruction\": %.3f, \"fp32_lanes_per_clk_per_sm\": %.1f}\n",
         name, warps / 4, cpi, 4.0 * 32.0 * lanes_per_instr / cpi);
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 1024 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = (float)((i * 7919) % 1000) / 4000.f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int warps : {4, 8, 16}) {
    run<0>("FFMA", 1, warps, in, out, cyc);
    run<1>("FFMA2", 2, warps, in, out, cyc);
    run<2>("FMUL2", 2, warps, in, out, cyc);
    run<3>("FADD2", 2, warps, in, out, cyc);
  }
  return 0;
}
