// Internal interface of the nearest-neighbour kernels.
#pragma once
#include "common.cuh"

// (score, index) packed so that 64-bit max == (highest score, then lowest index).
__host__ __device__ __forceinline__ unsigned long long nn_pack_bits(uint32_t ordered, int j) {
  return ((unsigned long long)ordered << 32) | (unsigned long long)(0xffffffffu - (uint32_t)j);
}
#if defined(__CUDACC__) || defined(ST3R_HOST_EMU)
__device__ __forceinline__ unsigned long long nn_pack(float s, int j) {
  return nn_pack_bits(float_to_ordered(s), j);
}
__device__ __forceinline__ int nn_unpack_idx(unsigned long long key) {
  return (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull));
}
__device__ __forceinline__ float nn_unpack_score(unsigned long long key) {
  return ordered_to_float((uint32_t)(key >> 32));
}
#endif

// Q row i is Qsrc[qidx ? qidx[i] : i]; only rows i < min(*count_ptr, Mmax) are
// processed (count_ptr may be NULL -> Mmax).  `packed` must be zeroed beforehand.
int nn_simt_launch(const float* Qsrc, const int32_t* qidx, const int32_t* count_ptr, int Mmax,
                   const float* DB, int N, int d, unsigned long long* packed, cudaStream_t stream);

// tcgen05 (TF32 tensor-core) candidate search + exact fp32 re-score.  Same contract.
// `db_norm_bound` = device pointer to max_j ||DB_j||_2^2 (float), see nn_db_norm_launch.
int nn_tc_launch(const float* Qsrc, const int32_t* qidx, const int32_t* count_ptr, int Mmax,
                 const float* DB, int N, int d, const float* db_norm_bound,
                 unsigned long long* packed, cudaStream_t stream, const float* DB_hi = nullptr,
                 const float* DB_lo = nullptr);
int nn_db_norm_launch(const float* DB, int N, int d, float* out_bound, cudaStream_t stream);
bool nn_tc_supported(int d);

// Batched form: up to NN_MAX_BATCH independent NN problems in ONE launch (the four seeded searches of an image
// pair advance in lock-step, so batching them divides the length of the launch chain by four).
constexpr int NN_MAX_BATCH = 4;
struct NnBatchItem {
  const float* Qsrc; const int32_t* qidx; const int32_t* count_ptr; int Mmax;
  const float* DB; int N; const float* db_norm_bound;
  unsigned long long* packed;
  // Optional tf32 head / tail of DB (nn_tc_split_launch): with both present and st3r_nn_tc_set_split(1) the
  // split-precision variant of the tcgen05 kernel runs (3 x TF32 products, ~40x narrower candidate band).
  const float* DB_hi = nullptr; const float* DB_lo = nullptr;
};
int nn_tc_split_launch(const float* DB, int N, int d, float* hi, float* lo, cudaStream_t stream);
bool nn_tc_split_enabled();
int nn_tc_launch_batch(const NnBatchItem* items, int n, int d, cudaStream_t stream);
