// Stable LSD radix sort of (uint64 key, uint32 value) pairs; see radix_sort.cu.
#pragma once
#include "common.cuh"

size_t radix_sort_ws_bytes(int n_cap);

// Sorts keys[0..n) (n = n_ptr ? min(*n_ptr, n_cap) : n_cap) ascending on bits
// [begin_bit, end_bit), stably; vals (optional, may be NULL together with
// vals_alt) are permuted alongside.  Result always ends in keys / vals.
int radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                     const int* n_ptr, int n_cap, int begin_bit, int end_bit, void* ws,
                     size_t ws_bytes, cudaStream_t stream);
