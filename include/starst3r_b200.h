/* starst3r_b200 — C ABI of the B200 (sm_100a) hot-path library `libstarst3r_b200.so`.
 *
 * This is the drop-in boundary for the two data-parallel hot loops of
 * phuang1024/Starst3r (see DESIGN.md §2 and INTEGRATION.md):
 *   MATCH  : mast3r/mast3r/fast_nn.py, mast3r/mast3r/cloud_opt/sparse_ga.py:595-630
 *   ALIGN  : starster/reconstruct.py:116-457, sparse_ga.py:464-501,817-886,977-981
 *   RASTER : starster/gs.py:76-87 (gsplat.rasterization), gs.py:126-136 (loss),
 *            gs.py:37,159-161 (Adam)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named `h_*`; tensors are dense,
 *     row-major, fp32 / int32 / int64 as typed;
 *   - every entry point enqueues work on `stream` and returns without
 *     synchronising; element counts that are only known on the device are
 *     written to device `int32_t*` outputs;
 *   - no hidden allocation: scratch comes from the caller (`ws`, `ws_bytes`;
 *     query the size with the matching `*_ws_bytes`);
 *   - return 0 on success, a negative ST3R_ERR_* otherwise; the message is
 *     available from st3r_last_error() (thread-local); nothing throws.
 */
#ifndef STARST3R_B200_H_
#define STARST3R_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define ST3R_ABI_VERSION 1

#if defined(__GNUC__)
#define ST3R_API __attribute__((visibility("default")))
#else
#define ST3R_API
#endif

#define ST3R_NN_AUTO 0    /* tcgen05 when supported (d == 24), else SIMT */
#define ST3R_NN_SIMT 1    /* exact fp32 FMA-chain on CUDA cores */
#define ST3R_NN_TCGEN05 2 /* TF32 tcgen05 candidate search + exact fp32 re-score */

ST3R_API const char* st3r_last_error(void);
ST3R_API int st3r_abi_version(void);
ST3R_API int st3r_device_sm_count(void);

/* ------------------------------------------------------------------ MATCH */

/* Row arg-max of Q·DBᵀ (dot-product nearest neighbour), ties -> lowest index.
 * Replaces bruteforce_reciprocal_nns(A, B, dist='dot') -> nn_A
 * (mast3r/mast3r/fast_nn.py:16-70; the column arg-max it also computes is
 * discarded by its only caller, fast_nn.py:82-84).
 * Q [M,d], DB [N,d] fp32; idx [M] int32; best [M] fp32 (may be NULL). */
ST3R_API size_t st3r_nn_argmax_ws_bytes(int M, int N, int d);
ST3R_API int st3r_nn_argmax(const float* Q, int M, const float* DB, int N, int d, int32_t* idx, float* best,
                   void* ws, size_t ws_bytes, int impl, cudaStream_t stream);

/* Number of grid seeds np.mgrid[S//2:H:S, S//2:W:S] yields (fast_nn.py:118-121). */
ST3R_API int st3r_recip_seed_count(int H, int W, int subsample);

/* Seeded iterative reciprocal NN search, fast_reciprocal_NNs(pts1, pts2,
 * subsample_or_initxy1, ret_xy=False, pixel_tol=0, ret_basin=False)
 * (fast_nn.py:109-188).  P1 [H1*W1,d], P2 [H2*W2,d].  If `seeds` is NULL the
 * seed grid of `subsample` is used, else `seeds[nseeds]` (sorted unique flat
 * indices into map 1).  Outputs: unique (idx1, idx2) sorted by (idx1, idx2),
 * capacity = number of seeds; *n_out = how many. */
ST3R_API size_t st3r_recip_nn_ws_bytes(int nseed_max, int key_cap, int max_iter);
ST3R_API int st3r_recip_nn(const float* P1, int H1, int W1, const float* P2, int H2, int W2, int d, int subsample,
                  const int32_t* seeds, int nseeds, int max_iter, int32_t* out_idx1, int32_t* out_idx2,
                  int32_t* n_out, void* ws, size_t ws_bytes, int impl, cudaStream_t stream);

/* merge_corres(idx1, idx2, ret_xy=False, ret_index=True) (fast_nn.py:87-106):
 * unique pairs sorted by (idx1, idx2) and the position of each pair's first
 * occurrence.  hw1/hw2 = exclusive upper bounds of the index values.
 * out_index may be NULL. */
ST3R_API size_t st3r_merge_corres_ws_bytes(int n);
ST3R_API int st3r_merge_corres(const int32_t* idx1, const int32_t* idx2, int n, int hw1, int hw2, int32_t* out_idx1,
                      int32_t* out_idx2, int32_t* out_index, int32_t* n_out, void* ws, size_t ws_bytes,
                      cudaStream_t stream);

/* extract_correspondences(feats, qonfs, subsample) (sparse_ga.py:595-630) for one
 * image pair: both descriptor sets, both directions, merged; conf = sqrt(q1*q2).
 * feat11/feat12 [H1,W1,d], feat21/feat22 [H2,W2,d]; qonf* matching [H,W].
 * Outputs (capacity st3r_extract_corres_cap rows): xy1/xy2 [cap,2] int64 (x,y),
 * conf [cap] fp32; *n_out rows are valid. */
ST3R_API int st3r_extract_corres_cap(int H1, int W1, int H2, int W2, int subsample);
ST3R_API size_t st3r_extract_corres_ws_bytes(int H1, int W1, int H2, int W2, int subsample, int max_iter);
ST3R_API int st3r_extract_corres(const float* feat11, const float* feat21, const float* feat22, const float* feat12,
                        const float* qonf11, const float* qonf21, const float* qonf22, const float* qonf12,
                        int H1, int W1, int H2, int W2, int d, int subsample, int max_iter,
                        int64_t* out_xy1, int64_t* out_xy2, float* out_conf, int32_t* n_out,
                        void* ws, size_t ws_bytes, int impl, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STARST3R_B200_H_ */
