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
/* Number of CUDA kernels this library has launched in this process (all entry points). */
ST3R_API uint64_t st3r_launch_count(void);
/* Adds n to that counter: a replayed CUDA graph launches the kernels that were counted once, at capture. */
ST3R_API void st3r_launch_count_add(uint64_t n);

/* Diagnostic: cycle counters of the tcgen05 matcher's epilogue warp in CTA 0, h_out4 (HOST) = {tiles, cycles
 * waiting for an accumulator, cycles in the arg-max epilogue, cycles of the whole tile loop}; reset != 0 clears. */
ST3R_API int st3r_debug_nn_tc_cycles(unsigned long long* h_out4, int reset);

/* The tcgen05 matcher has two variants of its rare path (exact re-scoring of near-tie columns): per-thread (best when
 * 1-3 columns per query row tie within the TF32 error band: random-like descriptors) and warp-cooperative (best for
 * smooth descriptor fields, where ~100 do).  Results are identical.  st3r_nn_tc_stats copies {query rows scanned, exact
 * list resolutions} accumulated by all launches so far to h_out2 (HOST; synchronises), optionally resetting them;
 * st3r_nn_tc_set_cooperative selects the variant for subsequent launches of this process. */
ST3R_API int st3r_nn_tc_stats(unsigned long long* h_out2, int reset);
ST3R_API int st3r_nn_tc_set_cooperative(int on);
/* Split-precision variant of the tcgen05 matcher (off by default): every operand is split into its tf32 head and the
 * tf32 head of the remainder and the kernel accumulates hi.hi + hi.lo + lo.hi (3 x the tensor work), which narrows
 * the band of columns that need an exact fp32 re-score ~40x - for smooth descriptor fields, where ~100 columns per
 * row sit inside the plain TF32 band.  Results are identical (the exact re-score decides either way).  While it is
 * on, st3r_nn_argmax_ws_bytes and st3r_extract_corres_ws_bytes include two [rows, 24] float arrays per descriptor
 * map; st3r_recip_nn always runs the plain variant. */
ST3R_API int st3r_nn_tc_set_split(int on);
/* Diagnostic: candidate-band width (relative to |q| max|db|) of the plain / split-precision tcgen05 kernel for the
 * launches that follow; 0 restores the built-in value.  scripts/nn_split_margin.py narrows the band until results stop
 * being exact, which measures the tensor core's real error and hence the safety margin of the built-in values. */
ST3R_API int st3r_debug_nn_tc_set_delta_coef(float plain, float split);

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

/* Implementation variant of the sort + unique step of st3r_merge_corres / st3r_recip_nn / st3r_extract_corres for
 * subsequent calls of this process.  2 (default): lists of up to 16384 keys (every pair up to 512 x 512 at subsample 8)
 * are sorted in registers and compacted by one CTA in one launch, through 32-bit surrogate words (idx1 | position)
 * and repair passes where idx1 and the position fit into 32 bits; 1: the same CTA on the 64-bit words; 0: two LSD radix
 * sorts + a compaction kernel for every size (the generic chain, also the path of longer lists).  Identical outputs. */
ST3R_API int st3r_recip_set_variant(int variant);

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

/* ------------------------------------------------------------------ RASTER + ADAM
 * Replaces gsplat.rasterization(means, quats, scales, opacities, colors=shN, viewmats, Ks, width,
 * height, sh_degree=1) as called at starster/gs.py:76-87 (gsplat 1.4 semantics, SURVEY.md
 * Appendix A; packed=True, tile_size=16, eps2d=0.3, near=0.01, far=1e10, radius_clip=0, classic
 * mode, no background), its backward (loss.backward(), gs.py:153), the loss of gs.py:126-136 and
 * the six Adam optimisers of gs.py:37,159-161.
 *
 * Intermediate layout (dense, entry e = camera * N + gaussian):
 *   radii int32 [C*N] (0 = culled); geomA float4 [C*N] = (mean2d.x, mean2d.y, opacity, depth);
 *   geomB float4 [C*N] = (conic a, b, c, 0); rgb float4 [C*N] = (r, g, b, 0);
 *   tiles int32 [C*N] = number of 16x16 tiles touched.
 * `cams` is [C, st3r_gs_cam_floats()] fp32: R (9, row-major world->camera), t (3), fx, fy, cx, cy,
 * camera centre (3).  Quaternions are wxyz.  shN is [N, sh_coeffs, 3]. */
ST3R_API int st3r_gs_cam_floats(void);
ST3R_API int st3r_gs_project(const float* means, const float* quats, const float* scales, const float* opacities,
                    const float* shN, int sh_coeffs, const float* cams, int N, int C, int width, int height,
                    int tile_size, float eps2d, float near_plane, float far_plane, float radius_clip,
                    int32_t* radii, float* geomA, float* geomB, float* rgb, int32_t* tiles, cudaStream_t stream);

/* Exclusive prefix sum (torch.cumsum of tiles_per_gauss); *total_out = sum (device). */
ST3R_API size_t st3r_scan_ws_bytes(size_t n);
ST3R_API int st3r_exclusive_scan_i32(const int32_t* in, int32_t* out, size_t n, int32_t* total_out, void* ws, size_t ws_bytes,
                            cudaStream_t stream);

/* isect_tiles: keys[i] = camera << (32 + tile_bits) | tile << 32 | bits(depth), vals[i] = entry e. */
ST3R_API int st3r_gs_isect(const int32_t* radii, const float* geomA, const int32_t* cum_tiles, int N, int C, int width,
                  int height, int tile_size, uint64_t* keys, uint32_t* vals, int n_cap, cudaStream_t stream);
ST3R_API int st3r_gs_sort_bits(int C, int width, int height, int tile_size);

/* Stable LSD radix sort of (key, value) pairs on bits [begin_bit, end_bit) (CUB SortPairs stand-in).
 * n = n_ptr ? min(*n_ptr, n_cap) : n_cap.  vals / vals_alt may both be NULL (keys only). */
ST3R_API size_t st3r_radix_sort_ws_bytes(int n_cap);
ST3R_API int st3r_radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_alt, uint32_t* vals_alt,
                          const int32_t* n_ptr, int n_cap, int begin_bit, int end_bit, void* ws, size_t ws_bytes,
                          cudaStream_t stream);

/* isect_offset_encode: offsets [C, tile_h, tile_w] int32. */
ST3R_API int st3r_gs_offsets(const uint64_t* keys, const int32_t* n_isect, int n_cap, int C, int width, int height,
                    int tile_size, int32_t* offsets, cudaStream_t stream);

/* isect_tiles + SortPairs + isect_offset_encode in one call (same outputs, bit for bit, as the three entry points
 * above): counting sort on (camera, tile), shared-memory sort on (depth, entry) inside each tile.  offsets
 * [C, tile_h, tile_w] int32; *n_isect_out (device) = true number of intersections; keys / vals capacity n_cap (pairs
 * beyond it are dropped and offsets is clamped to it; n_cap = 0 only counts). */
ST3R_API size_t st3r_gs_bin_ws_bytes(int C, int width, int height, int tile_size, int n_cap);
/* Implementation variant of the per-tile sort for subsequent calls of this process.  2 (default): segments of up to
 * 2048 pairs are sorted in registers through 32-bit surrogate keys (leading depth bits | position) and the full
 * (depth, entry) order is restored by odd-even passes in shared memory; 1: the same segments as 64-bit words (warp
 * shuffles, shared memory only for the longest spans); 0: the shared-memory / in-place comparator network for every
 * segment.  Identical outputs. */
ST3R_API int st3r_gs_bin_set_variant(int variant);
ST3R_API int st3r_gs_bin_tiles(const int32_t* radii, const float* geomA, int N, int C, int width, int height, int tile_size,
                      int32_t* offsets, int32_t* n_isect_out, uint64_t* keys, uint32_t* vals, int n_cap, void* ws,
                      size_t ws_bytes, cudaStream_t stream);

/* rasterize_to_pixels forward: render [C,H,W,3], alphas [C,H,W], last_ids [C,H,W] int32;
 * n_blend (optional, uint64, accumulated) counts blended (pixel, Gaussian) pairs. */
ST3R_API int st3r_gs_raster_fwd(const int32_t* offsets, const int32_t* n_isect, const uint32_t* flatten_ids,
                       const float* geomA, const float* geomB, const float* rgb, int C, int width, int height,
                       int tile_size, float* render, float* alphas, int32_t* last_ids, uint64_t* n_blend,
                       cudaStream_t stream);
/* Implementation of st3r_gs_raster_fwd / _bwd for subsequent calls of this process: 0 (default) = fragment-pool
 * kernels (alpha evaluated once per pixel of each splat's bounding box inside the tile into a shared-memory pool,
 * per-pixel recurrence over the pixel's own contributing Gaussians, gradients summed per Gaussian in registers; batches
 * of tile-sized splats are walked pixel-parallel); 1 = visit-list kernels (every warp tests every Gaussian that can touch
 * its two pixel rows; the first implementation, kept as an independent cross-check).  Same terms per (pixel, Gaussian):
 * the forward outputs are bit-identical, the backward differs by fp32 summation order. */
ST3R_API int st3r_gs_set_raster_variant(int variant);
/* rasterize_to_pixels backward: accumulates into v_geomA = (v_x, v_y, v_opacity, 0), v_geomB = v_conic,
 * v_rgb (all float4 [C*N], zeroed by the caller).  v_alphas may be NULL. */
ST3R_API int st3r_gs_raster_bwd(const int32_t* offsets, const int32_t* n_isect, const uint32_t* flatten_ids,
                       const float* geomA, const float* geomB, const float* rgb, int C, int width, int height,
                       int tile_size, const float* alphas, const int32_t* last_ids, const float* v_render,
                       const float* v_alphas, float* v_geomA, float* v_geomB, float* v_rgb, cudaStream_t stream);
/* projection + SH backward, summed over cameras, plus d/d(opacity, scale) of the regularisers
 * (reg_opac = C * fac / N, reg_scale = C * fac / (3N)).  v_sh is [N, 4, 3] (the SH coefficients degree 1
 * touches).  reg_sums (optional, float[2], accumulated): sum sigmoid(opacity), sum exp(scale). */
ST3R_API int st3r_gs_project_bwd(const float* means, const float* quats, const float* scales, const float* opacities,
                        const float* shN, int sh_coeffs, const float* cams, int N, int C, int width,
                        int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                        const int32_t* radii, const float* v_geomA, const float* v_geomB,
                        const float* v_rgb, float reg_opac, float reg_scale, float* v_means,
                        float* v_quats, float* v_scales, float* v_opacities, float* v_sh, float* reg_sums,
                        cudaStream_t stream);

/* (1-f) * L1 + f * (1 - SSIM) per view (gs.py:126-131; SSIM = torchmetrics, data_range 1).
 * render/truth [C,H,W,3]; dmaps: 9*C*H*W floats of scratch (layout private to the pair of calls); sums [C,2] (zeroed by caller) receives
 * (sum of interior SSIM, sum |truth - render|).  bwd writes v_render = dLoss/d(render). */
ST3R_API int st3r_gs_loss_fwd(const float* render, const float* truth, int C, int height, int width, float ssim_fac,
                     float* dmaps, float* sums, cudaStream_t stream);
ST3R_API int st3r_gs_loss_bwd(const float* render, const float* truth, const float* dmaps, int C, int height, int width,
                     float ssim_fac, float* v_render, cudaStream_t stream);

/* Scalar loss of one training iteration from the accumulators of st3r_gs_loss_fwd / st3r_gs_project_bwd:
 * sum over views of (1-f) * L1 + f * (1 - SSIM) (gs.py:126-131,149-152) + reg_opac * reg_sums[0] + reg_scale *
 * reg_sums[1] (gs.py:132-136; reg_sums may be NULL).  loss_out is one device float. */
ST3R_API int st3r_gs_loss_finalize(const float* sums, const float* reg_sums, int C, int height, int width, float ssim_fac,
                          float reg_opac, float reg_scale, float* loss_out, cudaStream_t stream);

/* torch.optim.Adam step over up to 8 tensors in one launch.  Arrays are HOST arrays of device pointers /
 * sizes; tensor i is rows[i] x cols[i] with leading dimensions ld_param[i] (param and both moments) and
 * ld_grad[i].  step is 1-based; the hyper-parameters are doubles (torch keeps them as Python floats and rounds
 * 1 - beta to fp32 once). */
ST3R_API int st3r_adam_step(int n_seg, float* const* h_params, const float* const* h_grads, float* const* h_exp_avg,
                   float* const* h_exp_avg_sq, const int* h_rows, const int* h_cols, const int* h_ld_param,
                   const int* h_ld_grad, double lr, double beta1, double beta2, double eps, int step,
                   cudaStream_t stream);

/* st3r_adam_step with the step number kept on the device: *steps_done (device int32) is the number of completed steps;
 * the update uses step = *steps_done + 1 (bias corrections evaluated on the device in double precision, like the host
 * form) and then increments the counter.  The call carries no per-step scalar, so a captured CUDA graph of the whole
 * training iteration replays unchanged (gs.TrainPlan). */
ST3R_API int st3r_adam_step_dev(int n_seg, float* const* h_params, const float* const* h_grads, float* const* h_exp_avg,
                       float* const* h_exp_avg_sq, const int* h_rows, const int* h_cols, const int* h_ld_param,
                       const int* h_ld_grad, double lr, double beta1, double beta2, double eps, int32_t* steps_done,
                       cudaStream_t stream);

/* Multi-GPU form of st3r_adam_step: gradient all-reduce fused with the Adam update over peer memory.  Every rank
 * keeps its per-Gaussian gradients in a symmetric buffer mapped by all ranks (NVLink P2P); h_peer_grad_bases[r] is
 * rank r's buffer as seen from THIS device, h_grad_offsets[i] the float offset of tensor i inside it (same on all
 * ranks; row stride h_ld_grad[i]).  Each element is summed over the ranks in rank order (identical on every replica)
 * and applied at once.  The caller orders the kernel after every rank's gradient writes (cross-device barrier). */
ST3R_API int st3r_adam_step_peers(int n_seg, float* const* h_params, const long long* h_grad_offsets, float* const* h_exp_avg,
                         float* const* h_exp_avg_sq, const int* h_rows, const int* h_cols, const int* h_ld_param,
                         const int* h_ld_grad, int world, const float* const* h_peer_grad_bases, double lr,
                         double beta1, double beta2, double eps, int step, cudaStream_t stream);

/* The same exchange as reduce-scatter + all-gather for larger nodes: this rank sums its 1/world slice of every rank's
 * gradient buffer (rank order: bit-identical to st3r_adam_step_peers) and stores it into every rank's `reduced` buffer
 * (both symmetric, n_floats a multiple of 4, 16-byte aligned); after a cross-device barrier st3r_adam_step runs on the
 * local reduced buffer.  Remote traffic per GPU: 2 (G-1)/G L instead of (G-1) L. */
ST3R_API int st3r_grad_reduce_scatter(int world, int rank, const float* const* h_peer_grad_bases,
                             float* const* h_peer_reduced_bases, int64_t n_floats, cudaStream_t stream);

/* NVLS form of st3r_grad_reduce_scatter: `mc_grads` / `mc_reduced` are the MULTICAST addresses of the symmetric gradient
 * and `reduced` buffers (NVSwitch multicast objects, e.g. torch symmetric memory's multicast_ptr).  This rank sums its
 * 1/world slice inside the switch (multimem.ld_reduce) and stores it to every rank (multimem.st): L / G floats in and
 * L / G out per GPU.  Replicas receive identical values; the in-switch summation order is unspecified. */
ST3R_API int st3r_grad_reduce_multimem(int world, int rank, const float* mc_grads, float* mc_reduced, int64_t n_floats,
                              cudaStream_t stream);

/* ------------------------------------------------------------------ MCMC strategy
 * gsplat.MCMCStrategy as driven by starster/gs.py:43-45 (construction), :146-147 (step_pre_backward, a no-op for
 * MCMC) and :163-164 (step_post_backward(..., lr=1e-3)); gsplat 1.4 strategy/mcmc.py + strategy/ops.py semantics
 * (SURVEY.md Appendix A.8).  Random draws (torch.randn_like, torch.multinomial) stay with the caller so the RNG
 * stream is the reference's; these entry points are deterministic given the draws. */

/* inject_noise_to_position: means += Sigma * (noise * op_sigmoid(1 - sigmoid(opacity)) * scaler), Sigma from
 * normalised wxyz quats and exp(scales); op_sigmoid(x) = 1 / (1 + exp(-100 (x - 0.995))).  noise [N,3] ~ N(0,1). */
ST3R_API int st3r_mcmc_inject_noise(float* means, const float* quats, const float* scales, const float* opacities,
                           const float* noise, int N, float scaler, cudaStream_t stream);
/* compute_relocation(opacities, scales, ratios, binoms): ACTIVATED opacities [N] / scales [N,3], ratios [N] int32
 * (clamped to [1, n_max]), binoms [n_max, n_max] (binoms[n,k] = C(n,k)) -> new_opacities [N], new_scales [N,3]. */
ST3R_API int st3r_mcmc_compute_relocation(const float* opacities, const float* scales, const int32_t* ratios,
                                 const float* binoms, int n_max, int N, float* new_opacities, float* new_scales,
                                 cudaStream_t stream);
/* Dead / alive split of MCMCStrategy._relocate_gs: probs[i] = sigmoid(opacities_raw[i]); dead <=> probs <= min_opacity.
 * dead_idx / alive_idx (capacity N, ascending = torch.nonzero order), alive_probs[k] = probs[alive_idx[k]],
 * *n_dead (device). */
ST3R_API size_t st3r_mcmc_partition_ws_bytes(int N);
ST3R_API int st3r_mcmc_partition(const float* opacities_raw, int N, float min_opacity, int32_t* dead_idx, int32_t* alive_idx,
                        float* probs, float* alive_probs, int32_t* n_dead, void* ws, size_t ws_bytes,
                        cudaStream_t stream);
/* relocate() / sample_add() of gsplat strategy/ops.py on RAW parameters (opacities = logit, scales = log).
 * sampled [n] int64 = torch.multinomial draws; source s_i = alive_idx ? alive_idx[sampled[i]] : sampled[i];
 * destination d_i = dst ? dst[i] : dst_base + i (dead rows for relocate; appended rows N.. for sample_add, the
 * tensors must already have room).  ratio = bincount(sources)[s_i] + 1; the sources AND destinations receive
 * logit(clamp(new_opacity, min_opacity, 1 - eps)) / log(new_scale); every other parameter tensor (h_row_ptrs[k],
 * h_row_cols[k] floats per row: means, quats, sh0, shN) copies row s_i -> d_i; every moment tensor
 * (h_moment_ptrs: exp_avg / exp_avg_sq of all parameters; pass 0 for sample_add) gets row s_i zeroed.
 * counts [N] int32 is scratch.  h_* are HOST arrays (at most 16 entries each). */
ST3R_API int st3r_mcmc_relocate(float* opacities_raw, float* scales_raw, int n_rows, float* const* h_row_ptrs,
                       const int* h_row_cols, int n_moments, float* const* h_moment_ptrs, const int* h_moment_cols,
                       const int64_t* sampled, const int32_t* alive_idx, const int32_t* dst, int dst_base, int n, int N,
                       const float* binoms, int n_max, float min_opacity, int32_t* counts, cudaStream_t stream);

/* ------------------------------------------------------------------ ALIGN
 * Fused sparse-global-alignment optimiser: optimize_loop of starster/reconstruct.py:371-406 with
 * make_K_cam_depth (:209-261), loss_3d / loss_2d / loss_dust3r (:311-369), make_pts3d / reproj2d
 * (mast3r/cloud_opt/sparse_ga.py:469-501,977-981), gamma_loss and the Adam(lr=1, betas=(.9,.9)) step with
 * quaternion re-normalisation (:373-395).  All arrays are device pointers; the problem description is the
 * flattened form of the reference's condense_data output (sparse_ga.py:729-814). */
typedef struct {
  float W, H;            /* image size in pixels                                   (imsizes)          */
  float base_focal;      /* Weiszfeld focal of the canonical point map             (base_focals)      */
  float median;          /* median of the core depth map before normalisation      (median_depths)    */
  float min_focal, max_focal; /* 0.25 / 10 x image diagonal                        (reconstruct.py:203-205) */
  int32_t core_off, n_core;   /* slice of `core` owned by this image                                  */
} St3rAlignImgConst;

typedef struct {
  int32_t n_img;
  const St3rAlignImgConst* img_const;   /* [n_img] */
  const float* core;                    /* concatenated core depth maps, each divided by its median */
  int32_t n_core_total;
  int32_t root;                         /* MST root and BFS edge list (parent, child), sparse_ga.py:991-1009 */
  const int32_t* edges;                 /* [2 * (n_img - 1)] */
  /* anchors (sparse_ga.py:770-779): pixel (u, v), core-depth index, depth-ratio offset */
  int32_t n_anchor;
  const int32_t* anc_img; const float* anc_uv; const int32_t* anc_k; const float* anc_off;
  /* loss_3d entries: pairs of anchors + confidence; norm3 = sum of confidences */
  int32_t n3; const int32_t* e3_a1; const int32_t* e3_a2; const float* e3_conf; float norm3;
  /* loss_2d entries: pixel in image img1, the matching anchor in the other image; norm2 = sum conf */
  int32_t n2; const int32_t* e2_img1; const float* e2_pix; const int32_t* e2_a2; const float* e2_conf; float norm2;
  /* loss_dust3r entries: anchor of img1, regression target (in img2's camera frame), confidence */
  int32_t nd; const int32_t* ed_a1; const int32_t* ed_img2; const float* ed_tgt; const float* ed_conf; float normd;
} St3rAlignProblem;

ST3R_API size_t st3r_align_ws_bytes(int n_img);
/* Workspace size that additionally lets st3r_align_optimize run all `niter` iterations of a call as ONE cooperative
 * launch (variant bit 2 of st3r_align_set_variant): packed per-entry records of the loss terms, per-CTA gradient rows,
 * the schedule.  With the smaller st3r_align_ws_bytes workspace the loop runs as three launches per iteration. */
ST3R_API size_t st3r_align_ws_bytes_for(const St3rAlignProblem* prob, int niter);
ST3R_API int st3r_align_cam_floats(void);          /* floats per camera record in cam_out: R(9) t(3) f cx cy A B bf pad(2) */
ST3R_API int st3r_align_img_const_bytes(void);
/* Implementation variants for subsequent ALIGN calls of this process (bit mask, 0 = default).  Bit 0: the
 * per-correspondence loss kernels of st3r_align_optimize keep the gradients of the current image pair in registers
 * along a contiguous entry range and spread the CTA sums over replicated tables (default: every row of 32 entries
 * is reduced across the warp).  Bit 1: st3r_focal_weiszfeld runs one 8-CTA thread-block cluster per image (default:
 * one CTA).  Bit 2: the whole loop of st3r_align_optimize is ONE cooperative launch - optimiser state replicated in the
 * shared memory of every CTA, entries packed once per call, one grid barrier per iteration (needs the workspace of
 * st3r_align_ws_bytes_for and at most 64 images; otherwise the call falls back to three launches per iteration).
 * Same arithmetic per element; only the floating-point summation order differs. */
ST3R_API int st3r_align_set_variant(int variant);
/* Runs `niter` iterations (mode 0: loss_3d, mode 1: loss_2d; both + dust3r_w * loss_dust3r) on the raw parameters
 * pp [N,2] (normalised principal points), log_focal [N], quat [N,4] (XYZW), trans [N,3], log_size [N]; Adam moments
 * adam_m / adam_v are [N,11] in that parameter order (zero them to start a phase).  train_mask bits: 1 pp, 2
 * log_focal, 4 quat, 8 trans, 16 log_size.  h_lr is a HOST array [niter] (cosine schedule).  Outputs (device,
 * optional): loss_hist [niter]; cam_out [N, st3r_align_cam_floats()] = camera records of the LAST forward (the
 * reference returns the state before the final step, reconstruct.py:379-380,405-406); pts3d_out [n_anchor,3];
 * depth_out [n_core_total]; grad_out [N,11] = parameter gradients of the last iteration.  niter = 0 only evaluates. */
ST3R_API int st3r_align_optimize(const St3rAlignProblem* prob, float* pp, float* log_focal, float* quat, float* trans,
                        float* log_size, float* adam_m, float* adam_v, int mode, int train_mask, float gamma,
                        float gamma_dust3r, float dust3r_w, const float* h_lr, int niter, double beta1,
                        double beta2, double eps, float* loss_hist, float* cam_out, float* pts3d_out,
                        float* depth_out, float* grad_out, void* ws, size_t ws_bytes, cudaStream_t stream);

/* canonical_view(ptmaps11, confs11, subsample, mode='avg-angle') (sparse_ga.py:817-855): ptmaps [P,H,W,3], confs
 * [P,H,W] -> canon [H,W,3], canon2 [H,W] (relative depth w.r.t. the 8x8 block centre), cconf [H,W]. */
ST3R_API int st3r_canonical_view(const float* ptmaps, const float* confs, int n_entries, int H, int W, int subsample,
                        float* canon, float* canon2, float* cconf, cudaStream_t stream);
/* estimate_focal_knowing_depth(canon, pp = image centre, 'weiszfeld', min_focal, max_focal) (dust3r/post_process.py:
 * 12-60) for n_img stacked [H,W,3] point maps. */
ST3R_API int st3r_focal_weiszfeld(const float* canon, int n_img, int H, int W, float min_focal, float max_focal, float* focal_out,
                         cudaStream_t stream);
/* SparseGA.get_dense_pts3d for one image (sparse_ga.py:70-93): every pixel anchored to its block's core depth.
 * h_cam2w (16 floats) and h_K (9 floats) are HOST arrays. */
ST3R_API int st3r_dense_points(const float* canon2, const float* core_depth, const float* h_cam2w, const float* h_K,
                      float base_focal, int H, int W, int subsample, float* pts3d, float* depth, cudaStream_t stream);
/* clean_pointcloud (dust3r/cloud_opt/base_opt.py:369-405) for n_img same-size views: pts3d [N,HW,3], confs [N,HW]
 * (updated in place), depthmaps [N,HW], cams [N, st3r_clean_cam_floats()] = world->camera R (9), t (3), K (9). */
ST3R_API int st3r_clean_cam_floats(void);
ST3R_API int st3r_clean_pointcloud(const float* pts3d, float* confs, const float* depthmaps, const float* cams, int n_img, int H,
                          int W, float tol, float bad_conf, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STARST3R_B200_H_ */
