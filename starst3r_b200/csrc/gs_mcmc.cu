// gsplat.MCMCStrategy device work (starster/gs.py:43-45 builds it, :146-147 / :163-164 call step_pre_backward /
// step_post_backward(..., lr=1e-3); gsplat 1.4 strategy/mcmc.py + strategy/ops.py):
//   * inject_noise_to_position : means += Sigma * (randn * op_sigmoid(1 - sigmoid(opacity)) * lr * noise_lr)
//   * compute_relocation       : new (opacity, scale) of a Gaussian split `ratio` ways (binomial-series formula of
//                                "3D Gaussian Splatting as Markov Chain Monte Carlo")
//   * relocate / sample_add    : dead/alive partition, bincount of the sampled sources, new opacity/scale written to
//                                the sources, source rows copied to the dead (or appended) rows, Adam moments reset.
// Random numbers stay with the caller (torch.randn / torch.multinomial on the reference's generator), so the RNG
// stream is the reference's; the kernels are deterministic given those draws.
// All of it is HBM-bound streaming: noise 56 B read + 12 B written per Gaussian; relocation moves one row of every
// tensor (83 floats) per relocated Gaussian.
#include "common.cuh"
#include "gs.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256)
mcmc_noise_kernel(float* __restrict__ means, const float* __restrict__ quats, const float* __restrict__ scales,
                  const float* __restrict__ opac, const float* __restrict__ noise, int N, float scaler) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float4 q4 = reinterpret_cast<const float4*>(quats)[i];
  float w = q4.x, x = q4.y, y = q4.z, z = q4.w;
  const float inv = 1.0f / fmaxf(sqrtf(w * w + x * x + y * y + z * z), 1e-12f);   // F.normalize
  w *= inv; x *= inv; y *= inv; z *= inv;
  const float R[9] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - w * z), 2.f * (x * z + w * y),
                      2.f * (x * y + w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - w * x),
                      2.f * (x * z - w * y), 2.f * (y * z + w * x), 1.f - 2.f * (x * x + y * y)};
  const float s0 = expf(scales[3 * i]), s1 = expf(scales[3 * i + 1]), s2 = expf(scales[3 * i + 2]);
  const float o = sigmoidf_(opac[i]);
  const float gate = 1.0f / (1.0f + expf(-100.0f * ((1.0f - o) - 0.995f)));   // op_sigmoid(1 - o)
  const float n0 = noise[3 * i] * gate * scaler, n1 = noise[3 * i + 1] * gate * scaler,
              n2 = noise[3 * i + 2] * gate * scaler;
  // Sigma n = M (M^T n), M = R diag(s)
  const float t0 = s0 * (R[0] * n0 + R[3] * n1 + R[6] * n2);
  const float t1 = s1 * (R[1] * n0 + R[4] * n1 + R[7] * n2);
  const float t2 = s2 * (R[2] * n0 + R[5] * n1 + R[8] * n2);
  means[3 * i] += R[0] * s0 * t0 + R[1] * s1 * t1 + R[2] * s2 * t2;
  means[3 * i + 1] += R[3] * s0 * t0 + R[4] * s1 * t1 + R[5] * s2 * t2;
  means[3 * i + 2] += R[6] * s0 * t0 + R[7] * s1 * t1 + R[8] * s2 * t2;
}

// gsplat compute_relocation for one Gaussian: opacity o split n ways.
__device__ __forceinline__ void relocation_one(float o, int n, const float* __restrict__ binoms, int n_max,
                                               float* new_o, float* coeff) {
  const float no = 1.0f - powf(1.0f - o, 1.0f / (float)n);
  float denom = 0.f;
  for (int i = 1; i <= n; ++i) {
    float p = no;                                   // no^(k+1)
    for (int k = 0; k <= i - 1; ++k) {
      const float term = ((k & 1) ? -1.0f : 1.0f) / sqrtf((float)(k + 1)) * p;
      denom += binoms[(i - 1) * n_max + k] * term;
      p *= no;
    }
  }
  *new_o = no;
  *coeff = o / denom;
}

__global__ void __launch_bounds__(256)
mcmc_relocation_kernel(const float* __restrict__ opac, const float* __restrict__ scales, const int32_t* __restrict__ ratios,
                       const float* __restrict__ binoms, int n_max, int N, float* __restrict__ new_opac,
                       float* __restrict__ new_scales) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int n = min(max(ratios[i], 1), n_max);
  float no, coeff;
  relocation_one(opac[i], n, binoms, n_max, &no, &coeff);
  new_opac[i] = no;
#pragma unroll
  for (int k = 0; k < 3; ++k) new_scales[3 * i + k] = coeff * scales[3 * i + k];
}

__global__ void __launch_bounds__(256)
mcmc_flag_kernel(const float* __restrict__ opac_raw, int N, float min_opacity, int32_t* __restrict__ flags,
                 float* __restrict__ probs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float o = sigmoidf_(opac_raw[i]);
  probs[i] = o;
  flags[i] = o <= min_opacity ? 1 : 0;
}

__global__ void __launch_bounds__(256)
mcmc_partition_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ cum, const float* __restrict__ probs,
                      int N, int32_t* __restrict__ dead_idx, int32_t* __restrict__ alive_idx,
                      float* __restrict__ alive_probs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int d = cum[i];
  if (flags[i]) {
    dead_idx[d] = i;
  } else {
    alive_idx[i - d] = i;
    alive_probs[i - d] = probs[i];
  }
}

__device__ __forceinline__ int source_of(const int64_t* __restrict__ sampled, const int32_t* __restrict__ alive_idx, int i) {
  const int s = (int)sampled[i];
  return alive_idx ? alive_idx[s] : s;
}

__global__ void __launch_bounds__(256)
mcmc_bincount_kernel(const int64_t* __restrict__ sampled, const int32_t* __restrict__ alive_idx, int n,
                     int32_t* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(counts + source_of(sampled, alive_idx, i), 1);
}

struct RowSeg { float* p; int cols; };
constexpr int MAX_SEGS = 16;
struct RowSegs { RowSeg s[MAX_SEGS]; int n; };

// One warp per relocated Gaussian i: source s = sampled[i], destination d = dst ? dst[i] : dst_base + i.
// Lane 0 computes the new (opacity, scale) of the source (every duplicate of s computes the same values, so the
// racing stores are benign), the warp then copies row s -> row d of every parameter tensor and (relocate only)
// zeroes row s of every Adam moment tensor.
__global__ void __launch_bounds__(256)
mcmc_relocate_kernel(float* __restrict__ opac_raw, float* __restrict__ scales_raw, RowSegs rows, RowSegs moments,
                     const int64_t* __restrict__ sampled, const int32_t* __restrict__ alive_idx,
                     const int32_t* __restrict__ dst, int dst_base, int n, const int32_t* __restrict__ counts,
                     const float* __restrict__ binoms, int n_max, float min_opacity) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int lane = lane_id();
  const int s = source_of(sampled, alive_idx, i);
  const int d = dst ? dst[i] : dst_base + i;
  float raw_o = 0.f, raw_s0 = 0.f, raw_s1 = 0.f, raw_s2 = 0.f;
  if (lane == 0) {
    const int ratio = min(max(counts[s] + 1, 1), n_max);
    const float o = sigmoidf_(opac_raw[s]);
    float no, coeff;
    relocation_one(o, ratio, binoms, n_max, &no, &coeff);
    no = fminf(fmaxf(no, min_opacity), 1.0f - 1.1920928955078125e-07f);
    raw_o = logf(no / (1.0f - no));                                     // torch.logit
    raw_s0 = logf(coeff * expf(scales_raw[3 * s]));
    raw_s1 = logf(coeff * expf(scales_raw[3 * s + 1]));
    raw_s2 = logf(coeff * expf(scales_raw[3 * s + 2]));
  }
  // Only destination rows (dead or appended, never a source) are written here, so every warp sharing a source
  // reads its old values; mcmc_writeback_kernel then gives the source the same new values.
  if (lane == 0) {
    opac_raw[d] = raw_o;
    scales_raw[3 * d] = raw_s0; scales_raw[3 * d + 1] = raw_s1; scales_raw[3 * d + 2] = raw_s2;
  }
  for (int k = 0; k < rows.n; ++k) {
    float* p = rows.s[k].p;
    const int cols = rows.s[k].cols;
    for (int c = lane; c < cols; c += 32) p[(size_t)d * cols + c] = p[(size_t)s * cols + c];
  }
  for (int k = 0; k < moments.n; ++k) {
    float* p = moments.s[k].p;
    const int cols = moments.s[k].cols;
    for (int c = lane; c < cols; c += 32) p[(size_t)s * cols + c] = 0.f;
  }
}

// Phase 1: the sources take the values their (first) destination received.
__global__ void __launch_bounds__(256)
mcmc_writeback_kernel(float* __restrict__ opac_raw, float* __restrict__ scales_raw, const int64_t* __restrict__ sampled,
                      const int32_t* __restrict__ alive_idx, const int32_t* __restrict__ dst, int dst_base, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = source_of(sampled, alive_idx, i);
  const int d = dst ? dst[i] : dst_base + i;
  opac_raw[s] = opac_raw[d];
  scales_raw[3 * s] = scales_raw[3 * d];
  scales_raw[3 * s + 1] = scales_raw[3 * d + 1];
  scales_raw[3 * s + 2] = scales_raw[3 * d + 2];
}

int fill_segs(RowSegs* out, int n, float* const* ptrs, const int* cols, const char* what) {
  ST3R_CHECK_ARG(n >= 0 && n <= MAX_SEGS, "%s: at most %d tensors", what, MAX_SEGS);
  out->n = n;
  for (int i = 0; i < n; ++i) {
    ST3R_CHECK_ARG(ptrs && cols && ptrs[i] && cols[i] > 0, "%s: bad tensor %d", what, i);
    out->s[i] = RowSeg{ptrs[i], cols[i]};
  }
  return ST3R_OK;
}

}  // namespace

extern "C" {

int st3r_mcmc_inject_noise(float* means, const float* quats, const float* scales, const float* opacities,
                           const float* noise, int N, float scaler, cudaStream_t stream) {
  ST3R_CHECK_ARG(N >= 0, "st3r_mcmc_inject_noise: N < 0");
  if (N == 0) return ST3R_OK;
  ST3R_CHECK_ARG(means && quats && scales && opacities && noise, "st3r_mcmc_inject_noise: null pointer");
  mcmc_noise_kernel<<<(N + 255) / 256, 256, 0, stream>>>(means, quats, scales, opacities, noise, N, scaler);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_mcmc_compute_relocation(const float* opacities, const float* scales, const int32_t* ratios, const float* binoms,
                                 int n_max, int N, float* new_opacities, float* new_scales, cudaStream_t stream) {
  ST3R_CHECK_ARG(N >= 0 && n_max >= 1, "st3r_mcmc_compute_relocation: bad sizes");
  if (N == 0) return ST3R_OK;
  ST3R_CHECK_ARG(opacities && scales && ratios && binoms && new_opacities && new_scales,
                 "st3r_mcmc_compute_relocation: null pointer");
  mcmc_relocation_kernel<<<(N + 255) / 256, 256, 0, stream>>>(opacities, scales, ratios, binoms, n_max, N,
                                                               new_opacities, new_scales);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

size_t st3r_mcmc_partition_ws_bytes(int N) {
  const size_t n = (size_t)(N > 0 ? N : 1);
  return st3r_align_up(n * 4, 256) * 2 + st3r_align_up(st3r_scan_ws_bytes(n), 256) + 1024;
}

int st3r_mcmc_partition(const float* opacities_raw, int N, float min_opacity, int32_t* dead_idx, int32_t* alive_idx,
                        float* probs, float* alive_probs, int32_t* n_dead, void* ws, size_t ws_bytes,
                        cudaStream_t stream) {
  ST3R_CHECK_ARG(N >= 0, "st3r_mcmc_partition: N < 0");
  ST3R_CHECK_ARG(n_dead, "st3r_mcmc_partition: null n_dead");
  if (N == 0) {
    ST3R_CHECK_CUDA(cudaMemsetAsync(n_dead, 0, sizeof(int32_t), stream));
    return ST3R_OK;
  }
  ST3R_CHECK_ARG(opacities_raw && dead_idx && alive_idx && probs && alive_probs && ws, "st3r_mcmc_partition: null pointer");
  WsAlloc a(ws, ws_bytes);
  int32_t* flags = a.take<int32_t>(N);
  int32_t* cum = a.take<int32_t>(N);
  const size_t scan_bytes = st3r_scan_ws_bytes(N);
  char* scan_ws = a.take<char>(scan_bytes);
  if (!a.ok()) {
    st3r_set_error("st3r_mcmc_partition: workspace too small (%zu < %zu)", ws_bytes, a.off);
    return ST3R_ERR_WORKSPACE;
  }
  mcmc_flag_kernel<<<(N + 255) / 256, 256, 0, stream>>>(opacities_raw, N, min_opacity, flags, probs);
  ST3R_CHECK_LAUNCH();
  int rc = st3r_exclusive_scan_i32(flags, cum, N, n_dead, scan_ws, scan_bytes, stream);
  if (rc != ST3R_OK) return rc;
  mcmc_partition_kernel<<<(N + 255) / 256, 256, 0, stream>>>(flags, cum, probs, N, dead_idx, alive_idx, alive_probs);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

int st3r_mcmc_relocate(float* opacities_raw, float* scales_raw, int n_rows, float* const* h_row_ptrs,
                       const int* h_row_cols, int n_moments, float* const* h_moment_ptrs, const int* h_moment_cols,
                       const int64_t* sampled, const int32_t* alive_idx, const int32_t* dst, int dst_base, int n, int N,
                       const float* binoms, int n_max, float min_opacity, int32_t* counts, cudaStream_t stream) {
  ST3R_CHECK_ARG(n >= 0 && N >= 0 && n_max >= 1, "st3r_mcmc_relocate: bad sizes");
  if (n == 0) return ST3R_OK;
  ST3R_CHECK_ARG(opacities_raw && scales_raw && sampled && binoms && counts, "st3r_mcmc_relocate: null pointer");
  RowSegs rows, moments;
  int rc = fill_segs(&rows, n_rows, h_row_ptrs, h_row_cols, "st3r_mcmc_relocate(rows)");
  if (rc != ST3R_OK) return rc;
  rc = fill_segs(&moments, n_moments, h_moment_ptrs, h_moment_cols, "st3r_mcmc_relocate(moments)");
  if (rc != ST3R_OK) return rc;
  ST3R_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)N, stream));
  mcmc_bincount_kernel<<<(n + 255) / 256, 256, 0, stream>>>(sampled, alive_idx, n, counts);
  ST3R_CHECK_LAUNCH();
  const long long threads = (long long)n * 32;
  mcmc_relocate_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
      opacities_raw, scales_raw, rows, moments, sampled, alive_idx, dst, dst_base, n, counts, binoms, n_max, min_opacity);
  ST3R_CHECK_LAUNCH();
  mcmc_writeback_kernel<<<(n + 255) / 256, 256, 0, stream>>>(opacities_raw, scales_raw, sampled, alive_idx, dst, dst_base, n);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
}
