// Fused per-Gaussian Adam (replaces the six torch.optim.Adam instances of starster/gs.py:37,159-161;
// lr 1e-3, betas (0.9, 0.999), eps 1e-8, no weight decay, bias-corrected, non-amsgrad).
// One launch updates every trainable tensor of the splat: segment s = (param, grad, exp_avg, exp_avg_sq)
// viewed as `rows` x `cols` with independent leading dimensions, so the first 4 SH coefficients of shN
// ([N,24,3], the only ones that ever receive a non-zero gradient: gs.py:81 renders with sh_degree=1)
// are updated in place without touching the other 60 floats.  `sh0` never receives a gradient
// (it is not passed to the renderer) and is skipped exactly like torch skips params with grad=None.
// HBM-bound: 7 floats moved per element (read p, g, m, v; write p, m, v) = 28 B.
#include "common.cuh"
#include "gs.cuh"

namespace {
struct AdamSeg {
  float* p; const float* g; float* m; float* v;
  int rows, cols, ld_p, ld_g;
  int vec;      // host-side verdict: every row can be moved as float4 (cols, leading dimensions and addresses multiples of 4)
};
struct AdamSegs { AdamSeg s[8]; int n; };

// Contiguous segments are presented as ONE row (no index division in the kernel); a segment is moved as float4 when its
// row length, leading dimensions and base addresses allow it.  `g_off`: element offset of the gradients inside the
// peers' buffers (st3r_adam_step_peers), 0 otherwise.
inline AdamSeg make_seg(float* p, const float* g, float* m, float* v, int rows, int cols, int ld_p, int ld_g,
                        long long g_off, bool g_aligned) {
  if (ld_p == cols && ld_g == cols && (long long)rows * cols < (1ll << 31)) {
    cols = ld_p = ld_g = rows * cols;
    rows = cols > 0 ? 1 : 0;
  }
  const bool al = ((uintptr_t)p % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0 && g_aligned &&
                  ((uintptr_t)g % 16) == 0 && (g_off % 4) == 0;
  // a single row may end in up to three scalar elements; several rows need whole float4s per row
  const bool shape = rows <= 1 ? true : ((cols | ld_p | ld_g) % 4) == 0;
  const bool small = (long long)rows * cols < (1ll << 31);
  return AdamSeg{p, g, m, v, rows, cols, ld_p, ld_g, (al && shape && small) ? 1 : 0};
}

struct AdamCoef { float lr_over_bc1, inv_sqrt_bc2, one_minus_b1, b2, one_minus_b2, eps; };

// Every operation is pinned (no compiler-chosen FMA contraction), so the float4, scalar and tail paths - and the host
// emulation of the tests - produce the same bits for the same element.
#ifdef ST3R_HOST_EMU
static inline float ad_fma(float a, float b, float c) { return fmaf(a, b, c); }
static inline float ad_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float ad_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float ad_div(float a, float b) { volatile float r = a / b; return r; }
#else
__device__ __forceinline__ float ad_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float ad_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float ad_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ad_div(float a, float b) { return __fdiv_rn(a, b); }
#endif
__device__ __forceinline__ void adam_update(float& p, const float g, float& m, float& v, const AdamCoef& k) {
  m = ad_fma(ad_sub(g, m), k.one_minus_b1, m);                        // exp_avg.lerp_(grad, 1 - beta1)
  v = ad_fma(ad_mul(k.one_minus_b2, g), g, ad_mul(v, k.b2));          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = ad_fma(sqrtf(v), k.inv_sqrt_bc2, k.eps);
  p = ad_fma(-k.lr_over_bc1, ad_div(m, denom), p);
}

// Gradient sources: the rank's own buffer, or the rank-ordered sum over the peers' symmetric buffers.
struct LocalGrad {
  const float* g;
  __device__ __forceinline__ float load(size_t i) const { return g[i]; }
  __device__ __forceinline__ float4 load4(size_t i) const { return *reinterpret_cast<const float4*>(g + i); }
};
constexpr int MAX_PEERS = 8;
struct PeerGrads { const float* base[MAX_PEERS]; int world; };
struct PeerOffsets { long long off[8]; };
struct PeerSumGrad {
  const PeerGrads& peers;
  long long base;
  __device__ __forceinline__ float load(size_t i) const {
    float g = 0.f;
#pragma unroll
    for (int k = 0; k < MAX_PEERS; ++k)
      if (k < peers.world) g += __ldcv(peers.base[k] + base + (long long)i);   // volatile-cached: never a stale L1 line of peer memory
    return g;
  }
  __device__ __forceinline__ float4 load4(size_t i) const {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < MAX_PEERS; ++k)
      if (k < peers.world) {
        const float4 x = __ldcv(reinterpret_cast<const float4*>(peers.base[k] + base + (long long)i));
        g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
      }
    return g;
  }
};

// One segment, grid-stride.  HBM-bound: 7 floats moved per element; 16-byte accesses wherever the layout allows, scalar
// for rows that are not multiples of 4 floats and for the up-to-3-element tail of a contiguous segment.
template <class G>
__device__ __forceinline__ void adam_segment(const AdamSeg& sg, const G& gl, const AdamCoef& k) {
  const unsigned stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (sg.vec) {
    const unsigned c4 = (unsigned)sg.cols >> 2, total4 = (unsigned)sg.rows * c4;
    for (unsigned i = i0; i < total4; i += stride) {
      size_t ip, ig;
      if (sg.rows == 1) {
        ip = ig = (size_t)i * 4;
      } else {
        const unsigned r = i / c4, c = (i - r * c4) * 4;
        ip = (size_t)r * sg.ld_p + c;
        ig = (size_t)r * sg.ld_g + c;
      }
      const float4 g = gl.load4(ig);
      float4 p = *reinterpret_cast<float4*>(sg.p + ip), m = *reinterpret_cast<float4*>(sg.m + ip),
             v = *reinterpret_cast<float4*>(sg.v + ip);
      adam_update(p.x, g.x, m.x, v.x, k);
      adam_update(p.y, g.y, m.y, v.y, k);
      adam_update(p.z, g.z, m.z, v.z, k);
      adam_update(p.w, g.w, m.w, v.w, k);
      *reinterpret_cast<float4*>(sg.p + ip) = p;
      *reinterpret_cast<float4*>(sg.m + ip) = m;
      *reinterpret_cast<float4*>(sg.v + ip) = v;
    }
    if (sg.rows == 1) {                                 // tail of a single row
      const unsigned i = (c4 << 2) + i0;
      if (i < (unsigned)sg.cols) {
        float p = sg.p[i], m = sg.m[i], v = sg.v[i];
        adam_update(p, gl.load(i), m, v, k);
        sg.p[i] = p; sg.m[i] = m; sg.v[i] = v;
      }
    }
    return;
  }
  const long long total = (long long)sg.rows * sg.cols;
  for (long long i = i0; i < total; i += stride) {
    size_t r, c;
    if (total < (1ll << 31)) { r = (unsigned)i / (unsigned)sg.cols; c = (unsigned)i - (unsigned)r * (unsigned)sg.cols; }
    else { r = (size_t)(i / sg.cols); c = (size_t)(i - (long long)r * sg.cols); }
    const size_t ip = r * sg.ld_p + c, ig = r * sg.ld_g + c;
    float p = sg.p[ip], m = sg.m[ip], v = sg.v[ip];
    adam_update(p, gl.load(ig), m, v, k);
    sg.p[ip] = p; sg.m[ip] = m; sg.v[ip] = v;
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(AdamSegs segs, AdamCoef k) {
  const AdamSeg sg = segs.s[blockIdx.y];
  adam_segment(sg, LocalGrad{sg.g}, k);
}
// The same update with the step number kept ON THE DEVICE (steps_done = number of completed steps): the bias
// corrections are evaluated by thread 0 of every block in double precision exactly as the host entry point does, so the
// launch carries no per-step scalar and a captured CUDA graph of the whole training iteration can be replayed
// unchanged (gs.TrainPlan's graph mode).  adam_advance_kernel increments the counter after the update.
__global__ void __launch_bounds__(256)
adam_dev_kernel(AdamSegs segs, const int* __restrict__ steps_done, double lr, double beta1, double beta2, AdamCoef k) {
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double step = (double)(*steps_done + 1);
    const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
    s_bc[0] = (float)(lr / bc1);
    s_bc[1] = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  k.lr_over_bc1 = s_bc[0];
  k.inv_sqrt_bc2 = s_bc[1];
  const AdamSeg sg = segs.s[blockIdx.y];
  adam_segment(sg, LocalGrad{sg.g}, k);
}
__global__ void adam_advance_kernel(int* steps_done) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *steps_done += 1;
}
// Fused gradient all-reduce + Adam over peer memory (multi-GPU training, SURVEY.md §8e: views are sharded, the splat is
// replicated, the only exchange is the sum of the per-Gaussian gradients).  Every rank's gradients live at the same
// offsets of a symmetric buffer that all ranks map (NVLink P2P through NVSwitch); this kernel reads element i of every
// peer, adds them in rank order (so all replicas compute bit-identical sums and stay in lock-step) and applies the
// Adam update in the same pass: the reduced gradient never exists in HBM and there is no separate collective.
__global__ void __launch_bounds__(256)
adam_peer_kernel(AdamSegs segs, PeerGrads peers, PeerOffsets goff, AdamCoef k) {
  const AdamSeg sg = segs.s[blockIdx.y];
  adam_segment(sg, PeerSumGrad{peers, goff.off[blockIdx.y]}, k);
}
// Reduce-scatter + all-gather form of the same exchange for larger node sizes: rank r sums elements [r L/G, (r+1) L/G)
// of every peer's gradient buffer (rank order, so the sums are bit-identical to adam_peer_kernel's) and stores the
// result into EVERY peer's `reduced` buffer.  Remote traffic per GPU: (G-1)/G L loads + (G-1)/G L stores instead of
// (G-1) L loads; the Adam kernel then runs on local memory.
__global__ void __launch_bounds__(256)
grad_reduce_scatter_kernel(PeerGrads grads, PeerGrads reduced_rw, long long begin4, long long end4) {
  for (long long i = begin4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < MAX_PEERS; ++k)
      if (k < grads.world) {
        const float4 v = __ldcv(reinterpret_cast<const float4*>(grads.base[k]) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
#pragma unroll
    for (int k = 0; k < MAX_PEERS; ++k)
      if (k < grads.world) reinterpret_cast<float4*>(const_cast<float*>(reduced_rw.base[k]))[i] = acc;
  }
}
// The same exchange through the NVSwitch itself (NVLS): the gradient buffers of all ranks are bound to one multicast
// address; multimem.ld_reduce makes the switch fetch element i from every GPU and return the SUM (one 16-byte response
// instead of G - 1 remote loads), multimem.st writes the result into every GPU's `reduced` buffer with one store.  Rank r
// handles elements [r L / G, (r + 1) L / G): per GPU L / G floats received and L / G sent instead of 2 (G - 1) / G L.
// Every rank receives the value its owner broadcast, so the replicas stay bit-identical; the order of the additions
// inside the switch is not specified, so the sum may differ from the rank-ordered one in the last bit.
#ifndef ST3R_HOST_EMU
constexpr int MM_UNROLL = 4;      // independent in-switch reductions in flight per thread (the round trip is microseconds)
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* p, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :
               : "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__global__ void __launch_bounds__(256)
grad_reduce_multimem_kernel(const float* mc_grads, float* mc_reduced, long long begin4, long long end4) {
  const float4* src = reinterpret_cast<const float4*>(mc_grads);
  float4* dst = reinterpret_cast<float4*>(mc_reduced);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = begin4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end4; i += stride * MM_UNROLL) {
    float4 v[MM_UNROLL];
#pragma unroll
    for (int k = 0; k < MM_UNROLL; ++k)
      if (i + k * stride < end4) v[k] = multimem_ld_reduce_add(src + i + k * stride);
#pragma unroll
    for (int k = 0; k < MM_UNROLL; ++k)
      if (i + k * stride < end4) multimem_st(dst + i + k * stride, v[k]);
  }
}
#endif
}  // namespace

extern "C" int st3r_grad_reduce_multimem(int world, int rank, const float* mc_grads, float* mc_reduced, int64_t n_floats,
                                         cudaStream_t stream) {
  ST3R_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "st3r_grad_reduce_multimem: bad world / rank");
  ST3R_CHECK_ARG(n_floats >= 0 && n_floats % 4 == 0, "st3r_grad_reduce_multimem: length must be a multiple of 4 floats");
  if (n_floats == 0) return ST3R_OK;
  ST3R_CHECK_ARG(mc_grads && mc_reduced && ((uintptr_t)mc_grads % 16) == 0 && ((uintptr_t)mc_reduced % 16) == 0,
                 "st3r_grad_reduce_multimem: multicast addresses must be non-null and 16-byte aligned");
#ifdef ST3R_HOST_EMU
  st3r_set_error("st3r_grad_reduce_multimem: needs NVSwitch multicast memory (not available in the host emulation)");
  return ST3R_ERR_UNSUPPORTED;
#else
  const long long n4 = n_floats / 4, chunk = (n4 + world - 1) / world;
  const long long begin4 = chunk * rank, end4 = begin4 + chunk < n4 ? begin4 + chunk : n4;
  if (begin4 >= end4) return ST3R_OK;
  long long blocks = (end4 - begin4 + 256 * MM_UNROLL - 1) / (256 * MM_UNROLL);
  const long long cap = (long long)st3r_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  grad_reduce_multimem_kernel<<<(unsigned)blocks, 256, 0, stream>>>(mc_grads, mc_reduced, begin4, end4);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
#endif
}

extern "C" int st3r_grad_reduce_scatter(int world, int rank, const float* const* peer_grad_bases,
                                        float* const* peer_reduced_bases, int64_t n_floats, cudaStream_t stream) {
  ST3R_CHECK_ARG(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world, "st3r_grad_reduce_scatter: bad world / rank");
  ST3R_CHECK_ARG(n_floats >= 0 && n_floats % 4 == 0, "st3r_grad_reduce_scatter: length must be a multiple of 4 floats");
  if (n_floats == 0) return ST3R_OK;
  ST3R_CHECK_ARG(peer_grad_bases && peer_reduced_bases, "st3r_grad_reduce_scatter: null");
  PeerGrads g, r;
  g.world = r.world = world;
  for (int k = 0; k < MAX_PEERS; ++k) {
    g.base[k] = peer_grad_bases[k < world ? k : 0];
    r.base[k] = peer_reduced_bases[k < world ? k : 0];
    ST3R_CHECK_ARG(g.base[k] && r.base[k] && ((uintptr_t)g.base[k] % 16) == 0 && ((uintptr_t)r.base[k] % 16) == 0,
                   "st3r_grad_reduce_scatter: peer buffers must be non-null and 16-byte aligned");
  }
  const long long n4 = n_floats / 4, chunk = (n4 + world - 1) / world;
  const long long begin4 = chunk * rank, end4 = begin4 + chunk < n4 ? begin4 + chunk : n4;
  if (begin4 >= end4) return ST3R_OK;
  long long blocks = (end4 - begin4 + 255) / 256;
  const long long cap = (long long)st3r_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  grad_reduce_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g, r, begin4, end4);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

extern "C" int st3r_adam_step_peers(int n_seg, float* const* params, const long long* grad_offsets, float* const* exp_avg,
                                    float* const* exp_avg_sq, const int* rows, const int* cols, const int* ld_param,
                                    const int* ld_grad, int world, const float* const* peer_grad_bases, double lr,
                                    double beta1, double beta2, double eps, int step, cudaStream_t stream) {
  ST3R_CHECK_ARG(n_seg >= 0 && n_seg <= 8 && step >= 1, "st3r_adam_step_peers: bad args (n_seg <= 8, step >= 1)");
  ST3R_CHECK_ARG(world >= 1 && world <= MAX_PEERS, "st3r_adam_step_peers: world size must be 1..%d", MAX_PEERS);
  if (n_seg == 0) return ST3R_OK;
  ST3R_CHECK_ARG(params && grad_offsets && exp_avg && exp_avg_sq && rows && cols && ld_param && ld_grad && peer_grad_bases,
                 "st3r_adam_step_peers: null");
  AdamSegs segs;
  PeerGrads peers;
  PeerOffsets goff;
  segs.n = n_seg;
  peers.world = world;
  for (int k = 0; k < MAX_PEERS; ++k) peers.base[k] = peer_grad_bases[k < world ? k : 0];
  for (int k = 0; k < world; ++k) ST3R_CHECK_ARG(peer_grad_bases[k], "st3r_adam_step_peers: null peer buffer %d", k);
  long long max_total = 0;
  for (int i = 0; i < 8; ++i) goff.off[i] = 0;
  for (int i = 0; i < n_seg; ++i) {
    ST3R_CHECK_ARG(params[i] && exp_avg[i] && exp_avg_sq[i] && rows[i] >= 0 && cols[i] > 0 && grad_offsets[i] >= 0,
                   "st3r_adam_step_peers: bad segment %d", i);
    bool bases_aligned = true;
    for (int k = 0; k < world; ++k) bases_aligned = bases_aligned && ((uintptr_t)peer_grad_bases[k] % 16) == 0;
    segs.s[i] = make_seg(params[i], nullptr, exp_avg[i], exp_avg_sq[i], rows[i], cols[i], ld_param[i], ld_grad[i],
                         grad_offsets[i], bases_aligned);
    goff.off[i] = grad_offsets[i];
    long long t = (long long)rows[i] * cols[i];
    if (t > max_total) max_total = t;
  }
  if (max_total == 0) return ST3R_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  int blocks = (int)((max_total + 255) / 256);
  int cap = st3r_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  adam_peer_kernel<<<dim3(blocks, n_seg), 256, 0, stream>>>(
      segs, peers, goff,
      AdamCoef{(float)(lr / bc1), (float)(1.0 / sqrt(bc2)), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps});
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

extern "C" int st3r_adam_step_dev(int n_seg, float* const* params, const float* const* grads, float* const* exp_avg,
                                  float* const* exp_avg_sq, const int* rows, const int* cols, const int* ld_param,
                                  const int* ld_grad, double lr, double beta1, double beta2, double eps,
                                  int32_t* steps_done, cudaStream_t stream) {
  ST3R_CHECK_ARG(n_seg >= 0 && n_seg <= 8 && steps_done, "st3r_adam_step_dev: bad args (n_seg <= 8, steps_done != NULL)");
  if (n_seg == 0) return ST3R_OK;
  ST3R_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && rows && cols && ld_param && ld_grad, "st3r_adam_step_dev: null");
  AdamSegs segs;
  segs.n = n_seg;
  long long max_total = 0;
  for (int i = 0; i < n_seg; ++i) {
    ST3R_CHECK_ARG(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && rows[i] >= 0 && cols[i] > 0,
                   "st3r_adam_step_dev: bad segment %d", i);
    segs.s[i] = make_seg(params[i], grads[i], exp_avg[i], exp_avg_sq[i], rows[i], cols[i], ld_param[i], ld_grad[i], 0, true);
    long long t = (long long)rows[i] * cols[i];
    if (t > max_total) max_total = t;
  }
  if (max_total > 0) {
    int blocks = (int)((max_total + 255) / 256);
    int cap = st3r_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    adam_dev_kernel<<<dim3(blocks, n_seg), 256, 0, stream>>>(
        segs, steps_done, lr, beta1, beta2,
        AdamCoef{0.f, 0.f, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps});
    ST3R_CHECK_LAUNCH();
  }
  adam_advance_kernel<<<1, 32, 0, stream>>>(steps_done);
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}

extern "C" int st3r_adam_step(int n_seg, float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, const int* rows, const int* cols, const int* ld_param,
                              const int* ld_grad, double lr, double beta1, double beta2, double eps, int step,
                              cudaStream_t stream) {
  ST3R_CHECK_ARG(n_seg >= 0 && n_seg <= 8 && step >= 1, "st3r_adam_step: bad args (n_seg <= 8, step >= 1)");
  if (n_seg == 0) return ST3R_OK;
  ST3R_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && rows && cols && ld_param && ld_grad, "st3r_adam_step: null");
  AdamSegs segs;
  segs.n = n_seg;
  long long max_total = 0;
  for (int i = 0; i < n_seg; ++i) {
    ST3R_CHECK_ARG(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && rows[i] >= 0 && cols[i] > 0,
                   "st3r_adam_step: bad segment %d", i);
    segs.s[i] = make_seg(params[i], grads[i], exp_avg[i], exp_avg_sq[i], rows[i], cols[i], ld_param[i], ld_grad[i], 0, true);
    long long t = (long long)rows[i] * cols[i];
    if (t > max_total) max_total = t;
  }
  if (max_total == 0) return ST3R_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float lr_over_bc1 = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  int blocks = (int)((max_total + 255) / 256);
  int cap = st3r_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  dim3 grid(blocks, n_seg);
  // torch evaluates 1 - beta in double and rounds once to fp32
  adam_kernel<<<grid, 256, 0, stream>>>(
      segs, AdamCoef{lr_over_bc1, inv_sqrt_bc2, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps});
  ST3R_CHECK_LAUNCH();
  return ST3R_OK;
}
