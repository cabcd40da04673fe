// TEST-ONLY: the ALIGN optimiser kernels of starst3r_b200/csrc/align.cu (camera forward / backward incl. Adam, the
// three per-correspondence loss kernels and their segmented variant) compiled for the host and run by the SIMT
// emulator in simt_emu.h.  One call = one optimiser iteration with the launch sequence of st3r_align_optimize.
// Never linked into the product library.
#include "simt_emu.h"
#define ST3R_HOST_EMU 1
static float g_dyn_smem[24 * 1024];                 // the 96 KB dynamic shared memory of the loss kernels
#define ST3R_DYN_SMEM(name) float* name = g_dyn_smem
#include "../../starst3r_b200/csrc/align.cu"

namespace {
template <typename K>
int run_grid(int blocks, int threads, K&& body) {
  emu::g_blockDim = dim3(threads, 1, 1);
  emu::g_gridDim = dim3(blocks, 1, 1);
  for (int b = 0; b < blocks; ++b) {
    emu::g_blockIdx = uint3{(unsigned)b, 0, 0};
    if (!emu::run_cta(threads, body)) return -1;
  }
  return 0;
}
void gamma_consts(float gm, float* off, float* offpow) {
  if (gm == 1.0f) { *off = 0.f; *offpow = 0.f; return; }
  double o = pow(1.0 / (double)gm, 1.0 / ((double)gm - 1.0));
  *off = (float)o;
  *offpow = (float)pow(o, (double)gm);
}
}  // namespace

extern "C" {

int emu_align_cam_grads(void) { return NG; }

// One iteration (step = 1-based Adam step).  variant 0: per-row loss kernels, global camera records; 1: segmented loss
// kernels + replicated tables + staged camera kernels.  seg_blocks / seg_per_warp override the launch shape of the
// segmented kernels when > 0 (to force ranges that straddle image pairs).  Returns 0, or -1 on an emulator deadlock.
int emu_align_iteration(int variant, const St3rAlignProblem* prob, float* pp, float* log_focal, float* quat, float* trans,
                        float* log_size, float* adam_m, float* adam_v, int mode, int train_mask, float gamma,
                        float gamma_dust3r, float dust3r_w, float lr, int step, double beta1, double beta2, double eps,
                        float* loss_out, float* grad_out, int seg_blocks, int seg_per_warp) {
  const St3rAlignProblem pb = *prob;
  const int N = pb.n_img;
  std::vector<AlignCam> cam(N);
  std::vector<AlignCamTmp> tmp(N);
  std::vector<AlignCamGrad> cgrad(N);
  std::vector<float> gcam((size_t)N * NG * ALIGN_REPL, 0.f), sums(4, 0.f), gscal(4, 0.f);
  Work w{cam.data(), tmp.data(), cgrad.data(), gcam.data(), sums.data(), gscal.data()};
  Params p{pp, log_focal, quat, trans, log_size};
  AdamState ad{adam_m, adam_v};
  float off_m, offp_m, off_d, offp_d;
  gamma_consts(gamma, &off_m, &offp_m);
  gamma_consts(gamma_dust3r, &off_d, &offp_d);
  const int n_main = mode == 0 ? pb.n3 : pb.n2;
  const float norm_main = mode == 0 ? pb.norm3 : pb.norm2;
  const float scale_main = (n_main > 0 && norm_main != 0.f) ? 1.0f / norm_main : 0.f;
  const float scale_d = (pb.nd > 0 && pb.normd != 0.f) ? 1.0f / pb.normd : 0.f;
  const int reps = variant == 1 ? ALIGN_REPL : 1;
  const bool stage = variant == 1 && N <= CAM_STAGE_MAX;
  constexpr int WARPS = LOSS_THREADS / 32;
  auto blocks_for = [](int n) { int b = (n + LOSS_THREADS * 2 - 1) / (LOSS_THREADS * 2); return b < 1 ? 1 : (b > 1184 ? 1184 : b); };
  auto segb = [&](int n) {
    if (seg_blocks > 0) return seg_blocks;
    int b = (n + WARPS * SEG_MIN_PER_WARP - 1) / (WARPS * SEG_MIN_PER_WARP);
    return b < 1 ? 1 : (b > SEG_MAX_CTAS ? SEG_MAX_CTAS : b);
  };
  auto segpw = [&](int n, int blocks) {
    if (seg_per_warp > 0) return seg_per_warp;
    const long long warps = (long long)blocks * WARPS;
    return (int)(((n + warps - 1) / warps + 31) / 32 * 32);
  };
  int rc;
  if (stage) rc = run_grid(1, CAM_THREADS, [&]() { align_cam_fwd_kernel<true>(pb, p, w); });
  else rc = run_grid(1, CAM_THREADS, [&]() { align_cam_fwd_kernel<false>(pb, p, w); });
  if (rc) return rc;
  if (n_main > 0) {
    if (variant == 1) {
      const int gb = segb(n_main), pw = segpw(n_main, gb);
      if ((long long)gb * WARPS * pw < n_main) return -2;      // the forced launch shape does not cover the entries
      if (mode == 0) rc = run_grid(gb, LOSS_THREADS, [&]() { align_loss_seg_kernel<0>(pb, w, gamma, off_m, offp_m, scale_main, pw, w.sums + 0); });
      else rc = run_grid(gb, LOSS_THREADS, [&]() { align_loss_seg_kernel<1>(pb, w, gamma, off_m, offp_m, scale_main, pw, w.sums + 0); });
    } else if (mode == 0) {
      rc = run_grid(blocks_for(n_main), LOSS_THREADS, [&]() { align_loss3d_kernel(pb, w, gamma, off_m, offp_m, scale_main); });
    } else {
      rc = run_grid(blocks_for(n_main), LOSS_THREADS, [&]() { align_loss2d_kernel(pb, w, gamma, off_m, offp_m, scale_main); });
    }
    if (rc) return rc;
  }
  if (pb.nd > 0 && dust3r_w != 0.f) {
    if (variant == 1) {
      const int gb = segb(pb.nd), pw = segpw(pb.nd, gb);
      if ((long long)gb * WARPS * pw < pb.nd) return -2;
      rc = run_grid(gb, LOSS_THREADS, [&]() { align_loss_seg_kernel<2>(pb, w, gamma_dust3r, off_d, offp_d, scale_d * dust3r_w, pw, w.sums + 1); });
    } else {
      rc = run_grid(blocks_for(pb.nd), LOSS_THREADS, [&]() { align_lossd_kernel(pb, w, gamma_dust3r, off_d, offp_d, scale_d * dust3r_w); });
    }
    if (rc) return rc;
  }
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float a1 = (float)((double)lr / bc1), a2 = (float)(1.0 / sqrt(bc2));
  if (stage)
    rc = run_grid(1, CAM_THREADS, [&]() {
      align_cam_bwd_kernel<true>(pb, p, ad, w, train_mask, a1, a2, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                 (float)eps, 1.0f, loss_out, 0, grad_out, reps);
    });
  else
    rc = run_grid(1, CAM_THREADS, [&]() {
      align_cam_bwd_kernel<false>(pb, p, ad, w, train_mask, a1, a2, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                  (float)eps, 1.0f, loss_out, 0, grad_out, reps);
    });
  if (rc) return rc;
  for (float v : gcam)
    if (v != 0.f) return -3;      // the backward kernel must leave every replica of the table zeroed for the next iteration
  return 0;
}
}
