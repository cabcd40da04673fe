/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the dot-product nearest neighbour of
 * mast3r/mast3r/fast_nn.py:16-70 (dist='dot'): scores = A @ B.T, row arg-max
 * (torch.max -> first maximal index; the blocked variant keeps the earlier block
 * on ties, fast_nn.py:60-61, i.e. the lowest index overall).
 * Score definition: sequential fp32 FMA chain over k (what MKL sgemm computes
 * for K=24 on the CPU the reference runs on; pinned against the reference's
 * own output in tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>

void oracle_nn_argmax_dot(const float* A, int M, const float* B, int N, int d, int32_t* idx, float* best) {
  for (int i = 0; i < M; ++i) {
    const float* a = A + (long)i * d;
    float bs = -INFINITY;
    int32_t bj = -1;
    for (int j = 0; j < N; ++j) {
      const float* b = B + (long)j * d;
      float s = 0.f;
      for (int k = 0; k < d; ++k) s = fmaf(a[k], b[k], s);
      if (s > bs) { bs = s; bj = j; }
    }
    idx[i] = bj;
    if (best) best[i] = bs;
  }
}
