/*
 * Plain-C CPU restatement of the integer/index part of the hot path — TEST
 * INFRASTRUCTURE ONLY (see oracle/ms_hgnn_oracle.py for the rules).
 *
 *   gn_oracle_corr    F.normalize(x, p=2, dim=2) followed by q @ q^T
 *                     (model/GroupNet_nba.py:284-286)
 *   gn_oracle_topk_h  init_adj_attention (model/MS_HGNN_batch.py:372-388):
 *                     scale == N -> one all-ones hyperedge (:375-377);
 *                     otherwise k = max(scale,1) (:378-380), per row the k
 *                     largest correlations are marked 1 (:382-385).
 * Ties are broken towards the lower agent index (the rule the CUDA kernel
 * documents); on tie-free rows this equals torch.topk + scatter.
 *
 * Parity status: pinned by tests/test_oracle.py against the golden vectors
 * generated from the live reference (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

int gn_oracle_corr(const float* x, int B, int N, int D, float* corr) {
  for (int b = 0; b < B; ++b) {
    const float* xb = x + (size_t)b * N * D;
    float* cb = corr + (size_t)b * N * N;
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < N; ++j) {
        /* float accumulation like the fp32 reference; the norm uses
         * max(||x||, 1e-12) as F.normalize does */
        float ni = 0.f, nj = 0.f;
        for (int k = 0; k < D; ++k) {
          ni += xb[i * D + k] * xb[i * D + k];
          nj += xb[j * D + k] * xb[j * D + k];
        }
        ni = fmaxf(sqrtf(ni), 1e-12f);
        nj = fmaxf(sqrtf(nj), 1e-12f);
        float s = 0.f;
        for (int k = 0; k < D; ++k) s += (xb[i * D + k] / ni) * (xb[j * D + k] / nj);
        cb[i * N + j] = s;
      }
    }
  }
  return 0;
}

/* returns 0, or -3 when scale > N (the reference raises at :382) */
int gn_oracle_topk_h(const float* corr, int B, int N, int scale, float* H) {
  if (scale > N) return -3;
  if (scale == N) {
    for (size_t i = 0; i < (size_t)B * N; ++i) H[i] = 1.0f;
    return 0;
  }
  int k = scale < 1 ? 1 : scale;
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < N; ++i) {
      const float* row = corr + ((size_t)b * N + i) * N;
      float* h = H + ((size_t)b * N + i) * N;
      for (int n = 0; n < N; ++n) {
        /* rank of n = number of agents that come before it in
         * (value desc, index asc) order */
        int rank = 0;
        for (int m = 0; m < N; ++m)
          if (row[m] > row[n] || (row[m] == row[n] && m < n)) ++rank;
        h[n] = rank < k ? 1.0f : 0.0f;
      }
    }
  }
  return 0;
}
