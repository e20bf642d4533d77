// fp32 (FFMA) kernels of one message-passing stage — the 1e-5 parity path.
//
// Stage = node2edge -> MLP_dict_softmax -> edge2node -> closing MLP
// (model/MS_HGNN_batch.py:173-195 / :425-441), restructured so that nothing of
// size (B,E,N,128) is ever materialised:
//
//   k1 node_pre     rows B*N   x' = node2edge_start_mlp(h); pq = W0split x';
//                              pairwise only: P = h W0_agg^T (aggregation collapse)
//   k2 node2edge    per scene  attention softmax over the incidence + weighted
//                              gather: edges (B*E,64); hyper: eo = H h (B*E,D)
//   k3 edge_mlp     rows B*E   init_MLP, [distribution|factor] MLPs, Gumbel
//                              softmax, sigmoid: dist (B,E,T), edge_feat (B*E,T)
//   k4 edge_agg     rows B*E   hyper only: ef = sum_t edge_feat_t * agg_mlp_t(eo)
//   k5 edge2node    per scene  pairwise: G = scatter of relu(P_i+P_j+b)*w (collapsed);
//                              hyper: agg = H^T ef
//   k6 node_post    rows B*N   pairwise: agg = G W1cat^T + S b1; incoming =
//                              [agg | h]/N; closing MLP 2D -> 128 -> Dout
//
// Algebra of the pairwise collapse (validated to 3e-7): the
// first Linear of every agg_mlp commutes with H @ ori and the second with
// H^T @ ., so both run on N rows/scene instead of N^2.
#include "gn_gemm_simt.cuh"

namespace gn {

// ===========================================================================
// k1: node_pre
// ===========================================================================
// smem: hT [Dp][LD] | hidT [128][LD] | xT [64][LD] | wp [2*KC*128]
template <int TM>
__global__ void __launch_bounds__(GN_THREADS)
node_pre_kernel(const float* __restrict__ h, int R, int D, int Dp, int T, int pairwise,
                gn_stage_weights W, float* __restrict__ xprime, float* __restrict__ pq,
                float* __restrict__ P) {
  constexpr int LD = TM + 4, RM = TM / 16;
  extern __shared__ __align__(16) float smem[];
  float* hT = smem;
  float* hidT = hT + Dp * LD;
  float* xT = hidT + 128 * LD;
  float* wp = xT + 64 * LD;
  const int ntiles = (R + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * TM, nrows = min(TM, R - row0);
    load_tile_kmajor<TM>(hT, h + static_cast<size_t>(row0) * D, D, nrows, D, Dp, 1.f, false);
    // x' = W1 relu(W0 h + b0) + b1, hidden processed in two 128-column halves
    float accx[RM][4];
    acc_zero(accx);
    for (int half = 0; half < 2; ++half) {
      float acc1[RM][8];
      acc_zero(acc1);
      gemm_accum<TM, 128>(acc1, hT, W.node_w0t, GN_NODE_HIDDEN, half * 128, Dp, wp);
      const float* b0 = W.node_b0 + half * 128;
      acc_store_kmajor<TM, 128>(acc1, hidT, 0,
                                [&](int col, int, float v) { return fmaxf(v + __ldg(b0 + col), 0.f); });
      gemm_accum<TM, 64>(accx, hidT, W.node_w1t + static_cast<size_t>(half) * 128 * GN_ATT_DIM,
                         GN_ATT_DIM, 0, 128, wp);
    }
    acc_foreach<TM, 64>(accx, [&](int r, int c, float& v) {
      v += __ldg(W.node_b1 + c);
      if (r < nrows) xprime[static_cast<size_t>(row0 + r) * GN_ATT_DIM + c] = v;
    });
    acc_store_kmajor<TM, 64>(accx, xT, 0, [](int, int, float v) { return v; });
    // pq = [W0a x' | W0b x'] (bias added per edge later)
    float accp[RM][4];
    acc_zero(accp);
    gemm_accum<TM, 64>(accp, xT, W.att_wpqt, GN_ATT_DIM, 0, GN_ATT_DIM, wp);
    acc_foreach<TM, 64>(accp, [&](int r, int c, float& v) {
      if (r < nrows) pq[static_cast<size_t>(row0 + r) * GN_ATT_DIM + c] = v;
    });
    if (pairwise) {
      const int ldp = T * GN_MLP_HIDDEN;
      for (int t = 0; t < T; ++t) {
        float accP[RM][8];
        acc_zero(accP);
        gemm_accum<TM, 128>(accP, hT, W.agg_w0t, ldp, t * 128, Dp, wp);
        acc_foreach<TM, 128>(accP, [&](int r, int c, float& v) {
          if (r < nrows) P[static_cast<size_t>(row0 + r) * ldp + t * 128 + c] = v;
        });
      }
    }
    __syncthreads();
  }
}

// ===========================================================================
// k2a: node2edge, pairwise (incidence implicit: edge e = i*N + j, self loops
// carry incidence 2).  Work item = (scene, chunk of EC edges).
// ===========================================================================
constexpr int N2E_LD = GN_ATT_DIM + 4;   // 68: rows 4 banks apart
constexpr int N2E_EC = 1024;

__global__ void __launch_bounds__(GN_THREADS)
node2edge_pair_kernel(const float* __restrict__ xprime, const float* __restrict__ pq,
                      int B, int N, gn_stage_weights W, float* __restrict__ edges) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                       // [N][68]
  float* ps = xs + N * N2E_LD;            // [N][68]  pn | q
  float* wts = ps + N * N2E_LD;           // [EC][2]
  __shared__ float sb0[GN_ATT_HIDDEN], sw1[GN_ATT_HIDDEN];
  const int tid = threadIdx.x;
  if (tid < GN_ATT_HIDDEN) { sb0[tid] = __ldg(W.att_b0 + tid); sw1[tid] = __ldg(W.att_w1 + tid); }
  const float b1 = __ldg(W.att_b1);
  const int E = N * N;
  const int nchunk = (E + N2E_EC - 1) / N2E_EC;
  const long long nwork = static_cast<long long>(B) * nchunk;
  for (long long wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int b = static_cast<int>(wi / nchunk), ch = static_cast<int>(wi - static_cast<long long>(b) * nchunk);
    const int e0 = ch * N2E_EC, ec = min(N2E_EC, E - e0);
    __syncthreads();
    for (int i = tid; i < N * 16; i += GN_THREADS) {
      int n = i >> 4, c = i & 15;
      size_t g = (static_cast<size_t>(b) * N + n) * GN_ATT_DIM + 4 * c;
      *reinterpret_cast<float4*>(xs + n * N2E_LD + 4 * c) = ldg_f4(xprime + g);
      *reinterpret_cast<float4*>(ps + n * N2E_LD + 4 * c) = ldg_f4(pq + g);
    }
    __syncthreads();
    // phase A: one thread per edge -> attention weights of its (<= 2) members
    for (int el = tid; el < ec; el += GN_THREADS) {
      const int e = e0 + el, i = e / N, j = e - i * N;
      const float* pi = ps + i * N2E_LD;
      const float* pj = ps + j * N2E_LD;
      const float hs = (i == j) ? 2.f : 1.f;      // self loop: edge_init = 2 x_i
      float ai = 0.f, aj = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < GN_ATT_HIDDEN; k4 += 4) {
        float4 ni = *reinterpret_cast<const float4*>(pi + k4);
        float4 nj = *reinterpret_cast<const float4*>(pj + k4);
        float4 qi = *reinterpret_cast<const float4*>(pi + 32 + k4);
        float4 qj = *reinterpret_cast<const float4*>(pj + 32 + k4);
        float pe0, pe1, pe2, pe3;
        if (i == j) {
          pe0 = 2.f * qi.x; pe1 = 2.f * qi.y; pe2 = 2.f * qi.z; pe3 = 2.f * qi.w;
        } else {
          pe0 = qi.x + qj.x; pe1 = qi.y + qj.y; pe2 = qi.z + qj.z; pe3 = qi.w + qj.w;
        }
        pe0 += sb0[k4]; pe1 += sb0[k4 + 1]; pe2 += sb0[k4 + 2]; pe3 += sb0[k4 + 3];
        ai = fmaf(fmaxf(ni.x + pe0, 0.f), sw1[k4], ai);
        ai = fmaf(fmaxf(ni.y + pe1, 0.f), sw1[k4 + 1], ai);
        ai = fmaf(fmaxf(ni.z + pe2, 0.f), sw1[k4 + 2], ai);
        ai = fmaf(fmaxf(ni.w + pe3, 0.f), sw1[k4 + 3], ai);
        aj = fmaf(fmaxf(nj.x + pe0, 0.f), sw1[k4], aj);
        aj = fmaf(fmaxf(nj.y + pe1, 0.f), sw1[k4 + 1], aj);
        aj = fmaf(fmaxf(nj.z + pe2, 0.f), sw1[k4 + 2], aj);
        aj = fmaf(fmaxf(nj.w + pe3, 0.f), sw1[k4 + 3], aj);
      }
      // softmax over ALL N nodes of (a * H): non-members enter with logit 0 (:135-137)
      float si = (ai + b1) * hs, sj = (aj + b1) * hs;
      float wi_, wj_;
      if (i == j) {
        float mx = (N > 1) ? fmaxf(si, 0.f) : si;
        float ei = expf(si - mx);
        float den = ei + static_cast<float>(N - 1) * expf(-mx);
        wi_ = ei / den * 2.f; wj_ = 0.f;
      } else {
        float mx = fmaxf(si, sj);
        if (N > 2) mx = fmaxf(mx, 0.f);
        float ei = expf(si - mx), ej = expf(sj - mx);
        float den = ei + ej + static_cast<float>(N - 2) * expf(-mx);
        wi_ = ei / den; wj_ = ej / den;
      }
      wts[2 * el] = wi_; wts[2 * el + 1] = wj_;
    }
    __syncthreads();
    // phase B: edges[e][:] = w_i x'_i + w_j x'_j, one float4 per task, coalesced stores
    for (int t = tid; t < ec * 16; t += GN_THREADS) {
      const int el = t >> 4, c = t & 15;
      const int e = e0 + el, i = e / N, j = e - i * N;
      float wi_ = wts[2 * el], wj_ = wts[2 * el + 1];
      float4 xi = *reinterpret_cast<const float4*>(xs + i * N2E_LD + 4 * c);
      float4 xj = *reinterpret_cast<const float4*>(xs + j * N2E_LD + 4 * c);
      float4 o;
      o.x = fmaf(wi_, xi.x, wj_ * xj.x); o.y = fmaf(wi_, xi.y, wj_ * xj.y);
      o.z = fmaf(wi_, xi.z, wj_ * xj.z); o.w = fmaf(wi_, xi.w, wj_ * xj.w);
      *reinterpret_cast<float4*>(edges + (static_cast<size_t>(b) * E + e) * GN_ATT_DIM + 4 * c) = o;
    }
  }
}

constexpr int N2H_THREADS = 128;                        // block size of edge2node_hyper

// ===========================================================================
// k2c: node2edge, hyper, ONE WARP PER HYPEREDGE (N <= 64, D % 4 == 0, D <= 256).
// model/MS_HGNN_batch.py:357-370, with the 32 lanes of a warp across the attention
// hidden units (32) / the feature columns, so a CTA of 8 warps keeps every lane busy whatever E is:
// a thread-per-hyperedge form left the crowd shape (N = E = 64, h_dim 256) with 64 active threads per SM.  A CTA stages SC
// whole scenes (x', pq, H and — only when the caller wants eo = H @ h — h) in shared memory with
// 128-bit loads; the member list of an edge is two ballots over its incidence row.
// ===========================================================================
constexpr int N2W_THREADS = 256;

__global__ void __launch_bounds__(N2W_THREADS)
node2edge_hyper_warp_kernel(const float* __restrict__ xprime, const float* __restrict__ pq,
                            const float* __restrict__ h, const float* __restrict__ H,
                            int B, int N, int E, int D, long long hstride, int SC, gn_stage_weights W,
                            float* __restrict__ edges, float* __restrict__ eo) {
  extern __shared__ __align__(16) float smem[];
  const int maxnodes = SC * N, maxedges = SC * E;
  float* xs = smem;                               // [SC*N][64]
  float* ps = xs + maxnodes * 64;                 // [SC*N][64]  pn | q
  float* Hs = ps + maxnodes * 64;                 // [SC*E][N]
  float* hs = Hs + ((maxedges * N + 3) & ~3);     // [SC*N][D]   (only when eo != nullptr)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float b0 = __ldg(W.att_b0 + lane), w1 = __ldg(W.att_w1 + lane), b1 = __ldg(W.att_b1);
  const int ntiles = (B + SC - 1) / SC;
  const int dch = (D + 31) >> 5;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0s = tile * SC, ns = min(SC, B - b0s);
    const int nn = ns * N, ne = ns * E;
    __syncthreads();
    for (int i = tid; i < nn * 16; i += N2W_THREADS) {
      const size_t g = static_cast<size_t>(b0s) * N * 64 + 4 * static_cast<size_t>(i);
      *reinterpret_cast<float4*>(xs + 4 * i) = ldg_f4(xprime + g);
      *reinterpret_cast<float4*>(ps + 4 * i) = ldg_f4(pq + g);
    }
    if (eo != nullptr) {
      const int tot4 = nn * (D >> 2);
      const float* src = h + static_cast<size_t>(b0s) * N * D;
      for (int i = tid; i < tot4; i += N2W_THREADS)
        *reinterpret_cast<float4*>(hs + 4 * i) = ldg_f4(src + 4 * static_cast<size_t>(i));
    }
    {
      const int per = E * N;
      for (int i = tid; i < ns * per; i += N2W_THREADS) {
        const int sc = i / per, r = i - sc * per;
        Hs[i] = __ldg(H + static_cast<size_t>(b0s + sc) * hstride + r);
      }
    }
    __syncthreads();
    for (int e = warp; e < ne; e += N2W_THREADS / 32) {
      const int sc = e / E, nb = sc * N;
      const float* Hr = Hs + e * N;
      const float v0 = lane < N ? Hr[lane] : 0.f;
      const float v1 = lane + 32 < N ? Hr[lane + 32] : 0.f;
      const unsigned m0 = __ballot_sync(0xffffffffu, v0 != 0.f), m1 = __ballot_sync(0xffffffffu, v1 != 0.f);
      const int cnt = __popc(m0) + __popc(m1);
      // pe[k] = b0[k] + sum_m H[e,m] q_m[k]      (lane = k)
      float pe = b0;
      for (unsigned mm = m0; mm; mm &= mm - 1) {
        const int n = __ffs(mm) - 1;
        pe = fmaf(__shfl_sync(0xffffffffu, v0, n), ps[(nb + n) * 64 + 32 + lane], pe);
      }
      for (unsigned mm = m1; mm; mm &= mm - 1) {
        const int n = __ffs(mm) - 1;
        pe = fmaf(__shfl_sync(0xffffffffu, v1, n), ps[(nb + n + 32) * 64 + 32 + lane], pe);
      }
      // a_m = (w1 . relu(pn_m + pe) + b1) * H[e,m]; member position p lives in lane p & 31, slot p >> 5
      float a_lo = 0.f, a_hi = 0.f, hv_lo = 0.f, hv_hi = 0.f;
      float mx = (cnt < N) ? 0.f : -INFINITY;
      int p = 0;
      for (unsigned mm = m0; mm; mm &= mm - 1, ++p) {
        const int n = __ffs(mm) - 1;
        const float hv = __shfl_sync(0xffffffffu, v0, n);
        const float a = (warp_sum(fmaxf(ps[(nb + n) * 64 + lane] + pe, 0.f) * w1) + b1) * hv;
        mx = fmaxf(mx, a);
        if (lane == (p & 31)) { if (p < 32) { a_lo = a; hv_lo = hv; } else { a_hi = a; hv_hi = hv; } }
      }
      for (unsigned mm = m1; mm; mm &= mm - 1, ++p) {
        const int n = __ffs(mm) - 1 + 32;
        const float hv = __shfl_sync(0xffffffffu, v1, n - 32);
        const float a = (warp_sum(fmaxf(ps[(nb + n) * 64 + lane] + pe, 0.f) * w1) + b1) * hv;
        mx = fmaxf(mx, a);
        if (lane == (p & 31)) { if (p < 32) { a_lo = a; hv_lo = hv; } else { a_hi = a; hv_hi = hv; } }
      }
      // softmax over ALL N nodes (non-members enter with logit 0, :135-137), times H
      const float ex_lo = lane < cnt ? expf(a_lo - mx) : 0.f;
      const float ex_hi = lane + 32 < cnt ? expf(a_hi - mx) : 0.f;
      const float den = warp_sum(ex_lo + ex_hi) + static_cast<float>(N - cnt) * expf(-mx);
      const float w_lo = ex_lo / den * hv_lo, w_hi = ex_hi / den * hv_hi;
      // edges_e = sum_m w_m x'_m (lane owns columns lane, lane+32);  eo_e = sum_m H[e,m] h_m
      float acc0 = 0.f, acc1 = 0.f;
      float eacc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) eacc[j] = 0.f;
      p = 0;
      for (int half = 0; half < 2; ++half) {
        for (unsigned mm = half ? m1 : m0; mm; mm &= mm - 1, ++p) {
          const int n = __ffs(mm) - 1 + 32 * half;
          const float wgt = __shfl_sync(0xffffffffu, p < 32 ? w_lo : w_hi, p & 31);
          const float* xr = xs + (nb + n) * 64;
          acc0 = fmaf(wgt, xr[lane], acc0);
          acc1 = fmaf(wgt, xr[lane + 32], acc1);
          if (eo != nullptr) {
            const float hv = __shfl_sync(0xffffffffu, half ? v1 : v0, n & 31);
            const float* hr = hs + (nb + n) * D;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < dch && j * 32 + lane < D) eacc[j] = fmaf(hv, hr[j * 32 + lane], eacc[j]);
          }
        }
      }
      float* dst = edges + (static_cast<size_t>(b0s) * E + e) * GN_ATT_DIM;
      dst[lane] = acc0;
      dst[lane + 32] = acc1;
      if (eo != nullptr) {
        float* de = eo + (static_cast<size_t>(b0s) * E + e) * D;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < dch && j * 32 + lane < D) de[j * 32 + lane] = eacc[j];
      }
    }
  }
}

// ===========================================================================
// k2d: node2edge, hyper, FOUR LANES PER HYPEREDGE (E >= 4, N <= 64).  Same math as k2c.
// Lane q of a quad owns attention hidden units [8q, 8q+8) and feature columns [16q, 16q+16) (and a
// quarter of the h columns when eo = H @ h is wanted), so the per-member work is 128-bit shared-memory
// loads + FMAs with one 2-step quad reduction for the logit: ~7x fewer warp instructions per edge than
// the warp-per-edge form, and a 256-thread CTA is fully busy at E = 64 (crowd) and at 5 NBA scenes.
// The softmax over all N nodes (:135-137) is accumulated online (running max, rescale on change), so
// the member list — a 64-bit mask from the incidence row — is walked only twice.
// ===========================================================================
template <int EOC>   // float4 chunks of eo per lane = D / 16 (0: eo not wanted)
__global__ void __launch_bounds__(256, EOC <= 4 ? 3 : 2)
node2edge_hyper_quad_kernel(const float* __restrict__ xprime, const float* __restrict__ pq,
                            const float* __restrict__ h, const float* __restrict__ H,
                            int B, int N, int E, long long hstride, int SC, gn_stage_weights W,
                            float* __restrict__ edges, float* __restrict__ eo) {
  extern __shared__ __align__(16) float smem[];
  constexpr int D = EOC * 16, LDH = D + 4;
  const int maxnodes = SC * N, maxedges = SC * E;
  float* xs = smem;                               // [SC*N][68]
  float* ps = xs + maxnodes * N2E_LD;             // [SC*N][68]  pn | q
  float* Hs = ps + maxnodes * N2E_LD;             // [SC*E][N]
  float* hs = Hs + ((maxedges * N + 3) & ~3);     // [SC*N][D+4]
  const int tid = threadIdx.x, q = tid & 3, slot = tid >> 2;
  const unsigned qmask = 0xFu << (tid & 28);      // the quad's lanes: member loops diverge between quads
  float b0[8], w1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { b0[k] = __ldg(W.att_b0 + q * 8 + k); w1[k] = __ldg(W.att_w1 + q * 8 + k); }
  const float b1 = __ldg(W.att_b1);
  const int ntiles = (B + SC - 1) / SC;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0s = tile * SC, ns = min(SC, B - b0s);
    const int nn = ns * N, ne = ns * E;
    __syncthreads();
    for (int i = tid; i < nn * 16; i += 256) {
      const int n = i >> 4, c = i & 15;
      const size_t g = (static_cast<size_t>(b0s) * N + n) * 64 + 4 * c;
      *reinterpret_cast<float4*>(xs + n * N2E_LD + 4 * c) = ldg_f4(xprime + g);
      *reinterpret_cast<float4*>(ps + n * N2E_LD + 4 * c) = ldg_f4(pq + g);
    }
    if (EOC > 0) {
      constexpr int d4 = D / 4 > 0 ? D / 4 : 1;
      const float* src = h + static_cast<size_t>(b0s) * N * D;
      for (int i = tid; i < nn * d4; i += 256) {
        const int n = i / d4, c = i - n * d4;
        *reinterpret_cast<float4*>(hs + n * LDH + 4 * c) = ldg_f4(src + static_cast<size_t>(n) * D + 4 * c);
      }
    }
    {
      const int per = E * N;
      if (((per | static_cast<int>(hstride & 3)) & 3) == 0 && (reinterpret_cast<uintptr_t>(H) & 15) == 0) {
        const int per4 = per >> 2;
        for (int i = tid; i < ns * per4; i += 256) {
          const int sc = i / per4, r = i - sc * per4;
          *reinterpret_cast<float4*>(Hs + 4 * i) = ldg_f4(H + static_cast<size_t>(b0s + sc) * hstride + 4 * r);
        }
      } else {
        for (int i = tid; i < ns * per; i += 256) {
          const int sc = i / per, r = i - sc * per;
          Hs[i] = __ldg(H + static_cast<size_t>(b0s + sc) * hstride + r);
        }
      }
    }
    __syncthreads();
    for (int e = slot; e < ne; e += 64) {
      const bool live = true;
      const int ee = e;
      const int sc = ee / E, nb = sc * N;
      const float* Hr = Hs + ee * N;
      // membership mask: lane q tests agents q, q+4, ... (ceil(N/4) branch-free steps; the [16q, 16q+16) split
      // this replaces spent 24 % of the kernel's samples in 16 predicated iterations, 3 lanes of 4 idle at N = 11)
      unsigned lo = 0u, hi = 0u;
      for (int n = q; n < N; n += 4) {
        const unsigned bit = (live && Hr[n] != 0.f) ? 1u : 0u;
        lo |= n < 32 ? bit << (n & 31) : 0u;
        hi |= n < 32 ? 0u : bit << (n & 31);
      }
      lo |= __shfl_xor_sync(qmask, lo, 1); hi |= __shfl_xor_sync(qmask, hi, 1);
      lo |= __shfl_xor_sync(qmask, lo, 2); hi |= __shfl_xor_sync(qmask, hi, 2);
      const int cnt = __popc(lo) + __popc(hi);
      // pe[k] = b0[k] + sum_m H[e,m] q_m[k]
      float pe[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) pe[k] = b0[k];
      for (int half = 0; half < 2; ++half)
        for (unsigned mm = half ? hi : lo; mm; mm &= mm - 1) {
          const int n = __ffs(mm) - 1 + 32 * half;
          const float hv = Hr[n];
          const float* qr = ps + (nb + n) * N2E_LD + 32 + q * 8;
          const float4 u = *reinterpret_cast<const float4*>(qr), v = *reinterpret_cast<const float4*>(qr + 4);
          pe[0] = fmaf(hv, u.x, pe[0]); pe[1] = fmaf(hv, u.y, pe[1]); pe[2] = fmaf(hv, u.z, pe[2]); pe[3] = fmaf(hv, u.w, pe[3]);
          pe[4] = fmaf(hv, v.x, pe[4]); pe[5] = fmaf(hv, v.y, pe[5]); pe[6] = fmaf(hv, v.z, pe[6]); pe[7] = fmaf(hv, v.w, pe[7]);
        }
      // logits + online softmax over ALL N nodes (non-members: logit 0) + weighted gather
      float mx = (cnt < N) ? 0.f : -INFINITY;
      float den = static_cast<float>(N - cnt);
      float acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = 0.f;
      float eacc[EOC > 0 ? EOC * 4 : 1];
#pragma unroll
      for (int c = 0; c < (EOC > 0 ? EOC * 4 : 1); ++c) eacc[c] = 0.f;
      for (int half = 0; half < 2; ++half)
        for (unsigned mm = half ? hi : lo; mm; mm &= mm - 1) {
          const int n = __ffs(mm) - 1 + 32 * half;
          const float hv = Hr[n];
          const float* pr = ps + (nb + n) * N2E_LD + q * 8;
          const float4 u = *reinterpret_cast<const float4*>(pr), v = *reinterpret_cast<const float4*>(pr + 4);
          float part = fmaxf(u.x + pe[0], 0.f) * w1[0];
          part = fmaf(fmaxf(u.y + pe[1], 0.f), w1[1], part); part = fmaf(fmaxf(u.z + pe[2], 0.f), w1[2], part);
          part = fmaf(fmaxf(u.w + pe[3], 0.f), w1[3], part); part = fmaf(fmaxf(v.x + pe[4], 0.f), w1[4], part);
          part = fmaf(fmaxf(v.y + pe[5], 0.f), w1[5], part); part = fmaf(fmaxf(v.z + pe[6], 0.f), w1[6], part);
          part = fmaf(fmaxf(v.w + pe[7], 0.f), w1[7], part);
          part += __shfl_xor_sync(qmask, part, 1);
          part += __shfl_xor_sync(qmask, part, 2);
          const float a = (part + b1) * hv;
          if (a > mx) {
            const float s = expf(mx - a);
            den *= s;
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] *= s;
            mx = a;
          }
          const float ex = expf(a - mx);
          den += ex;
          const float wgt = ex * hv;
          const float* xr = xs + (nb + n) * N2E_LD + q * 16;
#pragma unroll
          for (int c = 0; c < 16; c += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(xr + c);
            acc[c] = fmaf(wgt, xv.x, acc[c]); acc[c + 1] = fmaf(wgt, xv.y, acc[c + 1]);
            acc[c + 2] = fmaf(wgt, xv.z, acc[c + 2]); acc[c + 3] = fmaf(wgt, xv.w, acc[c + 3]);
          }
          if (EOC > 0) {
            const float* hr = hs + (nb + n) * LDH + q * (EOC * 4);
#pragma unroll
            for (int c = 0; c < EOC * 4; c += 4) {
              const float4 hv4 = *reinterpret_cast<const float4*>(hr + c);
              eacc[c] = fmaf(hv, hv4.x, eacc[c]); eacc[c + 1] = fmaf(hv, hv4.y, eacc[c + 1]);
              eacc[c + 2] = fmaf(hv, hv4.z, eacc[c + 2]); eacc[c + 3] = fmaf(hv, hv4.w, eacc[c + 3]);
            }
          }
        }
      if (live) {
        const float inv = 1.f / den;
        float* dst = edges + (static_cast<size_t>(b0s) * E + e) * GN_ATT_DIM + q * 16;
#pragma unroll
        for (int c = 0; c < 16; c += 4)
          *reinterpret_cast<float4*>(dst + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
        if (EOC > 0) {
          float* de = eo + (static_cast<size_t>(b0s) * E + e) * D + q * (EOC * 4);
#pragma unroll
          for (int c = 0; c < EOC * 4; c += 4)
            *reinterpret_cast<float4*>(de + c) = make_float4(eacc[c], eacc[c + 1], eacc[c + 2], eacc[c + 3]);
        }
      }
    }
  }
}

// ===========================================================================
// k3: edge_mlp — MLP_dict_softmax (:31-53) + Gumbel softmax (:446-520)
// rows = B*E edge rows, row independent.
// smem: eT [64][LD] (reused for z) | bufT [128][LD] | wp [2*KC*128] |
//       w1s [256][16] | part [TM][17] | io [TM*16]
// ===========================================================================
template <int TM>
__global__ void __launch_bounds__(GN_THREADS)
edge_mlp_kernel(const float* __restrict__ edges, long long R, int T, int E,
                gn_stage_weights W, const float* __restrict__ U, int noise_mode,
                unsigned long long seed, long long scene_offset, int stage_index,
                float* __restrict__ dist_out, float* __restrict__ edge_feat) {
  constexpr int LD = TM + 4, RM = TM / 16;
  constexpr int NPART = GN_THREADS / TM;            // threads per row in the small layer
  extern __shared__ __align__(16) float smem[];
  float* eT = smem;
  float* bufT = eT + 64 * LD;
  float* wp = bufT + 128 * LD;
  float* w1s = wp + 2 * KC * 128;
  float* part = w1s + 256 * GN_SMALL_OUT;
  float* io = part + TM * 17;
  const int tid = threadIdx.x;
  if (noise_mode == GN_NOISE_PHILOX_DEVICE_SEED) seed = __ldg(reinterpret_cast<const unsigned long long*>(U));
  for (int i = tid; i < 256 * GN_SMALL_OUT; i += GN_THREADS) w1s[i] = __ldg(W.df_w1 + i);
  const long long ntiles = (R + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    const int nrows = static_cast<int>(min(static_cast<long long>(TM), R - row0));
    load_tile_kmajor<TM>(eT, edges + static_cast<size_t>(row0) * GN_ATT_DIM, GN_ATT_DIM, nrows,
                         GN_ATT_DIM, GN_ATT_DIM, 1.f, false);
    if (noise_mode == GN_NOISE_GIVEN) {
      const float* us = U + static_cast<size_t>(row0) * T;
      for (int i = tid; i < nrows * T; i += GN_THREADS) io[i] = __ldg(us + i);
    }
    // init_MLP: 64 -> 128 -> 64
    {
      float acc[RM][8];
      acc_zero(acc);
      gemm_accum<TM, 128>(acc, eT, W.init_w0t, 128, 0, 64, wp);
      acc_store_kmajor<TM, 128>(acc, bufT, 0, [&](int col, int, float v) {
        return fmaxf(v + __ldg(W.init_b0 + col), 0.f);
      });
    }
    {
      float acc[RM][4];
      acc_zero(acc);
      gemm_accum<TM, 64>(acc, bufT, W.init_w1t, 64, 0, 128, wp);
      acc_store_kmajor<TM, 64>(acc, eT, 0, [&](int col, int, float v) {
        return v + __ldg(W.init_b1 + col);
      });
    }
    // [MLP_distribution | MLP_factor] first layers (64 -> 256) in two halves, each
    // followed by its slice of the 256 -> 16 second layer
    float lg[GN_SMALL_OUT];
#pragma unroll
    for (int o = 0; o < GN_SMALL_OUT; ++o) lg[o] = 0.f;
    const int prow = tid % TM, ppart = tid / TM;
    for (int half = 0; half < 2; ++half) {
      float acc[RM][8];
      acc_zero(acc);
      gemm_accum<TM, 128>(acc, eT, W.df_w0t, 256, half * 128, 64, wp);
      const float* b0 = W.df_b0 + half * 128;
      acc_store_kmajor<TM, 128>(acc, bufT, 0, [&](int col, int, float v) {
        return fmaxf(v + __ldg(b0 + col), 0.f);
      });
      __syncthreads();
      constexpr int KPP = 128 / NPART;
      const float* wsl = w1s + (half * 128 + ppart * KPP) * GN_SMALL_OUT;
      const float* asl = bufT + (ppart * KPP) * LD + prow;
#pragma unroll 4
      for (int k = 0; k < KPP; ++k) {
        float a = asl[k * LD];
        const float4* wr = reinterpret_cast<const float4*>(wsl + k * GN_SMALL_OUT);
        float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
        lg[0] = fmaf(a, w0.x, lg[0]); lg[1] = fmaf(a, w0.y, lg[1]);
        lg[2] = fmaf(a, w0.z, lg[2]); lg[3] = fmaf(a, w0.w, lg[3]);
        lg[4] = fmaf(a, w1.x, lg[4]); lg[5] = fmaf(a, w1.y, lg[5]);
        lg[6] = fmaf(a, w1.z, lg[6]); lg[7] = fmaf(a, w1.w, lg[7]);
        lg[8] = fmaf(a, w2.x, lg[8]); lg[9] = fmaf(a, w2.y, lg[9]);
        lg[10] = fmaf(a, w2.z, lg[10]); lg[11] = fmaf(a, w2.w, lg[11]);
        lg[12] = fmaf(a, w3.x, lg[12]); lg[13] = fmaf(a, w3.y, lg[13]);
        lg[14] = fmaf(a, w3.z, lg[14]); lg[15] = fmaf(a, w3.w, lg[15]);
      }
      __syncthreads();                    // bufT is overwritten by the next half
    }
    // combine the NPART partial sums of each row (parts live in different warps)
    for (int pp = 1; pp < NPART; ++pp) {
      if (ppart == pp) {
#pragma unroll
        for (int o = 0; o < GN_SMALL_OUT; ++o) part[prow * 17 + o] = lg[o];
      }
      __syncthreads();
      if (ppart == 0) {
#pragma unroll
        for (int o = 0; o < GN_SMALL_OUT; ++o) lg[o] += part[prow * 17 + o];
      }
      __syncthreads();
    }
    if (ppart == 0 && prow < nrows) {
      // y = (logits + g) / tau, tau = 1/2; dist = softmax(y); factor = sigmoid(.)
      const long long grow = row0 + prow;
      float y[GN_SMALL_OUT - 1];
      float mx = -INFINITY;
#pragma unroll
      for (int t = 0; t < GN_SMALL_OUT - 1; ++t) {
        if (t < T) {
          float u;
          if (noise_mode == GN_NOISE_GIVEN) {
            u = io[prow * T + t];
          } else {
            unsigned long long el =
                (static_cast<unsigned long long>(scene_offset) * E + grow) * T + t;
            u = Philox::uniform(el, static_cast<uint32_t>(stage_index), seed);
          }
          y[t] = (lg[t] + __ldg(W.df_b1 + t) + gumbel_from_uniform(u)) / 0.5f;
          mx = fmaxf(mx, y[t]);
        }
      }
      float den = 0.f;
#pragma unroll
      for (int t = 0; t < GN_SMALL_OUT - 1; ++t)
        if (t < T) { y[t] = expf(y[t] - mx); den += y[t]; }
      float fl = 0.f;
#pragma unroll
      for (int o = 0; o < GN_SMALL_OUT; ++o)
        if (o == T) fl = lg[o] + __ldg(W.df_b1 + o);
      float factor = 1.f / (1.f + expf(-fl));
#pragma unroll
      for (int t = 0; t < GN_SMALL_OUT - 1; ++t)
        if (t < T) {
          float d = y[t] / den;
          io[prow * T + t] = d;                      // dist (staged for a coalesced copy)
          part[prow * 17 + t] = factor * d;          // edge_feat
        }
    }
    __syncthreads();
    {
      float* ef = edge_feat + static_cast<size_t>(row0) * T;
      for (int i = tid; i < nrows * T; i += GN_THREADS) {
        int r = i / T, t = i - r * T;
        ef[i] = part[r * 17 + t];
      }
      if (dist_out != nullptr) {
        float* dd = dist_out + static_cast<size_t>(row0) * T;
        for (int i = tid; i < nrows * T; i += GN_THREADS) dd[i] = io[i];
      }
    }
    __syncthreads();
  }
}

// ===========================================================================
// k4: edge_agg (hyper, as written :259-265): rows = B*E,
//   ef = sum_t edge_feat[:,t] * (W1_t relu(W0_t eo + b0_t) + b1_t)
// The per-row scale is applied to the hidden activations, so the T second
// Linears accumulate into ONE output tile (K = T*128).
// smem: eoT [Dp][LD] | hidT [128][LD] | wp | sc [TM][16]
// ===========================================================================
template <int TM, int TN, int NCH>
__global__ void __launch_bounds__(GN_THREADS)
edge_agg_kernel(const float* __restrict__ eo, const float* __restrict__ edge_feat, long long R,
                int D, int Dp, int Dc, int T, gn_stage_weights W, float* __restrict__ ef) {
  constexpr int LD = TM + 4, RM = TM / 16, RN = TN / 16;
  extern __shared__ __align__(16) float smem[];
  float* eoT = smem;
  float* hidT = eoT + Dp * LD;
  float* wp = hidT + 128 * LD;
  float* sc = wp + 2 * KC * 128;
  const int tid = threadIdx.x;
  const long long ntiles = (R + TM - 1) / TM;
  const int ldw0 = T * GN_MLP_HIDDEN;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    const int nrows = static_cast<int>(min(static_cast<long long>(TM), R - row0));
    load_tile_kmajor<TM>(eoT, eo + static_cast<size_t>(row0) * D, D, nrows, D, Dp, 1.f, false);
    for (int i = tid; i < TM * 16; i += GN_THREADS) {
      int r = i >> 4, t = i & 15;
      sc[i] = (r < nrows && t < T) ? __ldg(edge_feat + static_cast<size_t>(row0 + r) * T + t) : 0.f;
    }
    float acc2[NCH][RM][RN];
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc_zero(acc2[c]);
    for (int t = 0; t < T; ++t) {
      float acc1[RM][8];
      acc_zero(acc1);
      gemm_accum<TM, 128>(acc1, eoT, W.agg_w0t, ldw0, t * 128, Dp, wp);
      const float* b0 = W.agg_b0 + t * 128;
      acc_store_kmajor<TM, 128>(acc1, hidT, 0, [&](int col, int row, float v) {
        return fmaxf(v + __ldg(b0 + col), 0.f) * sc[row * 16 + t];
      });
      const float* w1 = W.agg_w1t + static_cast<size_t>(t) * 128 * Dc;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (c * TN < Dc) gemm_accum<TM, TN>(acc2[c], hidT, w1, Dc, c * TN, 128, wp);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      acc_foreach<TM, TN>(acc2[c], [&](int r, int cc, float& v) {
        int col = c * TN + cc;
        if (r < nrows && col < D) {
          float bsum = 0.f;
          for (int t = 0; t < T; ++t) bsum = fmaf(sc[r * 16 + t], __ldg(W.agg_b1 + t * D + col), bsum);
          ef[static_cast<size_t>(row0 + r) * D + col] = v + bsum;
        }
      });
    }
    __syncthreads();
  }
}

// ===========================================================================
// k5a: edge2node, pairwise collapse.  Work item = scene.
//   wsym[n][j] = ef[(n,j),t] + ef[(j,n),t]          (= sum_e H[e,n] ef[e,t] split by partner j)
//   G[n][t][c] = sum_j wsym[n][j] * relu(P[n][t][c] + P[j][t][c] + b0[t][c])
//   S[n][t]    = sum_j wsym[n][j]
// smem: efs [N*N*T] | Pt [N][128] | wsym [N][N+1]
// ===========================================================================
__global__ void __launch_bounds__(GN_THREADS)
edge2node_pair_kernel(const float* __restrict__ P, const float* __restrict__ edge_feat,
                      int B, int N, int T, gn_stage_weights W,
                      float* __restrict__ G, float* __restrict__ S) {
  extern __shared__ __align__(16) float smem[];
  const int E = N * N, ldn = N + 1, ldp = T * GN_MLP_HIDDEN;
  float* efs = smem;
  float* Pt = efs + ((E * T + 3) & ~3);
  float* wsym = Pt + N * GN_MLP_HIDDEN;
  const int tid = threadIdx.x, c = tid & 127, ng = tid >> 7;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    const float* efg = edge_feat + static_cast<size_t>(b) * E * T;
    for (int i = tid; i < E * T; i += GN_THREADS) efs[i] = __ldg(efg + i);
    for (int t = 0; t < T; ++t) {
      __syncthreads();
      for (int i = tid; i < N * 32; i += GN_THREADS) {
        int n = i >> 5, c4 = i & 31;
        *reinterpret_cast<float4*>(Pt + n * GN_MLP_HIDDEN + 4 * c4) =
            ldg_f4(P + (static_cast<size_t>(b) * N + n) * ldp + t * GN_MLP_HIDDEN + 4 * c4);
      }
      for (int i = tid; i < E; i += GN_THREADS) {
        int n = i / N, j = i - n * N;
        wsym[n * ldn + j] = efs[(n * N + j) * T + t] + efs[(j * N + n) * T + t];
      }
      __syncthreads();
      const float bb = __ldg(W.agg_b0 + t * GN_MLP_HIDDEN + c);
      for (int n = ng; n < N; n += 2) {
        const float pn = Pt[n * GN_MLP_HIDDEN + c] + bb;
        float g = 0.f;
        for (int j = 0; j < N; ++j)
          g = fmaf(wsym[n * ldn + j], fmaxf(pn + Pt[j * GN_MLP_HIDDEN + c], 0.f), g);
        G[(static_cast<size_t>(b) * N + n) * ldp + t * GN_MLP_HIDDEN + c] = g;
      }
      if (tid < N) {
        float s = 0.f;
        for (int j = 0; j < N; ++j) s += wsym[tid * ldn + j];
        S[(static_cast<size_t>(b) * N + tid) * 16 + t] = s;
      }
    }
  }
}

// ===========================================================================
// k5b: edge2node, hyper: agg[n][:] = sum_e H[e,n] ef[e][:]   (:267)
// A CTA of 128 threads owns SC whole scenes at a time; their ef rows and
// incidence rows (both contiguous in HBM) are staged in shared memory; one
// thread per (node, 4-column group) accumulates over the scene's edges.
// ===========================================================================
__global__ void __launch_bounds__(N2H_THREADS)
edge2node_hyper_kernel(const float* __restrict__ ef, const float* __restrict__ H,
                       int B, int N, int E, int D, long long hstride, int SC, float* __restrict__ agg) {
  extern __shared__ __align__(16) float smem[];
  const int ldn = N + 1, lde = D + 4, d4 = D >> 2;
  float* es = smem;                                   // [SC*E][D+4]
  float* Hs = es + SC * E * lde;                      // [SC*E][N+1]
  const int tid = threadIdx.x;
  const int ntiles = (B + SC - 1) / SC;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0s = tile * SC, ns = min(SC, B - b0s);
    const int ne = ns * E, nn = ns * N;
    __syncthreads();
    for (int i = tid; i < ne * d4; i += N2H_THREADS) {
      int e = i / d4, c = i - e * d4;
      *reinterpret_cast<float4*>(es + e * lde + 4 * c) =
          ldg_f4(ef + (static_cast<size_t>(b0s) * E + e) * D + 4 * c);
    }
    for (int i = tid; i < ne * N; i += N2H_THREADS) {
      int e = i / N, n = i - e * N, sc = e / E;
      Hs[e * ldn + n] = __ldg(H + static_cast<size_t>(b0s + sc) * hstride + static_cast<size_t>(e - sc * E) * N + n);
    }
    __syncthreads();
    for (int i = tid; i < nn * d4; i += N2H_THREADS) {
      int nr = i / d4, c4 = i - nr * d4;              // node row in tile, column group
      int sc = nr / N, n = nr - sc * N;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e = 0; e < E; ++e) {
        float w = Hs[(sc * E + e) * ldn + n];
        if (w != 0.f) {
          float4 v = *reinterpret_cast<const float4*>(es + (sc * E + e) * lde + 4 * c4);
          o.x = fmaf(w, v.x, o.x); o.y = fmaf(w, v.y, o.y);
          o.z = fmaf(w, v.z, o.z); o.w = fmaf(w, v.w, o.w);
        }
      }
      *reinterpret_cast<float4*>(agg + (static_cast<size_t>(b0s) * N + nr) * D + 4 * c4) = o;
    }
  }
}

// ===========================================================================
// k6: node_post — incoming = cat(agg, h) / N (:267,:120,:355); closing MLP
// 2D -> 128 -> Dout (:195).  rows = B*N.
// pairwise: agg = G W1cat^T + S b1 (second half of the collapse), K = T*128
// streamed from HBM in 64-wide chunks.
// smem: incT [K2p][LD] | hidT [128][LD] | stg [64][LD] | wp | Ss [TM][16]
// ===========================================================================
template <int TM, int TN, int NCH>
__global__ void __launch_bounds__(GN_THREADS)
node_post_kernel(const float* __restrict__ h, const float* __restrict__ aggin,
                 const float* __restrict__ G, const float* __restrict__ S,
                 int R, int N, int D, int Dc, int K2p, int T, int Dout, int Doutc, int pairwise,
                 gn_stage_weights W, float* __restrict__ out, int ld_out, float* __restrict__ agg_out) {
  constexpr int LD = TM + 4, RM = TM / 16, RN = TN / 16;
  extern __shared__ __align__(16) float smem[];
  float* incT = smem;
  float* hidT = incT + K2p * LD;
  float* stg = hidT + 128 * LD;
  float* wp = stg + 64 * LD;
  float* Ss = wp + 2 * KC * 128;
  const int tid = threadIdx.x;
  const float fN = static_cast<float>(N);
  const int ntiles = (R + TM - 1) / TM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * TM, nrows = min(TM, R - row0);
    // zero the K padding rows [2D, K2p)
    for (int i = tid; i < (K2p - 2 * D) * LD; i += GN_THREADS) incT[2 * D * LD + i] = 0.f;
    // skip half: h / N
    load_tile_kmajor<TM>(incT + D * LD, h + static_cast<size_t>(row0) * D, D, nrows, D, D, fN, true);
    if (!pairwise) {
      load_tile_kmajor<TM>(incT, aggin + static_cast<size_t>(row0) * D, D, nrows, D, D, fN, true);
    } else {
      const int Kg = T * GN_MLP_HIDDEN;
      for (int i = tid; i < TM * 16; i += GN_THREADS) {
        int r = i >> 4, t = i & 15;
        Ss[i] = (r < nrows && t < T) ? __ldg(S + static_cast<size_t>(row0 + r) * 16 + t) : 0.f;
      }
      float acc[NCH][RM][RN];
#pragma unroll
      for (int c = 0; c < NCH; ++c) acc_zero(acc[c]);
      for (int k0 = 0; k0 < Kg; k0 += 64) {
        load_tile_kmajor<TM>(stg, G + static_cast<size_t>(row0) * Kg + k0, Kg, nrows, 64, 64, 1.f, false);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          if (c * TN < Dc)
            gemm_accum<TM, TN>(acc[c], stg, W.agg_w1t + static_cast<size_t>(k0) * Dc, Dc, c * TN, 64, wp);
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c * TN < Dc) {
          acc_foreach<TM, TN>(acc[c], [&](int r, int cc, float& v) {
            int col = c * TN + cc;
            if (col < D) {
              float bsum = 0.f;
              for (int t = 0; t < T; ++t) bsum = fmaf(Ss[r * 16 + t], __ldg(W.agg_b1 + t * D + col), bsum);
              incT[col * LD + r] = __fdiv_rn(v + bsum, fN);
              if (agg_out != nullptr && r < nrows) agg_out[static_cast<size_t>(row0 + r) * D + col] = v + bsum;
            }
          });
        }
      }
    }
    // closing MLP
    {
      float acc[RM][8];
      acc_zero(acc);
      gemm_accum<TM, 128>(acc, incT, W.post_w0t, 128, 0, K2p, wp);
      acc_store_kmajor<TM, 128>(acc, hidT, 0, [&](int col, int, float v) {
        return fmaxf(v + __ldg(W.post_b0 + col), 0.f);
      });
    }
    for (int c0 = 0; c0 < Doutc; c0 += 64) {
      float acc[RM][4];
      acc_zero(acc);
      gemm_accum<TM, 64>(acc, hidT, W.post_w1t, Doutc, c0, 128, wp);
      acc_foreach<TM, 64>(acc, [&](int r, int cc, float& v) {
        int col = c0 + cc;
        if (r < nrows && col < Dout)
          out[static_cast<size_t>(row0 + r) * ld_out + col] = v + __ldg(W.post_b1 + col);
      });
    }
    __syncthreads();
  }
}

}  // namespace gn

// ===========================================================================
// host side: one stage = a fixed sequence of launches on the caller's stream
// ===========================================================================
#include <cstdlib>
#include "gn_stage.h"

namespace gn {

#define GN_TRY(expr) do { int rc__ = (expr); if (rc__ != GN_OK) return rc__; } while (0)

struct StagePlan {
  int Dp, Dc, K2p, Doutc;
  bool tc_nodes;          // node-level / aggregation GEMMs on tcgen05 (bf16 path, D % 16 == 0, Dout % 16 == 0)
  size_t off_xprime, off_pq, off_P, off_edges, off_eo, off_efeat, off_ef, off_G, off_S, off_agg, off_Y;
  size_t off_hid, off_hid2, off_hidden;   // bf16 scratch of the tensor-core path
  size_t total;
};

static int make_plan(const gn_stage_cfg* c, StagePlan& p) {
  if (c->B < 0 || c->N < 1 || c->N > GN_MAX_AGENTS) return GN_E_SHAPE;
  if (c->D < 4 || (c->D & 3) || c->D > 256) return GN_E_SHAPE;
  if (c->Dout < 1 || c->T < 1 || c->T > GN_SMALL_OUT - 1) return GN_E_SHAPE;
  if (c->pairwise) { if (c->E != c->N * c->N) return GN_E_SHAPE; }
  else if (c->E < 1 || c->E > GN_MAX_AGENTS) return GN_E_SHAPE;
  if (c->precision != GN_FP32 && c->precision != GN_BF16_TC && c->precision != GN_TF32X3) return GN_E_PRECISION;
  p.Dp = round_up(c->D, 16);
  p.Dc = c->D <= 64 ? 64 : round_up(c->D, 128);   // multiple of the TN the agg_w1t GEMMs use
  p.K2p = round_up(2 * c->D, 16);
  p.Doutc = round_up(c->Dout, 64);
  p.tc_nodes = c->precision == GN_BF16_TC && (c->D % 16 == 0) && (c->Dout % 16 == 0) && c->Dout <= 256;
  const size_t R = static_cast<size_t>(c->B) * c->N, RE = static_cast<size_t>(c->B) * c->E;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += round_up_sz(bytes, 256); return at; };
  p.off_xprime = take(R * 64 * 4);
  p.off_pq = take(R * 64 * 4);
  p.off_edges = take(RE * 64 * 4);
  p.off_efeat = take(RE * c->T * 4);
  p.off_P = p.off_G = p.off_S = p.off_eo = p.off_ef = p.off_agg = p.off_Y = 0;
  if (c->pairwise && c->precision == GN_TF32X3) p.off_Y = take(R * 128 * 4);   // Y = x' W_init0^T per node
  p.off_hid = p.off_hid2 = p.off_hidden = 0;
  if (c->pairwise) {
    p.off_P = take(R * c->T * 128 * 4);
    p.off_G = take(R * c->T * 128 * 4);
    p.off_S = take(R * 16 * 4);
    p.off_agg = take(R * c->D * 4);
  } else if (p.tc_nodes && hyper_fused_fits(c->N, c->E, c->D, c->T)) {
    p.off_agg = take(R * c->D * 4);                  // eo, hidden and ef stay on chip
  } else {
    p.off_eo = take(RE * c->D * 4);
    p.off_ef = take(RE * c->D * 4);
    p.off_agg = take(R * c->D * 4);
    if (p.tc_nodes && !hyper_agg_fits(c->D, c->T)) p.off_hidden = take(RE * c->T * 128 * 2);
  }
  if (p.tc_nodes) {
    p.off_hid = take(R * 256 * 2);
    p.off_hid2 = take(R * 128 * 2);
  }
  p.total = o;
  return GN_OK;
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
  if (bytes > 227 * 1024) return GN_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(bytes));
  return e == cudaSuccess ? GN_OK : static_cast<int>(e);
}

static int grid_for(long long work, size_t smem) {
  int per_sm = smem > 113 * 1024 ? 1 : (smem > 75 * 1024 ? 2 : (smem > 56 * 1024 ? 3 : 4));
  long long g = static_cast<long long>(GN_SM_COUNT) * per_sm;
  if (work < g) g = work;
  return g < 1 ? 1 : static_cast<int>(g);
}

// ---- scene-kernel launchers (shared by the fp32 and the bf16 path) ----
static int launch_node2edge_pair(const float* xprime, const float* pq, int B, int N,
                                 const gn_stage_weights* w, float* edges, cudaStream_t st) {
  const int E = N * N, ecmax = E < N2E_EC ? E : N2E_EC;
  size_t smem = (static_cast<size_t>(2 * N) * N2E_LD + 2 * ecmax) * 4;
  GN_TRY(set_smem(node2edge_pair_kernel, smem));
  long long work = static_cast<long long>(B) * ((E + N2E_EC - 1) / N2E_EC);
  { ProfScope ps__("node2edge_pair", st);
    node2edge_pair_kernel<<<grid_for(work, smem), GN_THREADS, smem, st>>>(xprime, pq, B, N, *w, edges); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

static int launch_node2edge_hyper(const float* xprime, const float* pq, const float* h, const float* H,
                                  int B, int N, int E, int D, long long hstride, const gn_stage_weights* w,
                                  float* edges, float* eo, cudaStream_t st) {
  const int eoc = eo ? D / 16 : 0;
  if (N <= 64 && E >= 4 && E <= 64 && (!eo || D == 64 || D == 128 || D == 256)) {   // four lanes per hyperedge
    auto qbytes = [&](int sc) -> size_t {
      size_t nodes = static_cast<size_t>(sc) * N, ed = static_cast<size_t>(sc) * E;
      return (2 * nodes * N2E_LD + ((ed * N + 3) & ~size_t(3)) + (eo ? nodes * (D + 4) : 0)) * 4;
    };
    static const int cap_kb = getenv("GN_N2E_SMEM_KB") ? atoi(getenv("GN_N2E_SMEM_KB")) : 72;
    int SC = 128 / E;                                             // two 64-edge passes per tile: half the barriers per edge
    if (SC < 1) SC = 1;
    while (SC > 1 && qbytes(SC) > static_cast<size_t>(cap_kb) * 1024) --SC;
    const size_t smem = qbytes(SC);
    const int ntiles = (B + SC - 1) / SC;
    auto go = [&](auto kern) -> int {
      GN_TRY(set_smem(kern, smem));
      int occ = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem) != cudaSuccess || occ < 1) occ = 1;
      const int grid = ntiles < GN_SM_COUNT * occ ? ntiles : GN_SM_COUNT * occ;
      { ProfScope ps__("node2edge_hyper", st);
        kern<<<grid, 256, smem, st>>>(xprime, pq, h, H, B, N, E, hstride, SC, *w, edges, eo); }
      GN_LAUNCH_CHECK();
      return GN_OK;
    };
    switch (eoc) {
      case 0:  return go(node2edge_hyper_quad_kernel<0>);
      case 4:  return go(node2edge_hyper_quad_kernel<4>);
      case 8:  return go(node2edge_hyper_quad_kernel<8>);
      default: return go(node2edge_hyper_quad_kernel<16>);
    }
  }
  if (N <= 64 && E <= 64 && D <= 256 && (D & 3) == 0) {          // warp per hyperedge
    auto wbytes = [&](int sc) -> size_t {
      size_t nodes = static_cast<size_t>(sc) * N, ed = static_cast<size_t>(sc) * E;
      return (2 * nodes * 64 + ((ed * N + 3) & ~size_t(3)) + (eo ? nodes * D : 0)) * 4;
    };
    int SC = 64 / E;                                              // ~8 edges per warp per tile
    if (SC < 1) SC = 1;
    if (SC > 16) SC = 16;
    while (SC > 1 && wbytes(SC) > 72 * 1024) --SC;
    const size_t smem = wbytes(SC);
    GN_TRY(set_smem(node2edge_hyper_warp_kernel, smem));
    const int ntiles = (B + SC - 1) / SC;
    int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int grid = ntiles < GN_SM_COUNT * per_sm ? ntiles : GN_SM_COUNT * per_sm;
    { ProfScope ps__("node2edge_hyper", st);
      node2edge_hyper_warp_kernel<<<grid, N2W_THREADS, smem, st>>>(xprime, pq, h, H, B, N, E, D, hstride, SC,
                                                                   *w, edges, eo); }
    GN_LAUNCH_CHECK();
    return GN_OK;
  }
  return GN_E_SHAPE;                                     // unreachable for the shapes make_plan() admits
}

static int launch_edge2node_pair(const float* P, const float* efeat, int B, int N, int T,
                                 const gn_stage_weights* w, float* G, float* S, cudaStream_t st) {
  const int E = N * N;
  size_t smem = (((static_cast<size_t>(E) * T + 3) & ~size_t(3)) + static_cast<size_t>(N) * 128 +
                 static_cast<size_t>(N) * (N + 1)) * 4;
  GN_TRY(set_smem(edge2node_pair_kernel, smem));
  { ProfScope ps__("edge2node_pair", st);
    edge2node_pair_kernel<<<grid_for(B, smem), GN_THREADS, smem, st>>>(P, efeat, B, N, T, *w, G, S); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

static int launch_edge2node_hyper(const float* ef, const float* H, int B, int N, int E, int D,
                                  long long hstride, float* agg, cudaStream_t st) {
  auto bytes = [&](int sc) -> size_t {
    return (static_cast<size_t>(sc) * E * (D + 4) + static_cast<size_t>(sc) * E * (N + 1)) * 4;
  };
  // 24 KB tiles x 8 CTAs per SM: twice the loads in flight of the 48 KB x 4 configuration (0.260 -> 0.169 ms at the
  // NBA shape, profiles/hbm_kernel_sweep.py); the knobs stay for that sweep
  static const int cap_kb = getenv("GN_E2N_SMEM_KB") ? atoi(getenv("GN_E2N_SMEM_KB")) : 24;
  static const int per_sm_env = getenv("GN_E2N_CTAS") ? atoi(getenv("GN_E2N_CTAS")) : 8;
  int SC = N2H_THREADS / N; if (SC < 1) SC = 1; if (SC > 16) SC = 16;
  while (SC > 1 && bytes(SC) > static_cast<size_t>(cap_kb) * 1024) --SC;
  size_t smem = bytes(SC);
  GN_TRY(set_smem(edge2node_hyper_kernel, smem));
  const int ntiles = (B + SC - 1) / SC;
  int grid = ntiles < GN_SM_COUNT * per_sm_env ? ntiles : GN_SM_COUNT * per_sm_env;
  { ProfScope ps__("edge2node_hyper", st);
    edge2node_hyper_kernel<<<grid, N2H_THREADS, smem, st>>>(ef, H, B, N, E, D, hstride, SC, agg); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

static TcLinArgs lin_args() {
  TcLinArgs a;
  memset(&a, 0, sizeof(a));
  return a;
}

int stage_fwd(const gn_stage_cfg* c, const gn_stage_weights* w, const float* h,
              const float* H, const float* U, float* node_out, float* dist_out,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  StagePlan p;
  GN_TRY(make_plan(c, p));
  if (ws_bytes < p.total) return GN_E_WORKSPACE;
  if (c->B == 0) return GN_OK;
  char* base = static_cast<char*>(ws);
  auto fptr = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  float* xprime = fptr(p.off_xprime);
  float* pq = fptr(p.off_pq);
  float* edges = fptr(p.off_edges);
  float* efeat = fptr(p.off_efeat);
  float* P = fptr(p.off_P);
  float* G = fptr(p.off_G);
  float* S = fptr(p.off_S);
  float* eo = fptr(p.off_eo);
  float* ef = fptr(p.off_ef);
  float* agg = fptr(p.off_agg);
  float* Ypre = fptr(p.off_Y);
  __nv_bfloat16* hid = reinterpret_cast<__nv_bfloat16*>(base + p.off_hid);
  __nv_bfloat16* hid2 = reinterpret_cast<__nv_bfloat16*>(base + p.off_hid2);
  __nv_bfloat16* hidden = reinterpret_cast<__nv_bfloat16*>(base + p.off_hidden);
  const int B = c->B, N = c->N, D = c->D, E = c->E, T = c->T;
  const int R = B * N;
  const long long RE = static_cast<long long>(B) * E;
  const long long hstride = c->h_stride > 0 ? c->h_stride : static_cast<long long>(E) * N;
  if (hstride < static_cast<long long>(E) * N) return GN_E_SHAPE;
  const int ld_out = c->out_ld > 0 ? c->out_ld : c->Dout;
  if (ld_out < c->Dout) return GN_E_SHAPE;
  if (c->noise_mode < GN_NOISE_GIVEN || c->noise_mode > GN_NOISE_PHILOX_DEVICE_SEED) return GN_E_SHAPE;
  if (c->noise_mode != GN_NOISE_PHILOX && !U) return GN_E_NULL;
  if (c->noise_mode == GN_NOISE_PHILOX_DEVICE_SEED && (reinterpret_cast<uintptr_t>(U) & 7)) return GN_E_SHAPE;
  if (!c->pairwise && !H) return GN_E_NULL;
  const bool tcn = p.tc_nodes;
  const bool tf = c->precision == GN_TF32X3;       // fp32-grade tensor-core chains where the shape fits
  const long long RE_ = static_cast<long long>(B) * E;
  // pairwise aggregation fused into one kernel (P / G stay on chip) where the shape fits
  const bool tf_fused_agg = tf && c->pairwise && w->tf_pagg_w != nullptr && pair_agg_tf32_fits(N, D, T);
  const bool tf_pre = tf && w->tf_pre_w && node_pre_tf32_fits(D) &&
                      (!c->pairwise || tf_fused_agg || (w->tf_aggin_w && agg_in_tf32_fits(D, T)));
  // pairwise edge chain with node2edge and the first Linear of init_MLP folded into its staging pass: needs the
  // tensor-core prologue (which emits Y) and an h_dim of 64 (the prologue's Y op reads the 64-wide x' operand)
  const bool tf_pair_y = tf_pre && c->pairwise && w->tf_chain_w != nullptr && edge_chain_tf32_fits(true, N, T) &&
                         RE_ < (1LL << 31);
  const bool fused_agg = tcn && c->pairwise && pair_agg_fits(N, D, T);   // P / G stay on chip
  if (tcn && (!w->tc_node_w0 || !w->tc_node_w1 || !w->tc_att_wpq || !w->tc_agg_w0 || !w->tc_agg_w1 ||
              !w->tc_post_w0 || !w->tc_post_w1)) return GN_E_NULL;

  // ---- k1: node-level prologue: x', pq (and P for the pairwise collapse)
  const bool chain_ok = tcn && D == 64 && (c->Dout % 32 == 0) && c->Dout <= 128;
  if (chain_ok && (!c->pairwise || fused_agg)) {
    NodeChainArgs a;
    memset(&a, 0, sizeof(a));
    a.A0 = h; a.lda0 = D; a.K0 = D; a.R = R; a.nsteps = 3;
    a.step[0] = {static_cast<const __nv_bfloat16*>(w->tc_node_w0), w->node_b0, D, 256, 1, nullptr, 0, 0};
    a.step[1] = {static_cast<const __nv_bfloat16*>(w->tc_node_w1), w->node_b1, 256, 64, 0, xprime, 64, 0};
    a.step[2] = {static_cast<const __nv_bfloat16*>(w->tc_att_wpq), nullptr, 64, 64, 0, pq, 64, 0};
    GN_TRY(launch_node_chain_tc(a, "node_pre_chain_tc", st));
  } else if (tcn) {
    TcLinArgs a = lin_args();
    if (node_pre256_fits(D)) {
      GN_TRY(launch_node_pre256_tc(h, R, w, xprime, pq, st));
    } else {
    a.A0 = h; a.a0_is_f32 = 1; a.lda0 = D; a.K0 = D; a.R = R;
    a.W = static_cast<const __nv_bfloat16*>(w->tc_node_w0); a.Ntot = 256; a.N = 256;
    a.bias = w->node_b0; a.relu = 1; a.out = hid; a.out_is_f32 = 0; a.ldo = 256;
    GN_TRY(launch_tc_linear(a, "node_mlp0_tc", st));
    a = lin_args();
    a.A0 = hid; a.a0_is_f32 = 0; a.lda0 = 256; a.K0 = 256; a.R = R;
    a.W = static_cast<const __nv_bfloat16*>(w->tc_node_w1); a.Ntot = 64; a.N = 64;
    a.bias = w->node_b1; a.out = xprime; a.out_is_f32 = 1; a.ldo = 64;
    GN_TRY(launch_tc_linear(a, "node_mlp1_tc", st));
    a = lin_args();
    a.A0 = xprime; a.a0_is_f32 = 1; a.lda0 = 64; a.K0 = 64; a.R = R;
    a.W = static_cast<const __nv_bfloat16*>(w->tc_att_wpq); a.Ntot = 64; a.N = 64;
    a.out = pq; a.out_is_f32 = 1; a.ldo = 64;
    GN_TRY(launch_tc_linear(a, "att_proj_tc", st));
    }
    if (c->pairwise && !fused_agg) {
      const int NT = T * 128;
      for (int n0 = 0; n0 < NT; n0 += 256) {
        a = lin_args();
        a.A0 = h; a.a0_is_f32 = 1; a.lda0 = D; a.K0 = D; a.R = R;
        a.W = static_cast<const __nv_bfloat16*>(w->tc_agg_w0); a.Ntot = NT; a.n0 = n0;
        a.N = NT - n0 < 256 ? NT - n0 : 256;
        a.out = P; a.out_is_f32 = 1; a.ldo = NT; a.out_col0 = n0;
        GN_TRY(launch_tc_linear(a, "agg_in_tc", st));
      }
    }
  } else if (tf_pre) {
    GN_TRY(launch_node_pre_tf32(h, R, D, w, xprime, pq, tf_pair_y ? Ypre : nullptr, st));
    if (c->pairwise && !tf_fused_agg) GN_TRY(launch_agg_in_tf32(h, R, D, T, w, P, st));
  } else {
    constexpr int TM = 64, LD = TM + 4;
    size_t smem = static_cast<size_t>(p.Dp + 128 + 64) * LD * 4 + 2 * KC * 128 * 4;
    auto kern = node_pre_kernel<TM>;
    GN_TRY(set_smem(kern, smem));
    int grid = grid_for((R + TM - 1) / TM, smem);
    { ProfScope ps__("node_pre", st);
      kern<<<grid, GN_THREADS, smem, st>>>(h, R, D, p.Dp, T, c->pairwise, *w, xprime, pq, P); }
    GN_LAUNCH_CHECK();
  }

  // ---- k2: node2edge (fused into the tensor-core chain for the pairwise bf16 path)
  const bool tc = c->precision == GN_BF16_TC;
  const bool tf_chain = tf && w->tf_chain_w != nullptr;
  const bool fuse_pair = c->pairwise && ((tc && edge_chain_pair_fits(N)) || tf_pair_y);
  const bool out_aligned = (reinterpret_cast<uintptr_t>(node_out) & 15) == 0;   // fused tails store 128-bit rows
  const bool fused_hyper64 = tcn && !c->pairwise && out_aligned && hyper_fused64_fits(N, E, D, T, c->Dout, ld_out);
  const bool fused_hyper = fused_hyper64 || (tcn && !c->pairwise && hyper_fused_fits(N, E, D, T));   // eo / ef never leave the SM
  if (fused_hyper && !w->tc_hfuse_w) return GN_E_NULL;
  if (c->pairwise) { if (!fuse_pair) GN_TRY(launch_node2edge_pair(xprime, pq, B, N, w, edges, st)); }
  else GN_TRY(launch_node2edge_hyper(xprime, pq, h, H, B, N, E, D, hstride, w, edges,
                                     fused_hyper ? nullptr : eo, st));

  // ---- k3: per-edge MLP chain + Gumbel softmax
  if (tc) {
    GN_TRY(launch_edge_chain_tc(fuse_pair, edges, xprime, pq, N, E, T, RE, w, U, c->noise_mode,
                                c->seed, c->scene_offset, c->stage_index, dist_out, efeat, st));
  } else if (tf_chain) {
    GN_TRY(launch_edge_chain_tf32(fuse_pair, edges, Ypre, pq, N, E, T, RE, w, U, c->noise_mode,
                                  c->seed, c->scene_offset, c->stage_index, dist_out, efeat, st));
  } else {
    constexpr int TM = 128, LD = TM + 4;
    size_t smem = (static_cast<size_t>(64 + 128) * LD + 2 * KC * 128 + 256 * GN_SMALL_OUT +
                   TM * 17 + TM * 16) * 4;
    auto kern = edge_mlp_kernel<TM>;
    GN_TRY(set_smem(kern, smem));
    { ProfScope ps__("edge_mlp", st);
      kern<<<grid_for((RE + TM - 1) / TM, smem), GN_THREADS, smem, st>>>(
          edges, RE, T, E, *w, U, c->noise_mode, c->seed, c->scene_offset, c->stage_index,
          dist_out, efeat); }
    GN_LAUNCH_CHECK();
  }

  // ---- k4/k5: aggregation
  if (c->pairwise) {
    if (fused_agg) GN_TRY(launch_pair_agg_tc(h, efeat, B, N, T, w, agg, st));
    else if (tf_fused_agg && tf_pre) GN_TRY(launch_pair_agg_tf32(h, efeat, B, N, T, w, agg, st));
    else GN_TRY(launch_edge2node_pair(P, efeat, B, N, T, w, G, S, st));
  } else if (fused_hyper64) {
    return launch_hyper_fused64_tc(h, H, efeat, B, N, T, hstride, w, node_out, ld_out, c->Dout, st);
  } else if (fused_hyper) {
    const bool post_in = out_aligned && hyper_fused_post_fits(c->Dout, ld_out);
    GN_TRY(launch_hyper_fused_tc(h, H, efeat, B, N, T, hstride, w, agg, post_in ? node_out : nullptr,
                                 ld_out, c->Dout, st));
    if (post_in) return GN_OK;
  } else {
    if (tcn && hyper_agg_fits(D, T)) {
      GN_TRY(launch_hyper_agg_tc(eo, efeat, RE, T, w, ef, st));
    } else if (tcn) {
      const int NT = T * 128;
      for (int n0 = 0; n0 < NT; n0 += 256) {        // hidden = relu(eo W0^T + b0) * edge_feat[:, t]
        TcLinArgs a = lin_args();
        a.A0 = eo; a.a0_is_f32 = 1; a.lda0 = D; a.K0 = D; a.R = RE;
        a.W = static_cast<const __nv_bfloat16*>(w->tc_agg_w0); a.Ntot = NT; a.n0 = n0;
        a.N = NT - n0 < 256 ? NT - n0 : 256;
        a.bias = w->agg_b0; a.relu = 1; a.rowscale = efeat; a.rs_ld = T; a.rs_shift = 7;
        a.out = hidden; a.out_is_f32 = 0; a.ldo = NT; a.out_col0 = n0;
        GN_TRY(launch_tc_linear(a, "agg_hidden_tc", st));
      }
      TcLinArgs a = lin_args();                      // ef = hidden W1cat^T + sum_t edge_feat_t b1_t
      a.A0 = hidden; a.a0_is_f32 = 0; a.lda0 = NT; a.K0 = NT; a.R = RE;
      a.W = static_cast<const __nv_bfloat16*>(w->tc_agg_w1); a.Ntot = D; a.N = D;
      a.rowscale = efeat; a.rs_ld = T; a.bias_mat = w->agg_b1; a.bm_T = T; a.bm_ld = D;
      a.out = ef; a.out_is_f32 = 1; a.ldo = D;
      GN_TRY(launch_tc_linear(a, "agg_out_tc", st));
    } else if (tf && w->tf_hagg_w && hyper_agg_tf32_fits(D, T)) {
      GN_TRY(launch_hyper_agg_tf32(eo, efeat, RE, D, T, w, ef, st));
    } else if (D <= 64) {
      constexpr int TM = 128, LD = TM + 4;
      size_t smem = (static_cast<size_t>(p.Dp + 128) * LD + 2 * KC * 128 + TM * 16) * 4;
      auto kern = edge_agg_kernel<TM, 64, 1>;
      GN_TRY(set_smem(kern, smem));
      { ProfScope ps__("edge_agg", st);
        kern<<<grid_for((RE + TM - 1) / TM, smem), GN_THREADS, smem, st>>>(eo, efeat, RE, D, p.Dp, p.Dc, T, *w, ef); }
      GN_LAUNCH_CHECK();
    } else if (D <= 128) {
      constexpr int TM = 128, LD = TM + 4;
      size_t smem = (static_cast<size_t>(p.Dp + 128) * LD + 2 * KC * 128 + TM * 16) * 4;
      auto kern = edge_agg_kernel<TM, 128, 1>;
      GN_TRY(set_smem(kern, smem));
      { ProfScope ps__("edge_agg", st);
        kern<<<grid_for((RE + TM - 1) / TM, smem), GN_THREADS, smem, st>>>(eo, efeat, RE, D, p.Dp, p.Dc, T, *w, ef); }
      GN_LAUNCH_CHECK();
    } else {
      constexpr int TM = 64, LD = TM + 4;
      size_t smem = (static_cast<size_t>(p.Dp + 128) * LD + 2 * KC * 128 + TM * 16) * 4;
      auto kern = edge_agg_kernel<TM, 128, 2>;
      GN_TRY(set_smem(kern, smem));
      { ProfScope ps__("edge_agg", st);
        kern<<<grid_for((RE + TM - 1) / TM, smem), GN_THREADS, smem, st>>>(eo, efeat, RE, D, p.Dp, p.Dc, T, *w, ef); }
      GN_LAUNCH_CHECK();
    }
    GN_TRY(launch_edge2node_hyper(ef, H, B, N, E, D, hstride, agg, st));
  }

  // ---- k6: closing MLP on [agg | h] / N
  const bool pagg_done = tf_fused_agg && tf_pre;        // agg already holds G W1cat^T + S b1
  const int post_pair = (c->pairwise && !pagg_done) ? 1 : 0;
  if (tcn) {
    TcLinArgs a;
    if (c->pairwise && !fused_agg) {                 // agg = G W1cat^T + S b1  (second half of the collapse)
      const int NT = T * 128;
      a = lin_args();
      a.A0 = G; a.a0_is_f32 = 1; a.lda0 = NT; a.K0 = NT; a.R = R;
      a.W = static_cast<const __nv_bfloat16*>(w->tc_agg_w1); a.Ntot = D; a.N = D;
      a.rowscale = S; a.rs_ld = 16; a.bias_mat = w->agg_b1; a.bm_T = T; a.bm_ld = D;
      a.out = agg; a.out_is_f32 = 1; a.ldo = D;
      GN_TRY(launch_tc_linear(a, "agg_out_tc", st));
    }
    if (chain_ok) {
      NodeChainArgs n;
      memset(&n, 0, sizeof(n));
      n.A0 = agg; n.lda0 = D; n.K0 = D; n.A1 = h; n.lda1 = D; n.K1 = D; n.a_div = static_cast<float>(N);
      n.R = R; n.nsteps = 2;
      n.step[0] = {static_cast<const __nv_bfloat16*>(w->tc_post_w0), w->post_b0, 2 * D, 128, 1, nullptr, 0, 0};
      n.step[1] = {static_cast<const __nv_bfloat16*>(w->tc_post_w1), w->post_b1, 128, c->Dout, 0, node_out, ld_out, 0};
      GN_TRY(launch_node_chain_tc(n, "node_post_chain_tc", st));
      return GN_OK;
    }
    a = lin_args();
    a.A0 = agg; a.a0_is_f32 = 1; a.lda0 = D; a.K0 = D; a.A1 = h; a.lda1 = D; a.K1 = D;
    a.a_div = static_cast<float>(N); a.R = R;
    a.W = static_cast<const __nv_bfloat16*>(w->tc_post_w0); a.Ntot = 128; a.N = 128;
    a.bias = w->post_b0; a.relu = 1; a.out = hid2; a.out_is_f32 = 0; a.ldo = 128;
    GN_TRY(launch_tc_linear(a, "post_mlp0_tc", st));
    a = lin_args();
    a.A0 = hid2; a.a0_is_f32 = 0; a.lda0 = 128; a.K0 = 128; a.R = R;
    a.W = static_cast<const __nv_bfloat16*>(w->tc_post_w1); a.Ntot = c->Dout; a.N = c->Dout;
    a.bias = w->post_b1; a.out = node_out; a.out_is_f32 = 1; a.ldo = ld_out;
    GN_TRY(launch_tc_linear(a, "post_mlp1_tc", st));
  } else if (tf && w->tf_post_w && node_post_tf32_fits(D, c->Dout, ld_out, node_out) &&
             (!c->pairwise || pagg_done || (w->tf_aggout_w && agg_out_tf32_fits(D, T)))) {
    if (c->pairwise && !pagg_done) GN_TRY(launch_agg_out_tf32(G, S, R, D, T, w, agg, st));
    GN_TRY(launch_node_post_tf32(agg, h, R, D, N, c->Dout, w, node_out, ld_out, st));
  } else {
    constexpr int TM = 64, LD = TM + 4;
    size_t smem = (static_cast<size_t>(p.K2p + 128 + 64) * LD + 2 * KC * 128 + TM * 16) * 4;
    int grid = grid_for((R + TM - 1) / TM, smem);
    if (D <= 64) {
      auto kern = node_post_kernel<TM, 64, 1>;
      GN_TRY(set_smem(kern, smem));
      { ProfScope ps__("node_post", st);
        kern<<<grid, GN_THREADS, smem, st>>>(h, agg, G, S, R, N, D, p.Dc, p.K2p, T, c->Dout, p.Doutc, post_pair, *w, node_out, ld_out, post_pair ? agg : nullptr); }
    } else if (D <= 128) {
      auto kern = node_post_kernel<TM, 128, 1>;
      GN_TRY(set_smem(kern, smem));
      { ProfScope ps__("node_post", st);
        kern<<<grid, GN_THREADS, smem, st>>>(h, agg, G, S, R, N, D, p.Dc, p.K2p, T, c->Dout, p.Doutc, post_pair, *w, node_out, ld_out, post_pair ? agg : nullptr); }
    } else {
      auto kern = node_post_kernel<TM, 128, 2>;
      GN_TRY(set_smem(kern, smem));
      { ProfScope ps__("node_post", st);
        kern<<<grid, GN_THREADS, smem, st>>>(h, agg, G, S, R, N, D, p.Dc, p.K2p, T, c->Dout, p.Doutc, post_pair, *w, node_out, ld_out, post_pair ? agg : nullptr); }
    }
    GN_LAUNCH_CHECK();
  }
  return GN_OK;
}

int stage_saved_offsets(const gn_stage_cfg* c, size_t* out5) {
  StagePlan p;
  GN_TRY(make_plan(c, p));
  out5[0] = p.off_xprime; out5[1] = p.off_pq; out5[2] = p.off_edges; out5[3] = p.off_efeat; out5[4] = p.off_agg;
  return GN_OK;
}

size_t stage_workspace_bytes(const gn_stage_cfg* c) {
  StagePlan p;
  if (make_plan(c, p) != GN_OK) return 0;
  return p.total;
}

int stage_launch_count(const gn_stage_cfg* c) {
  StagePlan p;
  if (make_plan(c, p) != GN_OK) return 0;
  if (c->precision == GN_TF32X3) {
    // upper bound when every chain fits (weights present): pre (+P), chain (node2edge fused for the pairwise layer),
    // aggregation, post
    const bool fagg = c->pairwise && pair_agg_tf32_fits(c->N, c->D, c->T);
    const bool pre = node_pre_tf32_fits(c->D) && (!c->pairwise || fagg || agg_in_tf32_fits(c->D, c->T));
    const bool post = node_post_tf32_fits(c->D, c->Dout, c->out_ld > 0 ? c->out_ld : c->Dout, nullptr) &&
                      (!c->pairwise || fagg || agg_out_tf32_fits(c->D, c->T));
    const int fused_tf = (c->pairwise && edge_chain_tf32_fits(true, c->N, c->T)) ? 1 : 0;
    if (c->pairwise) return ((pre && !fagg) ? 2 : 1) + (2 - fused_tf) + 1 + ((post && !fagg) ? 2 : 1);
    return 1 + 1 + 1 + 1 + 1 + 1;
  }
  const int fused = (c->precision == GN_BF16_TC && c->pairwise && edge_chain_pair_fits(c->N)) ? 1 : 0;
  if (!p.tc_nodes) return (c->pairwise ? 5 : 6) - fused;
  const int chunks = (c->T * 128 + 255) / 256;
  const bool chain = c->D == 64 && (c->Dout % 32 == 0) && c->Dout <= 128;
  const int pre = (chain || node_pre256_fits(c->D)) ? 1 : 3, post = chain ? 1 : 2;
  if (c->pairwise && pair_agg_fits(c->N, c->D, c->T)) return pre + (2 - fused) + 1 + post;
  if (!c->pairwise && hyper_fused64_fits(c->N, c->E, c->D, c->T, c->Dout, c->out_ld > 0 ? c->out_ld : c->Dout))
    return pre + 1 + 1 + 1;
  if (!c->pairwise && hyper_fused_fits(c->N, c->E, c->D, c->T))
    return pre + 1 + 1 + 1 + (hyper_fused_post_fits(c->Dout, c->out_ld > 0 ? c->out_ld : c->Dout) ? 0 : post);
  if (!c->pairwise && hyper_agg_fits(c->D, c->T)) return pre + 1 + 1 + 1 + 1 + post;
  return c->pairwise ? (pre + chunks) + (2 - fused) + 1 + 3 : pre + 1 + 1 + (chunks + 1) + 1 + 2;
}

}  // namespace gn
