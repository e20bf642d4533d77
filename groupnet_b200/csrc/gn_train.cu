// Backward of one message-passing stage (fp32).  Hand-written CUDA; correctness-first.
//
// The forward math (model/MS_HGNN_batch.py:122-141, :41-53, :446-520, :259-268, :116-120, :220-229)
// is differentiated by hand.  The backward recomputes the cheap intermediates it needs from what the
// fp32 forward leaves in its workspace (x', pq, edges, edge_feat, agg) and runs, per stage:
//   closing MLP  ->  /N, split of the concat  ->  edge2node^T (gather to edges)  ->  T aggregation
//   MLPs as written (per edge row, eo = H @ h)  ->  edge2node/node2edge scatter of d_eo  ->
//   Gumbel-softmax / sigmoid  ->  MLP_dict_softmax MLPs  ->  attention softmax + weighted gather
//   (per edge, members from the incidence)  ->  attention projections  ->  node2edge_start_mlp.
// H and corr receive no gradient (the reference uses top-k INDICES only, :382); the Gumbel noise is
// a constant.  Dense parts use three generic SGEMM kernels that read nn.Linear parameters in their
// native (N_out, K) layout and accumulate dW / db straight into caller-provided gradient buffers
// (atomicAdd over row slices: summation order is not fixed, results agree to ~1e-6).
#include "gn_common.cuh"
#include "gn_stage.h"

namespace gn {

// ---------------------------------------------------------------------------
// SGEMM family, 64 x 64 tile, 256 threads, 4 x 4 per thread, k-step 16.
//   NT:  Y[M x N]  (op)= X[M x K] * W[N x K]^T (+ bias) (ReLU) ; flags: accumulate
//   NN:  Y[M x N]   =  X[M x K] * W[K x N]                 ; optional mask: Y *= (Ref > 0)
//   TN:  C[N x K]  +=  A[M x N]^T * B[M x K]               ; rows sliced over blockIdx.z, atomicAdd
// ---------------------------------------------------------------------------
constexpr int SG_T = 64, SG_K = 16;

template <bool WT>   // WT: W is [N x K] (use W[n][k]);  !WT: W is [K x N] (use W[k][n])
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ W, long long ldw,
             const float* __restrict__ bias, float* __restrict__ Y, long long ldy,
             long long M, int N, int K, int relu, int accumulate,
             const float* __restrict__ maskref, long long ldm) {
  __shared__ __align__(16) float xs[SG_K][SG_T + 4];
  __shared__ __align__(16) float ws[SG_K][SG_T + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long m0 = static_cast<long long>(blockIdx.x) * SG_T;
  const int n0 = blockIdx.y * SG_T;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_K) {
    for (int i = threadIdx.x; i < SG_T * SG_K; i += 256) {
      const int r = i / SG_K, k = i - r * SG_K;
      const long long m = m0 + r;
      xs[k][r] = (m < M && k0 + k < K) ? X[m * ldx + k0 + k] : 0.f;
    }
    if (WT) {
      for (int i = threadIdx.x; i < SG_T * SG_K; i += 256) {
        const int c = i / SG_K, k = i - c * SG_K;
        ws[k][c] = (n0 + c < N && k0 + k < K) ? W[static_cast<long long>(n0 + c) * ldw + k0 + k] : 0.f;
      }
    } else {
      for (int i = threadIdx.x; i < SG_T * SG_K; i += 256) {
        const int k = i / SG_T, c = i - k * SG_T;
        ws[k][c] = (n0 + c < N && k0 + k < K) ? W[static_cast<long long>(k0 + k) * ldw + n0 + c] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_K; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&ws[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long m = m0 + ty * 4 + u;
    if (m >= M) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int n = n0 + tx * 4 + v;
      if (n >= N) continue;
      float y = acc[u][v];
      if (bias != nullptr) y += bias[n];
      if (relu) y = fmaxf(y, 0.f);
      if (maskref != nullptr && !(maskref[m * ldm + n] > 0.f)) y = 0.f;
      if (accumulate) y += Y[m * ldy + n];
      Y[m * ldy + n] = y;
    }
  }
}

// C[N x K] += sum over rows m in this block's slice of A[m][n] * B[m][k];  cb[n] += sum_m A[m][n] (k-tile 0)
__global__ void __launch_bounds__(256)
sgemm_tn_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ Bm, long long ldb,
                float* __restrict__ C, long long ldc, float* __restrict__ cb,
                long long M, int N, int K, long long rows_per_slice) {
  __shared__ __align__(16) float as[SG_K][SG_T + 4];
  __shared__ __align__(16) float bs[SG_K][SG_T + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int n0 = blockIdx.x * SG_T, k0 = blockIdx.y * SG_T;
  const long long mbeg = static_cast<long long>(blockIdx.z) * rows_per_slice;
  const long long mend = min(M, mbeg + rows_per_slice);
  float acc[4][4] = {};
  float bsum[4] = {};
  for (long long m0 = mbeg; m0 < mend; m0 += SG_K) {
    for (int i = threadIdx.x; i < SG_K * SG_T; i += 256) {
      const int r = i / SG_T, c = i - r * SG_T;
      const long long m = m0 + r;
      as[r][c] = (m < mend && n0 + c < N) ? A[m * lda + n0 + c] : 0.f;
      bs[r][c] = (m < mend && k0 + c < K) ? Bm[m * ldb + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SG_K; ++r) {
      const float4 av = *reinterpret_cast<const float4*>(&as[r][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&bs[r][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (tx == 0) bsum[u] += a[u];
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int n = n0 + ty * 4 + u;
    if (n >= N) continue;
    if (cb != nullptr && blockIdx.y == 0 && tx == 0) atomicAdd(cb + n, bsum[u]);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int k = k0 + tx * 4 + v;
      if (k < K) atomicAdd(C + static_cast<long long>(n) * ldc + k, acc[u][v]);
    }
  }
}

// ---------------------------------------------------------------------------
// Fast variants for the aligned shapes of the backward (everything except the T- and 1-wide heads):
// 128 x 64 output tile, 8 x 4 per thread, k-step 16, shared memory double-buffered with a register
// prefetch of the next k-tile (one __syncthreads per step).  Same semantics as the kernels above.
// ---------------------------------------------------------------------------
template <bool WT>
__global__ void __launch_bounds__(256)
sgemm128_kernel(const float* __restrict__ X, long long ldx, const float* __restrict__ W, long long ldw,
                const float* __restrict__ bias, float* __restrict__ Y, long long ldy,
                long long M, int N, int K, int relu, int accumulate,
                const float* __restrict__ maskref, long long ldm) {
  __shared__ __align__(16) float xs[2][SG_K][128 + 4];
  __shared__ __align__(16) float ws[2][SG_K][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long m0 = static_cast<long long>(blockIdx.x) * 128;
  const int n0 = blockIdx.y * 64;
  float4 xr[2], wr;
  const int xk4 = tid & 3, xrow = tid >> 2;              // X tile: rows xrow, xrow + 64; 4 floats at k = 4 * xk4
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long m = m0 + xrow + 64 * i;
      xr[i] = m < M ? ldg_f4(X + m * ldx + k0 + 4 * xk4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (WT) {
      const int n = n0 + (tid >> 2);
      wr = n < N ? ldg_f4(W + static_cast<long long>(n) * ldw + k0 + 4 * (tid & 3)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const int k = tid >> 4, n = n0 + 4 * (tid & 15);
      const float* src = W + static_cast<long long>(k0 + k) * ldw + n;
      if (n + 3 < N) wr = ldg_f4(src);
      else wr = make_float4(n < N ? __ldg(src) : 0.f, n + 1 < N ? __ldg(src + 1) : 0.f, n + 2 < N ? __ldg(src + 2) : 0.f, 0.f);
    }
  };
  auto sstore = [&](int b) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = xrow + 64 * i;
      xs[b][4 * xk4][r] = xr[i].x; xs[b][4 * xk4 + 1][r] = xr[i].y; xs[b][4 * xk4 + 2][r] = xr[i].z; xs[b][4 * xk4 + 3][r] = xr[i].w;
    }
    if (WT) {
      const int c = tid >> 2, k4 = tid & 3;
      ws[b][4 * k4][c] = wr.x; ws[b][4 * k4 + 1][c] = wr.y; ws[b][4 * k4 + 2][c] = wr.z; ws[b][4 * k4 + 3][c] = wr.w;
    } else {
      *reinterpret_cast<float4*>(&ws[b][tid >> 4][4 * (tid & 15)]) = wr;
    }
  };
  float acc[8][4] = {};
  gload(0);
  sstore(0);
  __syncthreads();
  int b = 0;
  for (int k0 = 0; k0 < K; k0 += SG_K, b ^= 1) {
    const bool more = k0 + SG_K < K;
    if (more) gload(k0 + SG_K);
#pragma unroll
    for (int k = 0; k < SG_K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&xs[b][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&xs[b][k][ty * 8 + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&ws[b][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bw[v], acc[u][v]);
    }
    if (more) sstore(b ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const long long m = m0 + ty * 8 + u;
    if (m >= M) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int n = n0 + tx * 4 + v;
      if (n >= N) continue;
      float y = acc[u][v];
      if (bias != nullptr) y += bias[n];
      if (relu) y = fmaxf(y, 0.f);
      if (maskref != nullptr && !(maskref[m * ldm + n] > 0.f)) y = 0.f;
      if (accumulate) y += Y[m * ldy + n];
      Y[m * ldy + n] = y;
    }
  }
}

// C[N x K] += A^T B over this block's row slice; TN_ x TK_ output tile, (TN_/16) x (TK_/16) per thread
template <int TN_, int TK_>
__global__ void __launch_bounds__(256)
sgemm_tn_fast_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ Bm, long long ldb,
                     float* __restrict__ C, long long ldc, float* __restrict__ cb,
                     long long M, int N, int K, long long rows_per_slice) {
  constexpr int RN = TN_ / 16, RK = TK_ / 16, A4 = TN_ / 4, B4 = TK_ / 4;
  __shared__ __align__(16) float as[2][SG_K][TN_ + 4];
  __shared__ __align__(16) float bs[2][SG_K][TK_ + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * TN_, k0 = blockIdx.y * TK_;
  const long long mbeg = static_cast<long long>(blockIdx.z) * rows_per_slice;
  const long long mend = min(M, mbeg + rows_per_slice);
  constexpr int NA = SG_K * A4 / 256, NB = SG_K * B4 / 256;   // float4 loads per thread
  float4 ar[NA], br[NB];
  auto gload = [&](long long m0) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + 256 * i, r = idx / A4, c4 = idx - r * A4;
      const long long m = m0 + r;
      const int n = n0 + 4 * c4;
      ar[i] = (m < mend && n < N) ? ldg_f4(A + m * lda + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = tid + 256 * i, r = idx / B4, c4 = idx - r * B4;
      const long long m = m0 + r;
      const int k = k0 + 4 * c4;
      br[i] = (m < mend && k < K) ? ldg_f4(Bm + m * ldb + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int b) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int idx = tid + 256 * i, r = idx / A4, c4 = idx - r * A4;
      *reinterpret_cast<float4*>(&as[b][r][4 * c4]) = ar[i];
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int idx = tid + 256 * i, r = idx / B4, c4 = idx - r * B4;
      *reinterpret_cast<float4*>(&bs[b][r][4 * c4]) = br[i];
    }
  };
  float acc[RN][RK] = {};
  float bsum[RN] = {};
  gload(mbeg);
  sstore(0);
  __syncthreads();
  int b = 0;
  for (long long m0 = mbeg; m0 < mend; m0 += SG_K, b ^= 1) {
    const bool more = m0 + SG_K < mend;
    if (more) gload(m0 + SG_K);
#pragma unroll
    for (int r = 0; r < SG_K; ++r) {
      float av[RN], bw[RK];
#pragma unroll
      for (int u = 0; u < RN; u += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&as[b][r][ty * RN + u]);
        av[u] = t.x; av[u + 1] = t.y; av[u + 2] = t.z; av[u + 3] = t.w;
      }
#pragma unroll
      for (int v = 0; v < RK; v += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&bs[b][r][tx * RK + v]);
        bw[v] = t.x; bw[v + 1] = t.y; bw[v + 2] = t.z; bw[v + 3] = t.w;
      }
#pragma unroll
      for (int u = 0; u < RN; ++u) {
        if (tx == 0) bsum[u] += av[u];
#pragma unroll
        for (int v = 0; v < RK; ++v) acc[u][v] = fmaf(av[u], bw[v], acc[u][v]);
      }
    }
    if (more) sstore(b ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < RN; ++u) {
    const int n = n0 + ty * RN + u;
    if (n >= N) continue;
    if (cb != nullptr && blockIdx.y == 0 && tx == 0) atomicAdd(cb + n, bsum[u]);
#pragma unroll
    for (int v = 0; v < RK; ++v) {
      const int k = k0 + tx * RK + v;
      if (k < K) atomicAdd(C + static_cast<long long>(n) * ldc + k, acc[u][v]);
    }
  }
}

// ---------------------------------------------------------------------------
// Narrow heads (the T-wide distribution head and the 1-wide factor head, N <= 16): the 64-wide tiles above
// would idle 90 % of their lanes; these two are plain streaming kernels.
//   dgrad:  dX[m][k] (+)= sum_n dY[m][n] W[n][k], optional ReLU mask;  thread = (row, 4 columns)
//   wgrad:  dW[n][k] += sum_m dY[m][n] X[m][k], db[n] += sum_m dY[m][n];  thread = column k, rows sliced
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dgrad_narrow_kernel(const float* __restrict__ dY, long long ldy, const float* __restrict__ W, long long ldw,
                    float* __restrict__ dX, long long ldx, long long M, int N, int K, int accumulate,
                    const float* __restrict__ ref, long long ldr) {
  extern __shared__ __align__(16) float wsm[];          // [N][K]
  for (int i = threadIdx.x; i < N * K; i += 256) wsm[i] = W[static_cast<long long>(i / K) * ldw + (i % K)];
  __syncthreads();
  const int k4n = K >> 2;
  const long long total = M * k4n;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * 256) {
    const long long m = i / k4n;
    const int k = static_cast<int>(i - m * k4n) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = 0; n < N; ++n) {
      const float y = __ldg(dY + m * ldy + n);
      const float4 w = *reinterpret_cast<const float4*>(wsm + n * K + k);
      acc.x = fmaf(y, w.x, acc.x); acc.y = fmaf(y, w.y, acc.y); acc.z = fmaf(y, w.z, acc.z); acc.w = fmaf(y, w.w, acc.w);
    }
    if (ref != nullptr) {
      const float4 r = ldg_f4(ref + m * ldr + k);
      if (!(r.x > 0.f)) acc.x = 0.f;
      if (!(r.y > 0.f)) acc.y = 0.f;
      if (!(r.z > 0.f)) acc.z = 0.f;
      if (!(r.w > 0.f)) acc.w = 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(dX + m * ldx + k);
    if (accumulate) { const float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    *dst = acc;
  }
}

__global__ void __launch_bounds__(256)
wgrad_narrow_kernel(const float* __restrict__ dY, long long ldy, const float* __restrict__ X, long long ldx,
                    float* __restrict__ dW, long long ldc, float* __restrict__ db,
                    long long M, int N, int K, long long rows_per_slice) {
  const int k = threadIdx.x;                             // blockDim.x == K (<= 256)
  const long long mbeg = static_cast<long long>(blockIdx.x) * rows_per_slice;
  const long long mend = min(M, mbeg + rows_per_slice);
  float acc[16], bsum[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) { acc[n] = 0.f; bsum[n] = 0.f; }
  for (long long m0 = mbeg; m0 < mend; m0 += 8) {        // 8 independent row loads in flight per thread
    float x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = (m0 + r < mend) ? __ldg(X + (m0 + r) * ldx + k) : 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (m0 + r < mend) {
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          if (n < N) {
            const float y = __ldg(dY + (m0 + r) * ldy + n);
            acc[n] = fmaf(y, x[r], acc[n]);
            if (k == 0) bsum[n] += y;
          }
        }
      }
    }
  }
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    if (n < N) {
      atomicAdd(dW + static_cast<long long>(n) * ldc + k, acc[n]);
      if (k == 0 && db != nullptr) atomicAdd(db + n, bsum[n]);
    }
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int sgemm_nt(const float* X, long long ldx, const float* W, long long ldw, const float* bias,
                    float* Y, long long ldy, long long M, int N, int K, int relu, int accumulate,
                    cudaStream_t st) {
  if (M <= 0) return GN_OK;
  if ((K % SG_K) == 0 && (ldx & 3) == 0 && (ldw & 3) == 0 && al16(X) && al16(W) && N >= 32) {
    dim3 grid(static_cast<unsigned>((M + 127) / 128), (N + 63) / 64);
    { ProfScope ps__("bwd_sgemm", st);
      sgemm128_kernel<true><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, Y, ldy, M, N, K, relu, accumulate, nullptr, 0); }
    GN_LAUNCH_CHECK();
    return GN_OK;
  }
  dim3 grid(static_cast<unsigned>((M + SG_T - 1) / SG_T), (N + SG_T - 1) / SG_T);
  { ProfScope ps__("bwd_sgemm", st);
    sgemm_kernel<true><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, Y, ldy, M, N, K, relu, accumulate, nullptr, 0); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}
// dX[M x K] = dY[M x N] * W[N x K]  (W in nn.Linear layout), optional ReLU mask by `ref` (same shape as dX)
static int sgemm_dgrad(const float* dY, long long ldy, const float* W, long long ldw, float* dX, long long ldx,
                       long long M, int N, int K, int accumulate, const float* ref, long long ldr,
                       cudaStream_t st) {
  if (M <= 0) return GN_OK;
  if (N <= 16 && (K & 3) == 0 && (ldx & 3) == 0 && al16(dX) && (ref == nullptr || ((ldr & 3) == 0 && al16(ref))) &&
      static_cast<size_t>(N) * K * 4 <= 48 * 1024) {
    const long long total = M * (K >> 2);
    long long blocks = (total + 255) / 256;
    if (blocks > GN_SM_COUNT * 16) blocks = GN_SM_COUNT * 16;
    { ProfScope ps__("bwd_sgemm", st);
      dgrad_narrow_kernel<<<static_cast<unsigned>(blocks), 256, static_cast<size_t>(N) * K * 4, st>>>(
          dY, ldy, W, ldw, dX, ldx, M, N, K, accumulate, ref, ldr); }
    GN_LAUNCH_CHECK();
    return GN_OK;
  }
  // here the contraction runs over N (the Linear's outputs) and the output has K columns
  if ((N % SG_K) == 0 && (ldy & 3) == 0 && (ldw & 3) == 0 && al16(dY) && al16(W) && K >= 32) {
    dim3 grid(static_cast<unsigned>((M + 127) / 128), (K + 63) / 64);
    { ProfScope ps__("bwd_sgemm", st);
      sgemm128_kernel<false><<<grid, 256, 0, st>>>(dY, ldy, W, ldw, nullptr, dX, ldx, M, K, N, 0, accumulate, ref, ldr); }
    GN_LAUNCH_CHECK();
    return GN_OK;
  }
  dim3 grid(static_cast<unsigned>((M + SG_T - 1) / SG_T), (K + SG_T - 1) / SG_T);
  { ProfScope ps__("bwd_sgemm", st);
    sgemm_kernel<false><<<grid, 256, 0, st>>>(dY, ldy, W, ldw, nullptr, dX, ldx, M, K, N, 0, accumulate, ref, ldr); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}
static int sgemm_wgrad(const float* dY, long long ldy, const float* X, long long ldx, float* dW, float* db,
                       long long M, int N, int K, cudaStream_t st) {
  if (M <= 0 || dW == nullptr) return GN_OK;
  if (N <= 16 && K <= 256 && (K & 31) == 0) {
    long long slices = (M + 2047) / 2048;
    if (slices < 4 * GN_SM_COUNT) slices = 4 * GN_SM_COUNT;
    if (slices > (M + 63) / 64) slices = (M + 63) / 64;
    if (slices < 1) slices = 1;
    const long long rps = (M + slices - 1) / slices;
    slices = (M + rps - 1) / rps;
    { ProfScope ps__("bwd_wgrad", st);
      wgrad_narrow_kernel<<<static_cast<unsigned>(slices), K, 0, st>>>(dY, ldy, X, ldx, dW, K, db, M, N, K, rps); }
    GN_LAUNCH_CHECK();
    return GN_OK;
  }
  // row slices: at most 4096 rows each, but enough of them to fill the GPU twice when M is small
  // (the hyper layers have 12x / 121x fewer edge rows than the pairwise layer)
  const long long tiles = static_cast<long long>((N + 127) / 128) * ((K + 63) / 64);
  long long want = (2 * GN_SM_COUNT + tiles - 1) / tiles;
  long long slices = (M + 4095) / 4096;
  if (slices < want) slices = want;
  if (slices > (M + 255) / 256) slices = (M + 255) / 256;
  if (slices < 1) slices = 1;
  if (slices > 1024) slices = 1024;
  const long long rps = ((M + slices - 1) / slices + SG_K - 1) / SG_K * SG_K;
  slices = (M + rps - 1) / rps;
  if ((ldy & 3) == 0 && (ldx & 3) == 0 && (N & 3) == 0 && (K & 3) == 0 && al16(dY) && al16(X) && N >= 32 && K >= 32) {
    ProfScope ps__("bwd_wgrad", st);
    if (N >= 128) {
      dim3 grid((N + 127) / 128, (K + 63) / 64, static_cast<unsigned>(slices));
      sgemm_tn_fast_kernel<128, 64><<<grid, 256, 0, st>>>(dY, ldy, X, ldx, dW, K, db, M, N, K, rps);
    } else {
      dim3 grid((N + 63) / 64, (K + 127) / 128, static_cast<unsigned>(slices));
      sgemm_tn_fast_kernel<64, 128><<<grid, 256, 0, st>>>(dY, ldy, X, ldx, dW, K, db, M, N, K, rps);
    }
    GN_LAUNCH_CHECK();
    return GN_OK;
  }
  dim3 grid((N + SG_T - 1) / SG_T, (K + SG_T - 1) / SG_T, static_cast<unsigned>(slices));
  { ProfScope ps__("bwd_wgrad", st);
    sgemm_tn_kernel<<<grid, 256, 0, st>>>(dY, ldy, X, ldx, dW, K, db, M, N, K, rps); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

// ---------------------------------------------------------------------------
// structural kernels
// ---------------------------------------------------------------------------
// inc[r] = [agg[r] | h[r]] / N
__global__ void build_inc_kernel(const float* __restrict__ agg, const float* __restrict__ h, float* __restrict__ inc,
                                 long long R, int D, float n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * 2 * D) return;
  const long long r = i / (2 * D);
  const int c = static_cast<int>(i - r * 2 * D);
  inc[i] = __fdiv_rn(c < D ? agg[r * D + c] : h[r * D + c - D], n);
}
// d_agg = d_inc[:, :D] / N ; d_h = d_inc[:, D:] / N
__global__ void split_inc_kernel(const float* __restrict__ dinc, float* __restrict__ dagg, float* __restrict__ dh,
                                 long long R, int D, float n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * 2 * D) return;
  const long long r = i / (2 * D);
  const int c = static_cast<int>(i - r * 2 * D);
  const float v = __fdiv_rn(dinc[i], n);
  if (c < D) dagg[r * D + c] = v; else dh[r * D + c - D] = v;
}
// incidence value H[e,n] of scene b (pairwise: implicit, self loops 2)
__device__ __forceinline__ float inc_val(const float* __restrict__ H, long long hstride, int pairwise,
                                         long long b, int e, int n, int N) {
  if (pairwise) {
    const int i = e / N, j = e - i * N;
    return (n == i ? 1.f : 0.f) + (n == j ? 1.f : 0.f);
  }
  return H[b * hstride + static_cast<long long>(e) * N + n];
}
// out_e[b,e,:] = sum_n H[e,n] in_n[b,n,:]        (edges <- nodes)
__global__ void inc_gather_kernel(const float* __restrict__ H, long long hstride, int pairwise,
                                  const float* __restrict__ in_n, float* __restrict__ out_e,
                                  long long B, int N, int E, int D) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= B * E * D) return;
  const int c = static_cast<int>(i % D);
  const long long be = i / D;
  const int e = static_cast<int>(be % E);
  const long long b = be / E;
  float s = 0.f;
  if (pairwise) {
    const int p = e / N, q = e - p * N;
    s = in_n[(b * N + p) * D + c] + in_n[(b * N + q) * D + c];
  } else {
    for (int n = 0; n < N; ++n) {
      const float w = H[b * hstride + static_cast<long long>(e) * N + n];
      if (w != 0.f) s = fmaf(w, in_n[(b * N + n) * D + c], s);
    }
  }
  out_e[i] = s;
}
// out_n[b,n,:] (+)= sum_e H[e,n] in_e[b,e,:]    (nodes <- edges)
__global__ void inc_scatter_kernel(const float* __restrict__ H, long long hstride, int pairwise,
                                   const float* __restrict__ in_e, float* __restrict__ out_n,
                                   long long B, int N, int E, int D, int accumulate) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= B * N * D) return;
  const int c = static_cast<int>(i % D);
  const long long bn = i / D;
  const int n = static_cast<int>(bn % N);
  const long long b = bn / N;
  float s = 0.f;
  if (pairwise) {
    for (int j = 0; j < N; ++j)
      s += in_e[(b * E + n * N + j) * D + c] + in_e[(b * E + j * N + n) * D + c];
  } else {
    for (int e = 0; e < E; ++e) {
      const float w = H[b * hstride + static_cast<long long>(e) * N + n];
      if (w != 0.f) s = fmaf(w, in_e[(b * E + e) * D + c], s);
    }
  }
  out_n[i] = accumulate ? out_n[i] + s : s;
}
// col[r*ldc] = sum_c a[r][c] * b[r][c]
__global__ void rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ col,
                              long long ldc, long long R, int D) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float s = 0.f;
  for (int c = 0; c < D; ++c) s = fmaf(a[r * D + c], b[r * D + c], s);
  col[r * ldc] = s;
}
// out[r][c] = scale[r*lds] * in[r][c]
__global__ void rowscale_kernel(const float* __restrict__ in, const float* __restrict__ scale, long long lds,
                                float* __restrict__ out, long long R, int D) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= R * D) return;
  out[i] = scale[(i / D) * lds] * in[i];
}
// fused pair of the two above for the aggregation MLPs: one warp per row, one pass over d_ef:
//   col[r*ldc] = d_ef[r] . v[r]   and   dv[r] = scale[r*lds] * d_ef[r]
__global__ void __launch_bounds__(256)
rowdot_scale_kernel(const float* __restrict__ def, const float* __restrict__ v, float* __restrict__ col, long long ldc,
                    const float* __restrict__ scale, long long lds, float* __restrict__ dv, long long R, int D) {
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float sc = __ldg(scale + r * lds);
  float s = 0.f;
  for (int c = 4 * lane; c < D; c += 128) {
    const float4 a = ldg_f4(def + r * D + c), b = ldg_f4(v + r * D + c);
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    *reinterpret_cast<float4*>(dv + r * D + c) = make_float4(sc * a.x, sc * a.y, sc * a.z, sc * a.w);
  }
  s = warp_sum(s);
  if (lane == 0) col[r * ldc] = s;
}

// Row part of one aggregation MLP's backward (one warp per edge row, hidden width 128).  With
// g = d_ef W1 (the un-scaled hidden gradient), u = relu(W0 eo + b0), w = edge_feat_t:
//   d edge_feat_t = u . g + d_ef . b1     (= d_ef . (W1 u + b1): the second Linear is never re-run)
//   du = w * g * (u > 0)  (in place over g)      dv = w * d_ef  (the dY operand of dW1 / db1)
__global__ void __launch_bounds__(256)
agg_bwd_rows_kernel(const float* __restrict__ u, float* __restrict__ g, const float* __restrict__ def,
                    const float* __restrict__ b1, const float* __restrict__ w, long long ldw,
                    float* __restrict__ d_w, long long ldd, float* __restrict__ dv, long long R, int D) {
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  const float sc = __ldg(w + r * ldw);
  const float4 uu = ldg_f4(u + r * 128 + 4 * lane);
  float4 gg = *reinterpret_cast<const float4*>(g + r * 128 + 4 * lane);
  float s = uu.x * gg.x;
  s = fmaf(uu.y, gg.y, s); s = fmaf(uu.z, gg.z, s); s = fmaf(uu.w, gg.w, s);
  gg.x = uu.x > 0.f ? sc * gg.x : 0.f; gg.y = uu.y > 0.f ? sc * gg.y : 0.f;
  gg.z = uu.z > 0.f ? sc * gg.z : 0.f; gg.w = uu.w > 0.f ? sc * gg.w : 0.f;
  *reinterpret_cast<float4*>(g + r * 128 + 4 * lane) = gg;
  for (int c = 4 * lane; c < D; c += 128) {
    const float4 a = ldg_f4(def + r * D + c), b = ldg_f4(b1 + c);
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    *reinterpret_cast<float4*>(dv + r * D + c) = make_float4(sc * a.x, sc * a.y, sc * a.z, sc * a.w);
  }
  s = warp_sum(s);
  if (lane == 0) d_w[r * ldd] = s;
}

// Pairwise aggregation backward on NODE rows (the collapse of the forward): with
// P_t = h W0_t^T, Q_t = d_agg W1_t (both R x 128), c_t[n] = d_agg[n] . b1_t and, per edge e = (i, j) of a scene,
//   u = relu(P_t[i] + P_t[j] + b0_t),   d_ef_e = d_agg[i] + d_agg[j]      (self loop: incidence 2, :124)
//   d edge_feat[e, t] = (Q_t[i] + Q_t[j]) . u + c_t[i] + c_t[j]
//   d pre_e = edge_feat[e, t] * (Q_t[i] + Q_t[j]) * (pre > 0);   dP_t[i] += d pre_e,  dP_t[j] += d pre_e
//   G_t[i] += edge_feat[e, t] * u, G_t[j] likewise   (dW1_t = d_agg^T G_t, as in the forward)
// the edge-level GEMMs (N^2 rows) of the as-written backward become node-level ones (N rows); this kernel does
// the per-scene O(N^2 * 128) part.  One 128-thread CTA per (scene, t) item, thread = hidden column; (i, j)
// and (j, i) share u, so the loop runs over unordered pairs.
struct PairBwdPtrs { const float* b0[15]; const float* b1[15]; float* db0[15]; float* db1[15]; };

__global__ void __launch_bounds__(128)
pair_agg_bwd_kernel(const float* __restrict__ Pn, const float* __restrict__ Qn, const float* __restrict__ efeat,
                    const float* __restrict__ d_agg, PairBwdPtrs ptrs,
                    float* __restrict__ dPn, float* __restrict__ Gn, float* __restrict__ d_efeat,
                    long long B, int N, int T, int D) {
  extern __shared__ __align__(16) float sm[];
  float* Ps = sm;                    // [N][128]
  float* Qs = Ps + N * 128;          // [N][128]
  float* dPs = Qs + N * 128;         // [N][128]
  float* Gs = dPs + N * 128;         // [N][128]
  float* qs = Gs + N * 128;          // [N][D]
  float* ws = qs + N * D;            // [N*N]  edge_feat[:, t] of the scene
  float* dws = ws + N * N;           // [N*N]  (Q_i + Q_j) . u per unordered pair, stored at [i][j], i <= j
  float* cs = dws + N * N;           // [N]
  float* Ss = cs + N;                // [N]
  const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
  const int E = N * N, LD = T * 128;
  const long long items = B * T;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const long long b = item / T;
    const int t = static_cast<int>(item - b * T);
    __syncthreads();
    for (int n = 0; n < N; ++n) {
      const long long row = b * N + n;
      Ps[n * 128 + c] = Pn[row * LD + t * 128 + c];
      Qs[n * 128 + c] = Qn[row * LD + t * 128 + c];
      dPs[n * 128 + c] = 0.f;
      Gs[n * 128 + c] = 0.f;
    }
    for (int i = c; i < N * D; i += 128) qs[i] = d_agg[b * N * D + i];
    for (int e = c; e < E; e += 128) { ws[e] = efeat[(b * E + e) * T + t]; dws[e] = 0.f; }
    __syncthreads();
    for (int n = warp; n < N; n += 4) {                  // c_t[n] = d_agg[n] . b1_t
      float s = 0.f;
      for (int k = lane; k < D; k += 32) s = fmaf(qs[n * D + k], __ldg(ptrs.b1[t] + k), s);
      s = warp_sum(s);
      if (lane == 0) cs[n] = s;
    }
    const float bb = __ldg(ptrs.b0[t] + c);
    float db0 = 0.f;
    for (int i = 0; i < N; ++i) {
      const float pi = Ps[i * 128 + c], qi = Qs[i * 128 + c];
      float dpi = 0.f, gi = 0.f;
      {                                                  // self loop (i, i): eo = 2 h_i, d_ef = 2 d_agg[i]
        const float pre = 2.f * pi + bb, u = fmaxf(pre, 0.f), t1 = 2.f * qi, w = ws[i * N + i];
        const float dwp = warp_sum(t1 * u);
        if (lane == 0) atomicAdd(&dws[i * N + i], dwp);
        const float dpre = pre > 0.f ? w * t1 : 0.f;
        dpi += 2.f * dpre; gi += 2.f * w * u; db0 += dpre;
      }
      for (int j = i + 1; j < N; ++j) {
        const float pre = pi + Ps[j * 128 + c] + bb, u = fmaxf(pre, 0.f), t1 = qi + Qs[j * 128 + c];
        const float w = ws[i * N + j] + ws[j * N + i];
        const float dwp = warp_sum(t1 * u);
        if (lane == 0) atomicAdd(&dws[i * N + j], dwp);
        const float dpre = pre > 0.f ? w * t1 : 0.f;
        dpi += dpre; gi += w * u; db0 += dpre;
        dPs[j * 128 + c] += dpre;
        Gs[j * 128 + c] += w * u;
      }
      dPs[i * 128 + c] += dpi;
      Gs[i * 128 + c] += gi;
    }
    if (c < N) {                                         // S_t[n] = sum of edge_feat over the edges incident to n
      float s = 0.f;
      for (int j = 0; j < N; ++j) s += ws[c * N + j] + ws[j * N + c];
      Ss[c] = s;
    }
    __syncthreads();
    for (int n = 0; n < N; ++n) {
      const long long row = b * N + n;
      dPn[row * LD + t * 128 + c] = dPs[n * 128 + c];
      Gn[row * LD + t * 128 + c] = Gs[n * 128 + c];
    }
    for (int e = c; e < E; e += 128) {
      const int i = e / N, j = e - i * N;
      const float d = dws[i <= j ? e : j * N + i] + (i == j ? 2.f * cs[i] : cs[i] + cs[j]);
      d_efeat[(b * E + e) * T + t] = d;
    }
    if (ptrs.db0[t] != nullptr) atomicAdd(ptrs.db0[t] + c, db0);
    if (ptrs.db1[t] != nullptr) {
      for (int k = c; k < D; k += 128) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(Ss[n], qs[n * D + k], s);
        atomicAdd(ptrs.db1[t] + k, s);
      }
    }
  }
}

// Gumbel-softmax / sigmoid backward per edge row.  edge_feat = f * dist, sum_t dist = 1  =>
// f = sum_t edge_feat, dist = edge_feat / f.  y = (logits + g)/tau, tau = 1/2:
//   d_dist_t = d_ef_t * f (+ external d_dist_t);  d_f = sum_t d_ef_t dist_t
//   d_logits_t = 2 * dist_t * (d_dist_t - sum_s dist_s d_dist_s);  d_fl = d_f * f * (1 - f)
__global__ void gumbel_bwd_kernel(const float* __restrict__ efeat, const float* __restrict__ d_efeat,
                                  const float* __restrict__ d_dist_ext, float* __restrict__ d_logits,
                                  float* __restrict__ d_fl, long long R, int T) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float f = 0.f;
  for (int t = 0; t < T; ++t) f += efeat[r * T + t];
  // a sigmoid that underflowed to 0 makes dist = 0 / 0: the categorical is then unobservable from edge_feat, and every
  // gradient through this row carries the factor f = 0 anyway — use the uniform distribution instead of NaN
  const float inv = f > 0.f ? 1.f / f : 0.f, fallback = f > 0.f ? 0.f : 1.f / static_cast<float>(T);
  float df = 0.f, dot = 0.f;
  for (int t = 0; t < T; ++t) {
    const float dist = fmaf(efeat[r * T + t], inv, fallback);
    const float dd = d_efeat[r * T + t] * f + (d_dist_ext ? d_dist_ext[r * T + t] : 0.f);
    df = fmaf(d_efeat[r * T + t], dist, df);
    dot = fmaf(dist, dd, dot);
  }
  for (int t = 0; t < T; ++t) {
    const float dist = fmaf(efeat[r * T + t], inv, fallback);
    const float dd = d_efeat[r * T + t] * f + (d_dist_ext ? d_dist_ext[r * T + t] : 0.f);
    d_logits[r * T + t] = 2.f * dist * (dd - dot);
  }
  d_fl[r] = df * f * (1.f - f);
}

// node2edge backward, one thread per edge (members from the incidence), gradients accumulated into
// d_x', d_pq with atomics; attention-tail parameter gradients reduced per block then atomically.
//   forward: pe = sum_m H q_m + b0; a_m = w1 . relu(pn_m + pe) + b1; s_m = a_m H_m; p = softmax over all N
//            (non-members logit 0); wgt_m = p_m H_m; edges = sum_m wgt_m x'_m
__global__ void __launch_bounds__(128)
n2e_bwd_kernel(const float* __restrict__ xprime, const float* __restrict__ pq, const float* __restrict__ H,
               long long hstride, int pairwise, const float* __restrict__ att_b0, const float* __restrict__ att_w1,
               const float* __restrict__ att_b1, const float* __restrict__ d_edges,
               float* __restrict__ d_xprime, float* __restrict__ d_pq,
               float* __restrict__ d_b0, float* __restrict__ d_w1, float* __restrict__ d_b1,
               long long B, int N, int E) {
  __shared__ float sb0[32], sw1[32];
  __shared__ float g_b0[32], g_w1[32], g_b1;
  if (threadIdx.x < 32) { sb0[threadIdx.x] = att_b0[threadIdx.x]; sw1[threadIdx.x] = att_w1[threadIdx.x];
                          g_b0[threadIdx.x] = 0.f; g_w1[threadIdx.x] = 0.f; }
  if (threadIdx.x == 0) g_b1 = 0.f;
  __syncthreads();
  const float b1 = att_b1[0];
  const long long be = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float lw1[32], lb0[32], lb1 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) { lw1[k] = 0.f; lb0[k] = 0.f; }
  if (be < B * E) {
    const int e = static_cast<int>(be % E);
    const long long b = be / E;
    const float* xb = xprime + b * N * 64;
    const float* pb = pq + b * N * 64;
    // members
    int mem[GN_MAX_AGENTS]; float hv[GN_MAX_AGENTS];
    int cnt = 0;
    for (int n = 0; n < N; ++n) {
      const float w = inc_val(H, hstride, pairwise, b, e, n, N);
      if (w != 0.f) { mem[cnt] = n; hv[cnt] = w; ++cnt; }
    }
    float pe[32];
    for (int k = 0; k < 32; ++k) pe[k] = sb0[k];
    for (int m = 0; m < cnt; ++m)
      for (int k = 0; k < 32; ++k) pe[k] = fmaf(hv[m], pb[mem[m] * 64 + 32 + k], pe[k]);
    float s[GN_MAX_AGENTS];
    float mx = (cnt < N) ? 0.f : -INFINITY;
    for (int m = 0; m < cnt; ++m) {
      float a = 0.f;
      for (int k = 0; k < 32; ++k) a = fmaf(fmaxf(pb[mem[m] * 64 + k] + pe[k], 0.f), sw1[k], a);
      s[m] = (a + b1) * hv[m];
      mx = fmaxf(mx, s[m]);
    }
    float den = static_cast<float>(N - cnt) * expf(-mx);
    for (int m = 0; m < cnt; ++m) den += expf(s[m] - mx);
    // d_wgt_m = d_edges . x'_m ;  d_x'_m += wgt_m d_edges
    const float* de = d_edges + be * 64;
    float dp[GN_MAX_AGENTS];
    float pdot = 0.f;
    for (int m = 0; m < cnt; ++m) {
      const float p = expf(s[m] - mx) / den;
      const float wgt = p * hv[m];
      float dw = 0.f;
      for (int c = 0; c < 64; ++c) {
        dw = fmaf(de[c], xb[mem[m] * 64 + c], dw);
        atomicAdd(d_xprime + (b * N + mem[m]) * 64 + c, wgt * de[c]);
      }
      dp[m] = dw * hv[m];                 // d p_m
      pdot = fmaf(p, dp[m], pdot);
    }
    float dpe[32];
    for (int k = 0; k < 32; ++k) dpe[k] = 0.f;
    for (int m = 0; m < cnt; ++m) {
      const float p = expf(s[m] - mx) / den;
      const float da = p * (dp[m] - pdot) * hv[m];     // d a_m
      lb1 += da;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float pre = pb[mem[m] * 64 + k] + pe[k];
        const float act = fmaxf(pre, 0.f);
        lw1[k] = fmaf(act, da, lw1[k]);
        const float r = (pre > 0.f) ? sw1[k] * da : 0.f;
        atomicAdd(d_pq + (b * N + mem[m]) * 64 + k, r);       // d pn_m
        dpe[k] += r;
      }
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      lb0[k] += dpe[k];
      for (int m = 0; m < cnt; ++m) atomicAdd(d_pq + (b * N + mem[m]) * 64 + 32 + k, hv[m] * dpe[k]);   // d q_m
    }
  }
  // attention-tail parameter gradients: per-thread partials -> warp reduction -> one shared atomic per warp
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float a = warp_sum(lw1[k]), c = warp_sum(lb0[k]);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&g_w1[k], a); atomicAdd(&g_b0[k], c); }
  }
  {
    const float a = warp_sum(lb1);
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_b1, a);
  }
  __syncthreads();
  if (threadIdx.x < 32) { atomicAdd(d_b0 + threadIdx.x, g_b0[threadIdx.x]); atomicAdd(d_w1 + threadIdx.x, g_w1[threadIdx.x]); }
  if (threadIdx.x == 0) atomicAdd(d_b1, g_b1);
}

// ---------------------------------------------------------------------------
// host: workspace plan + orchestration
// ---------------------------------------------------------------------------
struct BwdPlan {
  size_t inc, o1, d_o1, d_inc, d_agg, d_ef, eo, d_eo, u, v, dv, du, d_efeat, d_logits, d_fl,
         z1, z, hd, hf, d_hd, d_z, d_z1, d_edges, d_x, d_pq, hid, d_hid, total;
  size_t Pn, Qn, dPn, Gn;            // pairwise collapse: node-level (R x T*128) tensors
};

static void make_bwd_plan(const gn_stage_cfg* c, BwdPlan& p) {
  const size_t R = static_cast<size_t>(c->B) * c->N, RE = static_cast<size_t>(c->B) * c->E;
  const size_t D = c->D, T = c->T;
  size_t o = 0;
  auto take = [&](size_t floats) { size_t at = o; o += round_up_sz(floats * 4, 256); return at; };
  p.inc = take(R * 2 * D); p.o1 = take(R * 128); p.d_o1 = take(R * 128); p.d_inc = take(R * 2 * D);
  p.d_agg = take(R * D);
  p.d_ef = p.eo = p.d_eo = p.u = p.v = p.dv = p.du = 0;
  p.Pn = p.Qn = p.dPn = p.Gn = 0;
  if (c->pairwise) {                 // the aggregation backward runs on node rows (see pair_agg_bwd_kernel)
    p.Pn = take(R * T * 128); p.Qn = take(R * T * 128); p.dPn = take(R * T * 128); p.Gn = take(R * T * 128);
  } else {
    p.d_ef = take(RE * D); p.eo = take(RE * D); p.d_eo = take(RE * D);
    p.u = take(RE * 128); p.v = take(RE * D); p.dv = take(RE * D); p.du = take(RE * 128);
  }
  p.d_efeat = take(RE * T); p.d_logits = take(RE * T); p.d_fl = take(RE);
  p.z1 = take(RE * 128); p.z = take(RE * 64); p.hd = take(RE * 128); p.hf = take(RE * 128);
  p.d_hd = take(RE * 128); p.d_z = take(RE * 64); p.d_z1 = take(RE * 128); p.d_edges = take(RE * 64);
  p.d_x = take(R * 64); p.d_pq = take(R * 64); p.hid = take(R * 256); p.d_hid = take(R * 256);
  p.total = o;
}

size_t stage_bwd_workspace_bytes(const gn_stage_cfg* c) {
  BwdPlan p;
  make_bwd_plan(c, p);
  return p.total;
}

#define GN_TRYB(expr) do { int rc__ = (expr); if (rc__ != GN_OK) return rc__; } while (0)

static inline unsigned nblk(long long n, int t = 256) { return static_cast<unsigned>((n + t - 1) / t); }

int stage_bwd(const gn_stage_cfg* c, const gn_train_params* P, const float* h, const float* H,
              const float* xprime, const float* pq, const float* edges, const float* efeat, const float* agg,
              const float* d_out, long long ld_dout, const float* d_dist, float* d_h,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  BwdPlan p;
  make_bwd_plan(c, p);
  if (ws_bytes < p.total) return GN_E_WORKSPACE;
  if (c->B == 0) return GN_OK;
  char* base = static_cast<char*>(ws);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(base + off); };
  const long long B = c->B;
  const int N = c->N, D = c->D, E = c->E, T = c->T, Dout = c->Dout;
  const long long R = B * N, RE = B * E;
  const long long hstride = c->h_stride > 0 ? c->h_stride : static_cast<long long>(E) * N;
  const float fN = static_cast<float>(N);
  const gn_lin& Ln0 = P->node0; const gn_lin& Ln1 = P->node1; const gn_lin& Lpq = P->attpq;
  const gn_lin& Li0 = P->init0; const gn_lin& Li1 = P->init1;
  const gn_lin& Ld0 = P->dist0; const gn_lin& Ld1 = P->dist1; const gn_lin& Lf0 = P->fac0; const gn_lin& Lf1 = P->fac1;
  const gn_lin& Lp0 = P->post0; const gn_lin& Lp1 = P->post1;

  // ---- closing MLP: out = W1 relu(W0 inc + b0) + b1, inc = [agg | h] / N
  build_inc_kernel<<<nblk(R * 2 * D), 256, 0, st>>>(agg, h, F(p.inc), R, D, fN);
  GN_TRYB(sgemm_nt(F(p.inc), 2 * D, Lp0.W, 2 * D, Lp0.b, F(p.o1), 128, R, 128, 2 * D, 1, 0, st));
  GN_TRYB(sgemm_wgrad(d_out, ld_dout, F(p.o1), 128, Lp1.dW, Lp1.db, R, Dout, 128, st));
  GN_TRYB(sgemm_dgrad(d_out, ld_dout, Lp1.W, 128, F(p.d_o1), 128, R, Dout, 128, 0, F(p.o1), 128, st));
  GN_TRYB(sgemm_wgrad(F(p.d_o1), 128, F(p.inc), 2 * D, Lp0.dW, Lp0.db, R, 128, 2 * D, st));
  GN_TRYB(sgemm_dgrad(F(p.d_o1), 128, Lp0.W, 2 * D, F(p.d_inc), 2 * D, R, 128, 2 * D, 0, nullptr, 0, st));
  split_inc_kernel<<<nblk(R * 2 * D), 256, 0, st>>>(F(p.d_inc), F(p.d_agg), d_h, R, D, fN);
  if (c->pairwise) {
    // ---- pairwise: aggregation backward on node rows (pair_agg_bwd_kernel)
    const int LD = T * 128;
    PairBwdPtrs pp;
    memset(&pp, 0, sizeof(pp));
    for (int t = 0; t < T; ++t) {
      const gn_lin& A0 = P->agg0[t]; const gn_lin& A1 = P->agg1[t];
      pp.b0[t] = A0.b; pp.b1[t] = A1.b; pp.db0[t] = A0.db; pp.db1[t] = A1.db;
      GN_TRYB(sgemm_nt(h, D, A0.W, D, nullptr, F(p.Pn) + t * 128, LD, R, 128, D, 0, 0, st));              // P_t = h W0_t^T
      GN_TRYB(sgemm_dgrad(F(p.d_agg), D, A1.W, 128, F(p.Qn) + t * 128, LD, R, D, 128, 0, nullptr, 0, st));  // Q_t = d_agg W1_t
    }
    {
      const size_t smem = (static_cast<size_t>(4) * N * 128 + static_cast<size_t>(N) * D + 2 * static_cast<size_t>(N) * N + 2 * N) * 4;
      cudaError_t e = cudaFuncSetAttribute(pair_agg_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return static_cast<int>(e);
      int per_sm = static_cast<int>((200 * 1024) / (smem + 1024));
      if (per_sm < 1) per_sm = 1;
      if (per_sm > 12) per_sm = 12;
      const long long items = B * T;
      const long long cap = static_cast<long long>(GN_SM_COUNT) * per_sm;
      const unsigned grid = static_cast<unsigned>(items < cap ? items : cap);
      ProfScope ps__("bwd_pair_agg", st);
      pair_agg_bwd_kernel<<<grid, 128, smem, st>>>(F(p.Pn), F(p.Qn), efeat, F(p.d_agg), pp, F(p.dPn), F(p.Gn),
                                                   F(p.d_efeat), B, N, T, D);
    }
    GN_LAUNCH_CHECK();
    for (int t = 0; t < T; ++t) {
      const gn_lin& A0 = P->agg0[t]; const gn_lin& A1 = P->agg1[t];
      GN_TRYB(sgemm_wgrad(F(p.dPn) + t * 128, LD, h, D, A0.dW, nullptr, R, 128, D, st));                  // dW0_t = dP_t^T h
      GN_TRYB(sgemm_dgrad(F(p.dPn) + t * 128, LD, A0.W, D, d_h, D, R, 128, D, 1, nullptr, 0, st));         // dh += dP_t W0_t
      GN_TRYB(sgemm_wgrad(F(p.d_agg), D, F(p.Gn) + t * 128, LD, A1.dW, nullptr, R, D, 128, st));          // dW1_t = d_agg^T G_t
    }
  } else {
  // ---- agg = H^T ef  =>  d_ef = H d_agg ;  eo = H h
    { ProfScope ps__("bwd_gather", st);
      inc_gather_kernel<<<nblk(RE * D), 256, 0, st>>>(H, hstride, c->pairwise, F(p.d_agg), F(p.d_ef), B, N, E, D);
      inc_gather_kernel<<<nblk(RE * D), 256, 0, st>>>(H, hstride, c->pairwise, h, F(p.eo), B, N, E, D); }
    GN_LAUNCH_CHECK();
    cudaMemsetAsync(F(p.d_eo), 0, static_cast<size_t>(RE) * D * 4, st);
    // ---- T aggregation MLPs as written: ef = sum_t efeat_t (W1_t relu(W0_t eo + b0_t) + b1_t)
    for (int t = 0; t < T; ++t) {
      const gn_lin& A0 = P->agg0[t]; const gn_lin& A1 = P->agg1[t];
      GN_TRYB(sgemm_nt(F(p.eo), D, A0.W, D, A0.b, F(p.u), 128, RE, 128, D, 1, 0, st));
      GN_TRYB(sgemm_dgrad(F(p.d_ef), D, A1.W, 128, F(p.du), 128, RE, D, 128, 0, nullptr, 0, st));   // g = d_ef W1
      { ProfScope ps__("bwd_rowops", st);
        agg_bwd_rows_kernel<<<nblk(RE, 8), 256, 0, st>>>(F(p.u), F(p.du), F(p.d_ef), A1.b, efeat + t, T,
                                                         F(p.d_efeat) + t, T, F(p.dv), RE, D); }
      GN_TRYB(sgemm_wgrad(F(p.dv), D, F(p.u), 128, A1.dW, A1.db, RE, D, 128, st));
      GN_TRYB(sgemm_wgrad(F(p.du), 128, F(p.eo), D, A0.dW, A0.db, RE, 128, D, st));
      GN_TRYB(sgemm_dgrad(F(p.du), 128, A0.W, D, F(p.d_eo), D, RE, 128, D, 1, nullptr, 0, st));
    }
    { ProfScope ps__("bwd_scatter", st);
      inc_scatter_kernel<<<nblk(R * D), 256, 0, st>>>(H, hstride, c->pairwise, F(p.d_eo), d_h, B, N, E, D, 1); }
  }
  // ---- edge_feat = sigmoid(fl) * softmax(2 (logits + g))
  gumbel_bwd_kernel<<<nblk(RE), 256, 0, st>>>(efeat, F(p.d_efeat), d_dist, F(p.d_logits), F(p.d_fl), RE, T);
  GN_LAUNCH_CHECK();
  // ---- MLP_dict_softmax MLPs (recompute z1, z, hd, hf)
  GN_TRYB(sgemm_nt(edges, 64, Li0.W, 64, Li0.b, F(p.z1), 128, RE, 128, 64, 1, 0, st));
  GN_TRYB(sgemm_nt(F(p.z1), 128, Li1.W, 128, Li1.b, F(p.z), 64, RE, 64, 128, 0, 0, st));
  GN_TRYB(sgemm_nt(F(p.z), 64, Ld0.W, 64, Ld0.b, F(p.hd), 128, RE, 128, 64, 1, 0, st));
  GN_TRYB(sgemm_nt(F(p.z), 64, Lf0.W, 64, Lf0.b, F(p.hf), 128, RE, 128, 64, 1, 0, st));
  GN_TRYB(sgemm_wgrad(F(p.d_logits), T, F(p.hd), 128, Ld1.dW, Ld1.db, RE, T, 128, st));
  GN_TRYB(sgemm_dgrad(F(p.d_logits), T, Ld1.W, 128, F(p.d_hd), 128, RE, T, 128, 0, F(p.hd), 128, st));
  GN_TRYB(sgemm_wgrad(F(p.d_hd), 128, F(p.z), 64, Ld0.dW, Ld0.db, RE, 128, 64, st));
  GN_TRYB(sgemm_dgrad(F(p.d_hd), 128, Ld0.W, 64, F(p.d_z), 64, RE, 128, 64, 0, nullptr, 0, st));
  GN_TRYB(sgemm_wgrad(F(p.d_fl), 1, F(p.hf), 128, Lf1.dW, Lf1.db, RE, 1, 128, st));
  GN_TRYB(sgemm_dgrad(F(p.d_fl), 1, Lf1.W, 128, F(p.d_hd), 128, RE, 1, 128, 0, F(p.hf), 128, st));
  GN_TRYB(sgemm_wgrad(F(p.d_hd), 128, F(p.z), 64, Lf0.dW, Lf0.db, RE, 128, 64, st));
  GN_TRYB(sgemm_dgrad(F(p.d_hd), 128, Lf0.W, 64, F(p.d_z), 64, RE, 128, 64, 1, nullptr, 0, st));
  GN_TRYB(sgemm_wgrad(F(p.d_z), 64, F(p.z1), 128, Li1.dW, Li1.db, RE, 64, 128, st));
  GN_TRYB(sgemm_dgrad(F(p.d_z), 64, Li1.W, 128, F(p.d_z1), 128, RE, 64, 128, 0, F(p.z1), 128, st));
  GN_TRYB(sgemm_wgrad(F(p.d_z1), 128, edges, 64, Li0.dW, Li0.db, RE, 128, 64, st));
  GN_TRYB(sgemm_dgrad(F(p.d_z1), 128, Li0.W, 64, F(p.d_edges), 64, RE, 128, 64, 0, nullptr, 0, st));
  // ---- node2edge: attention softmax + weighted gather
  cudaMemsetAsync(F(p.d_x), 0, static_cast<size_t>(R) * 64 * 4, st);
  cudaMemsetAsync(F(p.d_pq), 0, static_cast<size_t>(R) * 64 * 4, st);
  { ProfScope ps__("bwd_n2e", st);
  n2e_bwd_kernel<<<nblk(RE, 128), 128, 0, st>>>(xprime, pq, H, hstride, c->pairwise, P->att_b0, P->att_w1, P->att_b1,
                                               F(p.d_edges), F(p.d_x), F(p.d_pq), P->d_att_b0, P->d_att_w1,
                                               P->d_att_b1, B, N, E); }
  GN_LAUNCH_CHECK();
  // ---- pq = x' Wpq^T ; x' = W1 relu(W0 h + b0) + b1
  GN_TRYB(sgemm_wgrad(F(p.d_pq), 64, xprime, 64, Lpq.dW, nullptr, R, 64, 64, st));
  GN_TRYB(sgemm_dgrad(F(p.d_pq), 64, Lpq.W, 64, F(p.d_x), 64, R, 64, 64, 1, nullptr, 0, st));
  GN_TRYB(sgemm_nt(h, D, Ln0.W, D, Ln0.b, F(p.hid), 256, R, 256, D, 1, 0, st));
  GN_TRYB(sgemm_wgrad(F(p.d_x), 64, F(p.hid), 256, Ln1.dW, Ln1.db, R, 64, 256, st));
  GN_TRYB(sgemm_dgrad(F(p.d_x), 64, Ln1.W, 256, F(p.d_hid), 256, R, 64, 256, 0, F(p.hid), 256, st));
  GN_TRYB(sgemm_wgrad(F(p.d_hid), 256, h, D, Ln0.dW, Ln0.db, R, 256, D, st));
  GN_TRYB(sgemm_dgrad(F(p.d_hid), 256, Ln0.W, D, d_h, D, R, 256, D, 1, nullptr, 0, st));
  return GN_OK;
}

}  // namespace gn
