// fp32 FFMA micro-kernel used by every row-tile kernel of the 1e-5 parity path.
//
// A CTA of 256 threads (16 x 16) multiplies a TM-row activation tile held
// K-major in shared memory (At[k][row], leading dimension lda = TM + 4) with a
// K-major weight matrix streamed from L2 through a double-buffered cp.async
// panel of KC = 16 k-rows.  Thread (ty, tx) owns rows ty*RM..+RM and the RN
// columns { n0 + tx + 16*j }.  Per k it issues RM/4 + RN/4 128-bit shared loads
// for RM*RN FFMAs (8x8: 4 LDS.128 per 64 FFMA).
//
// The host packs each weight matrix with its columns permuted inside every
// TN-wide chunk (groupnet_b200/packing.py::permute_cols) so that a thread's RN
// columns are two contiguous 4-float groups in the panel: position
// (j/4)*64 + tx*4 + (j%4) holds original column tx + 16*j.  Panel loads are
// then conflict-free LDS.128 and the K-major epilogue stores (lane stride =
// one column = lda floats, lda % 32 == 4) are conflict-free STS.128.
#pragma once
#include "gn_common.cuh"

namespace gn {

constexpr int KC = 16;

template <int TM>
struct TileGeom {
  static constexpr int LD = TM + 4;     // K-major leading dimension (floats)
  static constexpr int RM = TM / 16;
};

template <int RM, int RN>
__device__ __forceinline__ void acc_zero(float (&acc)[RM][RN]) {
#pragma unroll
  for (int r = 0; r < RM; ++r)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[r][j] = 0.f;
}

// acc += At(TM x K) * Wt[:, n0 .. n0+TN)      K % 16 == 0
// Wt is the packed (column-permuted) K-major weight; wp is 2*KC*TN floats of smem.
template <int TM, int TN>
__device__ __forceinline__ void gemm_accum(float (&acc)[TM / 16][TN / 16],
                                           const float* At, const float* __restrict__ Wt,
                                           int ldw, int n0, int K, float* wp) {
  constexpr int RM = TM / 16, RN = TN / 16, LD = TM + 4;
  static_assert(RM % 4 == 0 && RN % 4 == 0, "thread tile must be float4-sized");
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int npanel = K / KC;
  auto load_panel = [&](int p, int buf) {
    const float* src = Wt + static_cast<size_t>(p) * KC * ldw + n0;
    float* dst = wp + buf * (KC * TN);
#pragma unroll
    for (int i = tid; i < KC * TN / 4; i += GN_THREADS) {
      int k = i / (TN / 4), c = i - k * (TN / 4);
      cp_async16(dst + k * TN + 4 * c, src + static_cast<size_t>(k) * ldw + 4 * c);
    }
    cp_async_commit();
  };
  load_panel(0, 0);
  for (int p = 0; p < npanel; ++p) {
    cp_async_wait<0>();
    __syncthreads();                       // panel p visible; everyone done with panel p-1
    if (p + 1 < npanel) load_panel(p + 1, (p + 1) & 1);
    const float* wb = wp + (p & 1) * (KC * TN) + tx * 4;
    const float* ab = At + static_cast<size_t>(p) * KC * LD + ty * RM;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      float a[RM], w[RN];
#pragma unroll
      for (int u = 0; u < RM / 4; ++u) {
        float4 v = *reinterpret_cast<const float4*>(ab + k * LD + 4 * u);
        a[4 * u] = v.x; a[4 * u + 1] = v.y; a[4 * u + 2] = v.z; a[4 * u + 3] = v.w;
      }
#pragma unroll
      for (int u = 0; u < RN / 4; ++u) {
        float4 v = *reinterpret_cast<const float4*>(wb + k * TN + 64 * u);
        w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
      }
#pragma unroll
      for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[r][j] = fmaf(a[r], w[j], acc[r][j]);
    }
  }
  __syncthreads();                         // At and wp may be overwritten by the caller
}

// Visit every accumulator: f(row_in_tile, col_in_chunk, value&).  col_in_chunk is
// the ORIGINAL (unpermuted) column offset inside the TN chunk: tx + 16*j.
template <int TM, int TN, typename F>
__device__ __forceinline__ void acc_foreach(float (&acc)[TM / 16][TN / 16], F f) {
  constexpr int RM = TM / 16, RN = TN / 16;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int j = 0; j < RN; ++j)
#pragma unroll
    for (int r = 0; r < RM; ++r) f(ty * RM + r, tx + 16 * j, acc[r][j]);
}

// Store relu/bias-processed accumulators K-major: dst[(c0 + col)*LD + row],
// 128-bit along rows.  g(col, value) -> value applies bias / activation.
template <int TM, int TN, typename G>
__device__ __forceinline__ void acc_store_kmajor(float (&acc)[TM / 16][TN / 16], float* dst,
                                                 int c0, G g) {
  constexpr int RM = TM / 16, RN = TN / 16, LD = TM + 4;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int j = 0; j < RN; ++j) {
    int col = tx + 16 * j;
    float* d = dst + static_cast<size_t>(c0 + col) * LD + ty * RM;
#pragma unroll
    for (int u = 0; u < RM / 4; ++u) {
      float4 v;
      v.x = g(col, ty * RM + 4 * u, acc[4 * u][j]);
      v.y = g(col, ty * RM + 4 * u + 1, acc[4 * u + 1][j]);
      v.z = g(col, ty * RM + 4 * u + 2, acc[4 * u + 2][j]);
      v.w = g(col, ty * RM + 4 * u + 3, acc[4 * u + 3][j]);
      *reinterpret_cast<float4*>(d + 4 * u) = v;
    }
  }
}

// Load a row-major global tile src[row0 .. row0+nrows)[0..K) (row stride lds)
// transposed into K-major shared memory dst[k][row], scaling by `scale`.
// Rows >= nrows and k in [K, Kp) are zero-filled.  K % 4 == 0.
template <int TM>
__device__ __forceinline__ void load_tile_kmajor(float* dst, const float* __restrict__ src,
                                                 size_t lds, int nrows, int K, int Kp, float scale,
                                                 bool divide) {
  constexpr int LD = TM + 4;
  constexpr int KG = GN_THREADS / TM;          // k-groups of 4 handled in parallel
  const int row = threadIdx.x % TM, kg = threadIdx.x / TM;
  for (int k4 = kg * 4; k4 < Kp; k4 += KG * 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < nrows && k4 < K) {
      v = ldg_f4(src + static_cast<size_t>(row) * lds + k4);
      if (divide) {
        v.x = __fdiv_rn(v.x, scale); v.y = __fdiv_rn(v.y, scale);
        v.z = __fdiv_rn(v.z, scale); v.w = __fdiv_rn(v.w, scale);
      }
    }
    dst[(k4 + 0) * LD + row] = v.x;
    dst[(k4 + 1) * LD + row] = v.y;
    dst[(k4 + 2) * LD + row] = v.z;
    dst[(k4 + 3) * LD + row] = v.w;
  }
}

}  // namespace gn
