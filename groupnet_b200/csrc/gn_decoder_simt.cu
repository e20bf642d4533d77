// Trajectory decoder, fp32 FFMA path: one launch per DecomposeBlock.
//
// model/GroupNet_nba.py:48-79 (DecomposeBlock.forward) and :461-505 (Decoder.forward): per row
// (= scene-agent x sample) and block
//   res   = x_true - x_hat                          (T_p, 2)
//   e_t   = relu(conv1d(res; 2 -> 32, k 3, pad 1))  (T_p, 32)
//   state = GRU(32 -> 96)(e_1 .. e_Tp), h0 = 0      gates (r, z, n): n = tanh(W_in e + b_in + r*(W_hn h + b_hn)),
//                                                   h' = (1 - z) n + z h
//   feat  = [past_feature ; z ; state]              (F + Z + 96)
//   x_hat = decoder_x(feat), y_hat = decoder_y(feat)   MLPs feat -> 512 -> 256 -> {2 T_p, 2 T_f}
// and over the blocks  reconstruction = sum x_hat,  prediction = sum y_hat,  out_seq = prediction + cur_location.
//
// A CTA owns a tile of 64 rows for the whole block: the feature tile, the GRU state and every hidden
// activation stay in shared memory (K-major), weights stream from L2 through the cp.async panels of
// gn_gemm_simt.cuh.  The three GRU gates are padded from 96 to 128 columns (zero weights) so each is one
// 128-column GEMM chunk; padded state columns stay exactly 0.  Hidden layer 1 (512) is produced 128 columns
// at a time and consumed immediately by the second Linear, so it is never materialised.
// out_seq / recover double as the running sums across blocks; x_hat lives in the caller's workspace.
#include "gn_gemm_simt.cuh"

namespace gn {

constexpr int DEC_TM = 64;
constexpr int DEC_LD = DEC_TM + 4;
constexpr int DEC_CONV = 32;        // conv_past channels
constexpr int DEC_STATE = 96;       // GRU hidden size
constexpr int DEC_GATE = 128;       // padded gate width
constexpr int DEC_H1 = 512, DEC_H2 = 256, DEC_OUTC = 64;
constexpr int DEC_RES_LD = 64;      // 2*T_p <= 64

struct DecArgs {
  gn_decoder_weights w;
  const float* past_feature; const float* z; const float* past_traj; const float* cur_location;
  float* x_hat; float* out_seq; float* recover;
  long long R;
  int S, F, Z, Tp, Tf, Kp, first, last;
};

__device__ __forceinline__ float dec_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }

// feat (TM x Kp, K-major) -> 512 -> 256 -> 64 (padded) ; returns the last Linear WITHOUT its bias
__device__ __forceinline__ void dec_mlp3(const float* featT, int Kp, float* hidT, float* wp,
                                         const float* __restrict__ w0, const float* __restrict__ b0,
                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                         const float* __restrict__ w2, float (&out)[DEC_TM / 16][4]) {
  constexpr int RM = DEC_TM / 16;
  float a2a[RM][8], a2b[RM][8];
  acc_zero(a2a);
  acc_zero(a2b);
  for (int c = 0; c < DEC_H1 / 128; ++c) {
    float a1[RM][8];
    acc_zero(a1);
    gemm_accum<DEC_TM, 128>(a1, featT, w0, DEC_H1, c * 128, Kp, wp);
    const float* bc = b0 + c * 128;
    acc_store_kmajor<DEC_TM, 128>(a1, hidT, 0, [&](int col, int, float v) { return fmaxf(v + __ldg(bc + col), 0.f); });
    const float* w1c = w1 + static_cast<size_t>(c) * 128 * DEC_H2;
    gemm_accum<DEC_TM, 128>(a2a, hidT, w1c, DEC_H2, 0, 128, wp);
    gemm_accum<DEC_TM, 128>(a2b, hidT, w1c, DEC_H2, 128, 128, wp);
  }
  acc_zero(out);
  acc_store_kmajor<DEC_TM, 128>(a2a, hidT, 0, [&](int col, int, float v) { return fmaxf(v + __ldg(b1 + col), 0.f); });
  gemm_accum<DEC_TM, 64>(out, hidT, w2, DEC_OUTC, 0, 128, wp);
  acc_store_kmajor<DEC_TM, 128>(a2b, hidT, 0,
                                [&](int col, int, float v) { return fmaxf(v + __ldg(b1 + 128 + col), 0.f); });
  gemm_accum<DEC_TM, 64>(out, hidT, w2 + static_cast<size_t>(128) * DEC_OUTC, DEC_OUTC, 0, 128, wp);
}

// smem (floats): featT [(F+Z) + 128][LD] (state rows behind the features) | hidT [128][LD] | xT [32][LD] |
//                wp [2*KC*128] | res [TM][64]
__global__ void __launch_bounds__(GN_THREADS)
decoder_block_kernel(DecArgs a) {
  constexpr int TM = DEC_TM, LD = DEC_LD, RM = TM / 16;
  extern __shared__ __align__(16) float smem[];
  const int Kf = a.F + a.Z;
  float* featT = smem;
  float* hT = featT + static_cast<size_t>(Kf) * LD;            // GRU state = rows Kf.. of the feature tile
  float* hidT = hT + DEC_GATE * LD;
  float* xT = hidT + 128 * LD;
  float* wp = xT + DEC_CONV * LD;
  float* res = wp + 2 * KC * 128;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int tp2 = 2 * a.Tp, tf2 = 2 * a.Tf;
  const float* br = a.w.gru_b;
  const float* bz = br + DEC_GATE;
  const float* bin = bz + DEC_GATE;
  const float* bhn = bin + DEC_GATE;
  const long long ntiles = (a.R + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    const int nrows = static_cast<int>(min(static_cast<long long>(TM), a.R - row0));
    load_tile_kmajor<TM>(featT, a.past_feature + static_cast<size_t>(row0) * a.F, a.F, nrows, a.F, a.F, 1.f, false);
    load_tile_kmajor<TM>(featT + static_cast<size_t>(a.F) * LD, a.z + static_cast<size_t>(row0) * a.Z, a.Z, nrows,
                         a.Z, a.Z, 1.f, false);
    for (int i = tid; i < TM * tp2; i += GN_THREADS) {
      const int r = i / tp2, c = i - r * tp2;
      float v = 0.f;
      if (r < nrows) {
        const long long row = row0 + r;
        v = __ldg(a.past_traj + (row / a.S) * tp2 + c);
        if (!a.first) v -= a.x_hat[row * tp2 + c];
      }
      res[r * DEC_RES_LD + c] = v;
    }
    for (int i = tid; i < DEC_GATE * LD; i += GN_THREADS) hT[i] = 0.f;
    __syncthreads();

    // ---- GRU over the T_p steps ----
    for (int t = 0; t < a.Tp; ++t) {
      for (int i = tid; i < DEC_CONV * TM; i += GN_THREADS) {
        const int c = i / TM, r = i - c * TM;
        const float* wc = a.w.conv_w + c * 6;
        float acc = __ldg(a.w.conv_b + c);
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int tt = t + kk - 1;
          if (tt >= 0 && tt < a.Tp) {
            acc = fmaf(__ldg(wc + kk), res[r * DEC_RES_LD + 2 * tt], acc);
            acc = fmaf(__ldg(wc + 3 + kk), res[r * DEC_RES_LD + 2 * tt + 1], acc);
          }
        }
        xT[c * LD + r] = fmaxf(acc, 0.f);
      }
      // (gemm_accum synchronises before its first read of xT / hT)
      float g[RM][8], acc[RM][8];
      acc_zero(acc);                                              // r gate
      gemm_accum<TM, 128>(acc, xT, a.w.gru_wx, 3 * DEC_GATE, 0, DEC_CONV, wp);
      gemm_accum<TM, 128>(acc, hT, a.w.gru_wh, 3 * DEC_GATE, 0, DEC_STATE, wp);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int r = 0; r < RM; ++r) g[r][j] = dec_sigmoid(acc[r][j] + __ldg(br + tx + 16 * j));
      acc_zero(acc);                                              // W_hn h + b_hn
      gemm_accum<TM, 128>(acc, hT, a.w.gru_wh, 3 * DEC_GATE, 2 * DEC_GATE, DEC_STATE, wp);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int r = 0; r < RM; ++r) g[r][j] *= acc[r][j] + __ldg(bhn + tx + 16 * j);
      acc_zero(acc);                                              // W_in e + b_in
      gemm_accum<TM, 128>(acc, xT, a.w.gru_wx, 3 * DEC_GATE, 2 * DEC_GATE, DEC_CONV, wp);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int r = 0; r < RM; ++r) g[r][j] = tanhf(acc[r][j] + __ldg(bin + tx + 16 * j) + g[r][j]);
      acc_zero(acc);                                              // z gate
      gemm_accum<TM, 128>(acc, xT, a.w.gru_wx, 3 * DEC_GATE, DEC_GATE, DEC_CONV, wp);
      gemm_accum<TM, 128>(acc, hT, a.w.gru_wh, 3 * DEC_GATE, DEC_GATE, DEC_STATE, wp);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = tx + 16 * j;
#pragma unroll
        for (int r = 0; r < RM; ++r) {
          const float zg = dec_sigmoid(acc[r][j] + __ldg(bz + col));
          const float hold = hT[col * LD + ty * RM + r];
          acc[r][j] = (1.f - zg) * g[r][j] + zg * hold;
        }
      }
      __syncthreads();                                            // every thread has read its old state
      acc_store_kmajor<TM, 128>(acc, hT, 0, [](int, int, float v) { return v; });
      __syncthreads();
    }

    // ---- decoder_x: x_hat, reconstruction ----
    {
      float o[RM][4];
      dec_mlp3(featT, a.Kp, hidT, wp, a.w.x_w0, a.w.x_b0, a.w.x_w1, a.w.x_b1, a.w.x_w2, o);
      acc_foreach<TM, 64>(o, [&](int r, int c, float& v) {
        if (r < nrows && c < tp2) {
          const size_t at = static_cast<size_t>(row0 + r) * tp2 + c;
          const float xv = v + __ldg(a.w.x_b2 + c);
          a.x_hat[at] = xv;
          a.recover[at] = a.first ? xv : a.recover[at] + xv;
        }
      });
    }
    // ---- decoder_y: prediction (+ cur_location after the last block) ----
    {
      float o[RM][4];
      dec_mlp3(featT, a.Kp, hidT, wp, a.w.y_w0, a.w.y_b0, a.w.y_w1, a.w.y_b1, a.w.y_w2, o);
      acc_foreach<TM, 64>(o, [&](int r, int c, float& v) {
        if (r < nrows && c < tf2) {
          const long long row = row0 + r;
          const size_t at = static_cast<size_t>(row) * tf2 + c;
          const float yv = v + __ldg(a.w.y_b2 + c);
          float p = a.first ? yv : a.out_seq[at] + yv;
          if (a.last) p += __ldg(a.cur_location + (row / a.S) * 2 + (c & 1));
          a.out_seq[at] = p;
        }
      });
    }
    __syncthreads();
  }
}

static size_t decoder_smem_bytes(int F, int Z) {
  return (static_cast<size_t>(F + Z + DEC_GATE + 128 + DEC_CONV) * DEC_LD + 2 * KC * 128 + DEC_TM * DEC_RES_LD) *
         sizeof(float);
}

}  // namespace gn

extern "C" size_t gn_decoder_workspace_bytes(int64_t A, int32_t S, int32_t Tp) {
  if (A <= 0 || S <= 0 || Tp <= 0) return 0;
  return static_cast<size_t>(A) * S * 2 * Tp * sizeof(float);
}

extern "C" int gn_decoder_fwd(const gn_decoder_weights* blocks, int32_t num_blocks, const float* past_feature,
                              const float* z, const float* past_traj, const float* cur_location, int64_t A,
                              int32_t S, int32_t F, int32_t Z, int32_t Tp, int32_t Tf, float* out_seq,
                              float* recover, void* workspace, size_t workspace_bytes, gn_stream_t stream) {
  using namespace gn;
  if (!blocks || !past_feature || !z || !past_traj || !cur_location || !out_seq || !recover || !workspace)
    return GN_E_NULL;
  if (num_blocks < 1 || A < 0 || S < 1 || Tp < 1 || Tf < 1 || 2 * Tp > DEC_RES_LD || 2 * Tf > DEC_OUTC ||
      F < 4 || Z < 4 || (F & 3) || (Z & 3))
    return GN_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(past_feature) | reinterpret_cast<uintptr_t>(z)) & 15) return GN_E_ALIGN;
  const size_t smem = decoder_smem_bytes(F, Z);
  if (smem > 227 * 1024) return GN_E_SHAPE;
  if (workspace_bytes < gn_decoder_workspace_bytes(A, S, Tp)) return GN_E_WORKSPACE;
  if (A == 0) return GN_OK;
  for (int b = 0; b < num_blocks; ++b) {
    const gn_decoder_weights& w = blocks[b];
    if (!w.conv_w || !w.conv_b || !w.gru_wx || !w.gru_wh || !w.gru_b || !w.x_w0 || !w.x_b0 || !w.x_w1 || !w.x_b1 ||
        !w.x_w2 || !w.x_b2 || !w.y_w0 || !w.y_b0 || !w.y_w1 || !w.y_b1 || !w.y_w2 || !w.y_b2)
      return GN_E_NULL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaFuncSetAttribute(decoder_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  DecArgs a;
  a.past_feature = past_feature; a.z = z; a.past_traj = past_traj; a.cur_location = cur_location;
  a.x_hat = static_cast<float*>(workspace); a.out_seq = out_seq; a.recover = recover;
  a.R = static_cast<long long>(A) * S;
  a.S = S; a.F = F; a.Z = Z; a.Tp = Tp; a.Tf = Tf;
  a.Kp = round_up(F + Z + DEC_STATE, KC);
  const long long ntiles = (a.R + DEC_TM - 1) / DEC_TM;
  const int grid = static_cast<int>(ntiles < GN_SM_COUNT ? ntiles : GN_SM_COUNT);
  for (int b = 0; b < num_blocks; ++b) {
    a.w = blocks[b];
    a.first = b == 0;
    a.last = b == num_blocks - 1;
    ProfScope prof("decoder_block", st);
    decoder_block_kernel<<<grid, GN_THREADS, smem, st>>>(a);
    GN_LAUNCH_CHECK();
  }
  return GN_OK;
}
