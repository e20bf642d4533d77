// Trajectory decoder, bf16 tensor-core path (tcgen05 / TMEM), 2e-2 parity.
//
// model/GroupNet_nba.py:48-79 (DecomposeBlock.forward) and :461-505 (Decoder.forward); the math is the one
// gn_decoder_simt.cu states.  Per DecomposeBlock:
//
//   decoder_gru_tc     tile = 128 rows (scene-agent x sample), 512 threads = 4 per row.  The GRU step is ONE
//                      K = 128 contraction per time step: A = [e_t (32) | h (96)] as a bf16 operand in shared
//                      memory, B = the resident gate matrix [r | z | n_x | n_h] (384 x 128, zero blocks where a
//                      gate does not see e or h; two N = 192 MMAs chains), accumulators in 384 TMEM columns; the
//                      gate math (sigmoid, tanh, blend) runs in the drain on the row's four threads, the state stays
//                      fp32 in their registers and only its bf16 operand copy goes back to shared memory.  conv1d +
//                      ReLU of the residual is evaluated per step while the operand is built.  Writes the bf16
//                      feature row [past_feature | z | state] the MLPs read.
//   tc_linear x 8      decoder_x / decoder_y as row-tile GEMMs (gn_tc_linear.cu): the first Linears of both MLPs
//                      as one K = F + Z + 96 -> 1024 contraction (4 launches of N = 256), 512 -> 256 per MLP,
//                      256 -> 2 T (zero-padded to a multiple of 16) per MLP; bf16 activations between them.
//   decoder_finish     x_hat, reconstruction += x_hat, out_seq += y_hat (+ cur_location after the last block).
#include "gn_tc.cuh"
#include "gn_stage.h"

#define GN_TRY(expr) do { int rc__ = (expr); if (rc__ != GN_OK) return rc__; } while (0)

namespace gn {

namespace dtc {
constexpr int TM = 128, THREADS = 512;
constexpr int CONV = 32, STATE = 96, KG = 128;            // K of the gate contraction: e_t | h
constexpr int NG = 192;                                   // one gate operand: two gates of 96 rows
constexpr int RES_LD = 65;                                // 2 * Tp <= 64 residual values per row (+1: bank spread)
constexpr uint32_t W_BYTES = NG * KG * 2;                 // 48 KB per gate operand
constexpr uint32_t A_BYTES = TM * KG * 2;                 // 32 KB
constexpr uint32_t OFF_W0 = 0, OFF_W1 = W_BYTES, OFF_A = 2 * W_BYTES;
constexpr uint32_t OFF_RES = OFF_A + A_BYTES;
constexpr uint32_t OFF_CONST = OFF_RES + TM * RES_LD * 4;  // gate biases [4][96] | conv_w [32*6] | conv_b [32]
constexpr uint32_t CONST_FLOATS = 4 * STATE + CONV * 6 + CONV;
constexpr uint32_t OFF_BAR = (OFF_CONST + CONST_FLOATS * 4 + 15) & ~15u;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 32;
}  // namespace dtc

struct DecGruArgs {
  const __nv_bfloat16* gru_w;          // two canonical [192 x 128] operands: (r | z), (n_x | n_h)
  const float* gru_b;                  // (4,128): b_r | b_z | b_in | b_hn
  const float* conv_w; const float* conv_b;
  const float* past_feature; const float* z; const float* past_traj; const float* x_hat;
  __nv_bfloat16* feat;                 // (R, F + Z + 96)
  long long R;
  int S, F, Z, Tp, first;
};

// MUFU-based gate functions (ex2 / rcp / tanh approximations: ~1e-6 .. 5e-4 absolute, far inside the bf16 operand rounding)
__device__ __forceinline__ float fast_tanh(float v);
// sigmoid(v) = 0.5 tanh(v / 2) + 0.5: ONE MUFU op instead of ex2 + rcp (the gate math is MUFU-bound: 5 -> 3 ops per column)
__device__ __forceinline__ float fast_sigmoid(float v) { return fmaf(0.5f, fast_tanh(0.5f * v), 0.5f); }
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ float fast_tanh(float v) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(v));
  return y;
}

__global__ void __launch_bounds__(dtc::THREADS, 1)
decoder_gru_tc_kernel(DecGruArgs a) {
  using namespace dtc;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem + OFF_A;
  float* res = reinterpret_cast<float*>(smem + OFF_RES);
  float* sgb = reinterpret_cast<float*>(smem + OFF_CONST);          // [4][96]
  float* scw = sgb + 4 * STATE;                                     // [32][2][3]
  float* scb = scw + CONV * 6;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hf = warp >> 2;                                         // which quarter of the row's columns
  const int row = (warp & 3) * 32 + lane;                           // tile row = TMEM lane
  const int Kf = a.F + a.Z + STATE, FZ = a.F + a.Z, tp2 = 2 * a.Tp;

  // resident operands and constants
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.gru_w);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_W0);
    for (int i = tid; i < static_cast<int>(2 * W_BYTES / 16); i += THREADS) dst[i] = __ldg(src + i);
    for (int i = tid; i < 4 * STATE; i += THREADS) sgb[i] = __ldg(a.gru_b + (i / STATE) * 128 + (i % STATE));
    for (int i = tid; i < CONV * 6; i += THREADS) scw[i] = __ldg(a.conv_w + i);
    for (int i = tid; i < CONV; i += THREADS) scb[i] = __ldg(a.conv_b + i);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) mbar_init(mbar, 1);
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t sA_addr = smem_u32(sA), sW0_addr = smem_u32(smem + OFF_W0), sW1_addr = smem_u32(smem + OFF_W1);
  uint32_t phase = 0;
  const long long ntiles = (a.R + TM - 1) / TM;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    const int nrows = static_cast<int>(min(static_cast<long long>(TM), a.R - row0));
    // residual x_true - x_hat of the tile's rows
    for (int i = tid; i < TM * tp2; i += THREADS) {
      const int r = i / tp2, c = i - r * tp2;
      float v = 0.f;
      if (r < nrows) {
        const long long g = row0 + r;
        v = __ldg(a.past_traj + (g / a.S) * tp2 + c);
        if (!a.first) v -= a.x_hat[g * tp2 + c];
      }
      res[r * RES_LD + c] = v;
    }
    // [past_feature | z] -> bf16 feature rows (the same for every block: written by the first one)
    if (a.first) {
      const int f8 = a.F >> 3, fz8 = FZ >> 3;
      for (int i = tid; i < nrows * fz8; i += THREADS) {
        const int r = i / fz8, k8 = i - r * fz8;
        const long long g = row0 + r;
        const float* src = k8 < f8 ? a.past_feature + g * a.F + 8 * k8 : a.z + g * a.Z + 8 * (k8 - f8);
        const float4 x = ldg_f4(src), y = ldg_f4(src + 4);
        *reinterpret_cast<uint4*>(a.feat + g * Kf + 8 * k8) =
            make_uint4(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w), pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
      }
    }
    // h_0 = 0: the state part of the operand (this thread's k-groups 4 + 3 hf ..)
#pragma unroll
    for (int g8 = 0; g8 < 3; ++g8)
      *reinterpret_cast<uint4*>(sA + canon_off(row, 4 + 3 * hf + g8, TM)) = make_uint4(0u, 0u, 0u, 0u);
    float hreg[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) hreg[j] = 0.f;
    __syncthreads();                                   // res complete

    for (int t = 0; t < a.Tp; ++t) {
      // e_t = relu(conv1d(res)) for this thread's 8 of the 32 channels -> k-group hf
      {
        float xin[6];                                  // res[t-1], res[t], res[t+1] x (x, y); zero padding
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int tt = t + kk - 1;
          const bool in = tt >= 0 && tt < a.Tp;
          xin[kk] = in ? res[row * RES_LD + 2 * tt] : 0.f;
          xin[3 + kk] = in ? res[row * RES_LD + 2 * tt + 1] : 0.f;
        }
        uint32_t pk[4];
#pragma unroll
        for (int c2 = 0; c2 < 4; ++c2) {
          float e[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = 8 * hf + 2 * c2 + u;
            const float* wc = scw + c * 6;
            float acc = scb[c];
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
              acc = fmaf(wc[kk], xin[kk], acc);
              acc = fmaf(wc[3 + kk], xin[3 + kk], acc);
            }
            e[u] = acc;
          }
          pk[c2] = pack_bf16_relu(e[0], e[1]);
        }
        *reinterpret_cast<uint4*>(sA + canon_off(row, hf, TM)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
      fence_before_thread_sync();
      __syncthreads();                                 // operand complete; previous drain done with the accumulators
      if (warp == 0) {
        fence_after_thread_sync();
        if (elect_one()) {
          issue_gemm(tmem, sA_addr, sW0_addr, NG, KG, false);          // r | z       -> columns   0..191
          issue_gemm(tmem + NG, sA_addr, sW1_addr, NG, KG, false);     // n_x | n_h   -> columns 192..383
          mma_commit(mbar);
        }
        __syncwarp();
      }
      mbar_wait(mbar, phase); phase ^= 1;
      fence_after_thread_sync();
      // gate math on this thread's 24 state columns [24 hf, 24 hf + 24), 8 (= one operand k-group) at a time
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int j0 = 24 * hf + 8 * c;
        uint32_t vr[8], vz[8], vx[8], vh[8];
        tmem_ld8_nowait(tmem_row + j0, vr);
        tmem_ld8_nowait(tmem_row + STATE + j0, vz);
        tmem_ld8_nowait(tmem_row + 2 * STATE + j0, vx);
        tmem_ld8_nowait(tmem_row + 3 * STATE + j0, vh);
        tmem_ld_wait();
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float hn[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = j0 + j + u;
            const float rg = fast_sigmoid(__uint_as_float(vr[j + u]) + sgb[col]);
            const float zg = fast_sigmoid(__uint_as_float(vz[j + u]) + sgb[STATE + col]);
            const float ng = fast_tanh(__uint_as_float(vx[j + u]) + sgb[2 * STATE + col] +
                                       rg * (__uint_as_float(vh[j + u]) + sgb[3 * STATE + col]));
            const float hold = hreg[8 * c + j + u];
            hn[u] = fmaf(zg, hold - ng, ng);           // (1 - z) n + z h
            hreg[8 * c + j + u] = hn[u];
          }
          pk[j >> 1] = pack_bf16(hn[0], hn[1]);
        }
        // the MMAs of this step have completed (mbarrier): the operand buffer is free for h_t
        *reinterpret_cast<uint4*>(sA + canon_off(row, 4 + 3 * hf + c, TM)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    // state -> feature columns [F + Z + 24 hf, + 24)
    if (row < nrows) {
      __nv_bfloat16* dst = a.feat + (row0 + row) * Kf + FZ + 24 * hf;
#pragma unroll
      for (int g8 = 0; g8 < 3; ++g8)
        *reinterpret_cast<uint4*>(dst + 8 * g8) =
            make_uint4(pack_bf16(hreg[8 * g8], hreg[8 * g8 + 1]), pack_bf16(hreg[8 * g8 + 2], hreg[8 * g8 + 3]),
                       pack_bf16(hreg[8 * g8 + 4], hreg[8 * g8 + 5]), pack_bf16(hreg[8 * g8 + 6], hreg[8 * g8 + 7]));
    }
    fence_before_thread_sync();
    __syncthreads();                                   // res / operand / accumulators are rewritten by the next tile
  }

  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_thread_sync();
    tmem_dealloc(tmem, 512);
  }
}

struct DecFinishArgs {
  const float* ox; const float* oy; int ldx, ldy;      // last Linears' outputs (R, ldx) / (R, ldy), biases included
  const float* cur_location;
  float* x_hat; float* recover; float* out_seq;
  long long R;
  int S, tp2, tf2, first, last;
};

__global__ void __launch_bounds__(256)
decoder_finish_kernel(DecFinishArgs a) {
  const int w = a.tp2 + a.tf2;
  const long long total = a.R * w;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / w;
    const int c = static_cast<int>(i - row * w);
    if (c < a.tp2) {
      const float xv = a.ox[row * a.ldx + c];
      const long long at = row * a.tp2 + c;
      a.x_hat[at] = xv;
      a.recover[at] = a.first ? xv : a.recover[at] + xv;
    } else {
      const int cy = c - a.tp2;
      const long long at = row * a.tf2 + cy;
      float p = a.oy[row * a.ldy + cy];
      if (!a.first) p += a.out_seq[at];
      if (a.last) p += __ldg(a.cur_location + (row / a.S) * 2 + (cy & 1));
      a.out_seq[at] = p;
    }
  }
}

static inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
static inline int pad16(int v) { return (v + 15) & ~15; }

struct DecTcLayout { size_t x_hat, feat, hid1, hid2, ox, oy, total; };
static DecTcLayout dec_tc_layout(long long R, int F, int Z, int Tp, int Tf) {
  DecTcLayout l;
  size_t off = 0;
  l.x_hat = off; off += align256(static_cast<size_t>(R) * 2 * Tp * 4);
  l.feat = off;  off += align256(static_cast<size_t>(R) * (F + Z + dtc::STATE) * 2);
  l.hid1 = off;  off += align256(static_cast<size_t>(R) * 1024 * 2);
  l.hid2 = off;  off += align256(static_cast<size_t>(R) * 512 * 2);
  l.ox = off;    off += align256(static_cast<size_t>(R) * pad16(2 * Tp) * 4);
  l.oy = off;    off += align256(static_cast<size_t>(R) * pad16(2 * Tf) * 4);
  l.total = off;
  return l;
}

}  // namespace gn

extern "C" size_t gn_decoder_tc_workspace_bytes(int64_t A, int32_t S, int32_t F, int32_t Z, int32_t Tp, int32_t Tf) {
  if (A <= 0 || S <= 0 || F <= 0 || Z <= 0 || Tp <= 0 || Tf <= 0) return 0;
  return gn::dec_tc_layout(static_cast<long long>(A) * S, F, Z, Tp, Tf).total;
}

extern "C" int gn_decoder_fwd_tc(const gn_decoder_tc_weights* blocks, int32_t num_blocks, const float* past_feature,
                                 const float* z, const float* past_traj, const float* cur_location, int64_t A,
                                 int32_t S, int32_t F, int32_t Z, int32_t Tp, int32_t Tf, float* out_seq,
                                 float* recover, void* workspace, size_t workspace_bytes, gn_stream_t stream) {
  using namespace gn;
  if (!blocks || !past_feature || !z || !past_traj || !cur_location || !out_seq || !recover || !workspace)
    return GN_E_NULL;
  if (num_blocks < 1 || A < 0 || S < 1 || Tp < 1 || Tf < 1 || Tp > 32 || Tf > 32 || F < 8 || Z < 8 || (F & 7) || (Z & 7) ||
      ((F + Z) & 15))
    return GN_E_SHAPE;
  if ((reinterpret_cast<uintptr_t>(past_feature) | reinterpret_cast<uintptr_t>(z) |
       reinterpret_cast<uintptr_t>(workspace)) & 15)
    return GN_E_ALIGN;
  if (A == 0) return GN_OK;
  const long long R = static_cast<long long>(A) * S;
  const DecTcLayout l = dec_tc_layout(R, F, Z, Tp, Tf);
  if (workspace_bytes < l.total) return GN_E_WORKSPACE;
  for (int b = 0; b < num_blocks; ++b) {
    const gn_decoder_tc_weights& w = blocks[b];
    if (!w.conv_w || !w.conv_b || !w.gru_w || !w.gru_b || !w.w0 || !w.b0 || !w.x_w1 || !w.x_b1 || !w.x_w2 ||
        !w.x_b2 || !w.y_w1 || !w.y_b1 || !w.y_w2 || !w.y_b2)
      return GN_E_NULL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float* x_hat = reinterpret_cast<float*>(ws + l.x_hat);
  __nv_bfloat16* feat = reinterpret_cast<__nv_bfloat16*>(ws + l.feat);
  __nv_bfloat16* hid1 = reinterpret_cast<__nv_bfloat16*>(ws + l.hid1);
  __nv_bfloat16* hid2 = reinterpret_cast<__nv_bfloat16*>(ws + l.hid2);
  float* ox = reinterpret_cast<float*>(ws + l.ox);
  float* oy = reinterpret_cast<float*>(ws + l.oy);
  const int Kf = F + Z + dtc::STATE, n3x = pad16(2 * Tp), n3y = pad16(2 * Tf);
  cudaError_t e = cudaFuncSetAttribute(decoder_gru_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(dtc::SMEM_BYTES));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int sms = sm_count();
  const long long ntiles = (R + dtc::TM - 1) / dtc::TM;
  const int grid = static_cast<int>(ntiles < sms ? ntiles : sms);

  auto linear = [&](const void* A0, long long lda, int K, const __nv_bfloat16* W, int Ntot, int n0, int N,
                    const float* bias, int relu, void* out, int out_f32, long long ldo, int col0,
                    const char* name) -> int {
    TcLinArgs t{};
    t.A0 = A0; t.a0_is_f32 = 0; t.lda0 = lda; t.K0 = K;
    t.A1 = nullptr; t.lda1 = 0; t.K1 = 0; t.a_div = 0.f;
    t.W = W; t.Ntot = Ntot; t.n0 = n0; t.N = N; t.bias = bias; t.relu = relu;
    t.rowscale = nullptr; t.rs_ld = 0; t.rs_shift = 0; t.bias_mat = nullptr; t.bm_T = 0; t.bm_ld = 0;
    t.out = out; t.out_is_f32 = out_f32; t.ldo = ldo; t.out_col0 = col0; t.out_div = 0.f; t.R = R;
    return launch_tc_linear(t, name, st);
  };

  for (int b = 0; b < num_blocks; ++b) {
    const gn_decoder_tc_weights& w = blocks[b];
    DecGruArgs g;
    g.gru_w = static_cast<const __nv_bfloat16*>(w.gru_w); g.gru_b = w.gru_b; g.conv_w = w.conv_w; g.conv_b = w.conv_b;
    g.past_feature = past_feature; g.z = z; g.past_traj = past_traj; g.x_hat = x_hat; g.feat = feat;
    g.R = R; g.S = S; g.F = F; g.Z = Z; g.Tp = Tp; g.first = b == 0;
    {
      ProfScope prof("decoder_gru_tc", st);
      decoder_gru_tc_kernel<<<grid, dtc::THREADS, dtc::SMEM_BYTES, st>>>(g);
    }
    GN_LAUNCH_CHECK();
    const __nv_bfloat16* w0 = static_cast<const __nv_bfloat16*>(w.w0);
    for (int c = 0; c < 4; ++c)
      GN_TRY(linear(feat, Kf, Kf, w0, 1024, 256 * c, 256, w.b0, 1, hid1, 0, 1024, 256 * c, "decoder_mlp0_tc"));
    GN_TRY(linear(hid1, 1024, 512, static_cast<const __nv_bfloat16*>(w.x_w1), 256, 0, 256, w.x_b1, 1, hid2, 0, 512, 0,
                  "decoder_mlp1_tc"));
    GN_TRY(linear(hid1 + 512, 1024, 512, static_cast<const __nv_bfloat16*>(w.y_w1), 256, 0, 256, w.y_b1, 1, hid2, 0,
                  512, 256, "decoder_mlp1_tc"));
    GN_TRY(linear(hid2, 512, 256, static_cast<const __nv_bfloat16*>(w.x_w2), n3x, 0, n3x, w.x_b2, 0, ox, 1, n3x, 0,
                  "decoder_mlp2_tc"));
    GN_TRY(linear(hid2 + 256, 512, 256, static_cast<const __nv_bfloat16*>(w.y_w2), n3y, 0, n3y, w.y_b2, 0, oy, 1, n3y,
                  0, "decoder_mlp2_tc"));
    DecFinishArgs f;
    f.ox = ox; f.oy = oy; f.ldx = n3x; f.ldy = n3y; f.cur_location = cur_location;
    f.x_hat = x_hat; f.recover = recover; f.out_seq = out_seq; f.R = R; f.S = S;
    f.tp2 = 2 * Tp; f.tf2 = 2 * Tf; f.first = b == 0; f.last = b == num_blocks - 1;
    const long long total = R * (f.tp2 + f.tf2);
    const long long want = (total + 255) / 256;
    const int fgrid = static_cast<int>(want < 8LL * sms ? want : 8LL * sms);
    {
      ProfScope prof("decoder_finish", st);
      decoder_finish_kernel<<<fgrid, 256, 0, st>>>(f);
    }
    GN_LAUNCH_CHECK();
  }
  return GN_OK;
}
