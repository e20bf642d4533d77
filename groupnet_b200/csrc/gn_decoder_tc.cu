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
//   decoder_mlp_fused  decoder_x and decoder_y (feat -> 512 -> 256 -> 2 T) of a 128-row tile in ONE role-split kernel:
//                      the hidden activations go from tensor memory to shared memory and back into the next MMA, the
//                      weights arrive as a host-packed stream of 16 KB stages (feature width <= 384, T <= 16);
//   tc_linear x 8      other shapes: the same Linears as row-tile GEMMs (gn_tc_linear.cu), the first Linears of both
//                      MLPs as one K = F + Z + 96 -> 1024 contraction (4 launches of N = 256), 512 -> 256 per MLP,
//                      256 -> 2 T (zero-padded to a multiple of 16) per MLP; bf16 activations between them.
//   decoder_finish     x_hat, reconstruction += x_hat, out_seq += y_hat (+ cur_location after the last block).
#include <cstdlib>
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

// ---------------------------------------------------------------------------------------------------------------
// decoder_x / decoder_y FUSED: feat -> 512 -> 256 -> 2 T for both MLPs of a 128-row tile in one kernel; the hidden
// activations never leave the SM (the row-tile GEMMs spent their time moving hid1 / hid2 and re-streaming weights).
//
//   warps 0-7   drain (warp w: TMEM lane quarter w & 3, column half w >> 2): D1 hidden-1 chunk (128 columns) -> bias, ReLU,
//               bf16 -> the A2 operand in shared memory; D2 hidden-2 (256 columns) -> the A3 operand; D3 the 32 output
//               columns -> + bias -> fp32 rows in HBM
//   warp 8      producer: the tile's feature rows (cp.async, once per tile for both MLPs) and the host-packed weight
//               stream — 16 KB stages in the issuer's consumption order, one bulk copy each — through a ring of 3
//   warp 9      issuer, per MLP:  G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) G2(3) G3
//                 G1(c)  acc1[c & 1] (128 cols) = feat (K = Kf) x W0[c]          K / 64 stages
//                 G2(c)  acc2 (256 cols)      += A2[c & 1] (K = 128) x W1[:, c]  4 stages (2 K halves x 2 N halves)
//                 G3     acc1[0] (32 cols)     = A3 (K = 256) x W2               4 stages
//               so D1(c) runs under G1(c + 1), and the second Linear consumes hidden-1 chunk by chunk
// TMEM: acc1[0] 0..127 | acc1[1] 128..255 | acc2 256..511.  A3 occupies both A2 buffers (all of G2 has completed).
namespace dfu {
constexpr int NST = 3, DRAIN_WARPS = 8, THREADS = (DRAIN_WARPS + 2) * 32;
constexpr int KF_MAX = 384;
constexpr uint32_t STG_BYTES = 16384;
constexpr uint32_t OFF_FEAT = 0;                              // [Kf/8][128][8] bf16, <= 96 KB
constexpr uint32_t OFF_A2 = 128 * KF_MAX * 2;                 // 2 x 32 KB
constexpr uint32_t OFF_W = OFF_A2 + 2 * 32768;                // NST x 16 KB
constexpr uint32_t OFF_BIAS = OFF_W + NST * STG_BYTES;        // b0[1024] | b1[512] | b2[64]
constexpr uint32_t OFF_BAR = OFF_BIAS + 1600 * 4;
enum { B_WFULL = 0, B_WEMPTY = 3, B_FEATFULL = 6, B_FEATEMPTY, B_ACC1FULL, B_ACC1EMPTY = 10, B_A2FULL = 12, B_A2EMPTY = 14,
       B_ACC2FULL = 16, B_ACC2EMPTY, B_COUNT };
constexpr uint32_t SMEM_BYTES = OFF_BAR + B_COUNT * 8 + 16;
}  // namespace dfu

struct DecFusedArgs {
  const __nv_bfloat16* feat; int Kf;            // (R, Kf) bf16, Kf % 64 == 0, Kf <= 384
  const __nv_bfloat16* stream;                  // per MLP: 4 * (Kf / 64) + 16 + 4 stages of 8,192 bf16
  const float* bias;                            // b0[1024] | b1x[256] b1y[256] | b2x[32] b2y[32]
  float* ox; float* oy;                         // (R, 32) fp32 each
  long long R;
};

__global__ void __launch_bounds__(dfu::THREADS, 1)
decoder_mlp_fused_kernel(DecFusedArgs a) {
  using namespace dfu;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sfeat = smem + OFF_FEAT;
  float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk1 = a.Kf >> 6;                    // stages of one G1
  const int per_mlp = 4 * nk1 + 16 + 4;

  for (int i = tid; i < 1600; i += THREADS) sbias[i] = __ldg(a.bias + i);
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) {
    for (int i = 0; i < NST; ++i) { mbar_init(bar + B_WFULL + i, 1); mbar_init(bar + B_WEMPTY + i, 1); }
    mbar_init(bar + B_FEATFULL, 32); mbar_init(bar + B_FEATEMPTY, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar + B_ACC1FULL + b, 1); mbar_init(bar + B_ACC1EMPTY + b, DRAIN_WARPS * 32);
      mbar_init(bar + B_A2FULL + b, DRAIN_WARPS * 32); mbar_init(bar + B_A2EMPTY + b, 1);
    }
    mbar_init(bar + B_ACC2FULL, 1); mbar_init(bar + B_ACC2EMPTY, DRAIN_WARPS * 32);
  }
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;
  const long long ntiles = (a.R + 127) / 128;
  // every role walks the same static schedule and counts its own barrier phases: cnt[b] = completed waits on bar[b].
  // "full" barriers are waited with the phase's own parity, "empty" ones with the opposite (a fresh barrier passes)
  uint32_t cnt[B_COUNT];
#pragma unroll
  for (int i = 0; i < B_COUNT; ++i) cnt[i] = 0;
  auto wait_full = [&](int b) { mbar_wait(bar + b, cnt[b] & 1u); ++cnt[b]; };
  auto wait_empty = [&](int b) { mbar_wait(bar + b, (cnt[b] & 1u) ^ 1u); ++cnt[b]; };

  if (warp == DRAIN_WARPS) {
    // ---------------- producer ----------------
    auto load_feat = [&](long long tile) {
      wait_empty(B_FEATEMPTY);
      const long long row0 = tile * 128;
      const int nrows = static_cast<int>(min(128LL, a.R - row0));
      const int nk8 = a.Kf >> 3;
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        const int r = rb * 32 + lane;
        if (r < nrows) {
          const __nv_bfloat16* src = a.feat + static_cast<size_t>(row0 + r) * a.Kf;
          for (int k8 = 0; k8 < nk8; ++k8) cp_async16(sfeat + canon_off(r, k8, 128), src + 8 * k8);
        } else {
          for (int k8 = 0; k8 < nk8; ++k8)
            *reinterpret_cast<uint4*>(sfeat + canon_off(r, k8, 128)) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      cp_async_commit();
    };
    auto publish_feat = [&]() {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(bar + B_FEATFULL);
    };
    uint32_t g = 0;
    // the next tile's rows are requested once the issuer is past the last read of this tile's (second MLP, G1(3))
    const int feat_at = per_mlp + 4 * nk1 + 8 + 3;
    if (static_cast<long long>(blockIdx.x) < ntiles) { load_feat(blockIdx.x); publish_feat(); }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const bool has_next = tile + gridDim.x < ntiles;
      for (int i = 0; i < 2 * per_mlp; ++i, ++g) {
        if (i == feat_at && has_next) load_feat(tile + gridDim.x);
        const uint32_t st = g % NST;
        mbar_wait(bar + B_WEMPTY + st, ((g / NST) & 1u) ^ 1u);
        const int im = i % per_mlp;
        const uint32_t bytes = im >= per_mlp - 4 ? 4096u : STG_BYTES;       // the last Linear's stages: 32 rows x 64 K
        tcu::expect_tx(bar + B_WFULL + st, bytes);
        tcu::bulk_g2s(smem_u32(smem + OFF_W + st * STG_BYTES), a.stream + static_cast<size_t>(i) * (STG_BYTES / 2), bytes,
                      bar + B_WFULL + st);
      }
      if (has_next) publish_feat();
    }
  } else if (warp == DRAIN_WARPS + 1) {
    // ---------------- MMA issuer ----------------
    uint32_t g = 0;
    const uint32_t feat_addr = smem_u32(sfeat), a2_addr = smem_u32(smem + OFF_A2);
    auto stage_mma = [&](uint32_t tmem_d, uint32_t a_addr, int N, bool accumulate) {   // one 64-wide K stage
      const uint32_t st = g % NST;
      mbar_wait(bar + B_WFULL + st, (g / NST) & 1u);
      fence_after_thread_sync();
      if (elect_one()) {
        issue_gemm(tmem_d, a_addr, smem_u32(smem + OFF_W + st * STG_BYTES), N, 64, accumulate);
        mma_commit(bar + B_WEMPTY + st);
      }
      __syncwarp();
      ++g;
    };
    auto commit = [&](int b) {
      if (elect_one()) mma_commit(bar + b);
      __syncwarp();
    };
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      wait_full(B_FEATFULL);
      fence_after_thread_sync();
      for (int m = 0; m < 2; ++m) {
        auto G1 = [&](int c) {
          const int b = c & 1;
          wait_empty(B_ACC1EMPTY + b);
          fence_after_thread_sync();
          for (int ks = 0; ks < nk1; ++ks) stage_mma(tmem + b * 128u, feat_addr + ks * (8u * 2048u), 128, ks > 0);
          commit(B_ACC1FULL + b);
        };
        auto G2 = [&](int c) {
          const int b = c & 1;
          wait_full(B_A2FULL + b);
          if (c == 0) wait_empty(B_ACC2EMPTY);
          fence_after_thread_sync();
          for (int kh = 0; kh < 2; ++kh)
            for (int nh = 0; nh < 2; ++nh)
              stage_mma(tmem + 256u + nh * 128u, a2_addr + b * 32768u + kh * (8u * 2048u), 128, c > 0 || kh > 0);
          commit(B_A2EMPTY + b);
        };
        G1(0); G1(1); G2(0); G1(2); G2(1); G1(3);
        if (m == 1) commit(B_FEATEMPTY);
        G2(2); G2(3);
        commit(B_ACC2FULL);
        // G3: A3 (K = 256) lives in both A2 buffers
        wait_full(B_A2FULL + 0); wait_full(B_A2FULL + 1);
        wait_empty(B_ACC1EMPTY + 0);
        fence_after_thread_sync();
        for (int ks = 0; ks < 4; ++ks) stage_mma(tmem, a2_addr + ks * (8u * 2048u), 32, ks > 0);
        commit(B_ACC1FULL + 0);
        commit(B_A2EMPTY + 0); commit(B_A2EMPTY + 1);
      }
    }
  } else {
    // ---------------- drain ----------------
    const int q = warp & 3, ch = warp >> 2, row = q * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
    unsigned char* sA2 = smem + OFF_A2;
    // `ncols` accumulator columns starting at TMEM column tcol (this warp's half) -> + bias, ReLU -> bf16 k-groups kg0.. of dst
    auto to_operand = [&](uint32_t tcol, int ncols, const float* bias, unsigned char* dst, int kg0) {
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        float v[16];
        tmem_ld16(trow + tcol + c0, v);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) pk[j >> 1] = pack_bf16_relu(v[j] + bias[c0 + j], v[j + 1] + bias[c0 + j + 1]);
        *reinterpret_cast<uint4*>(dst + canon_off(row, kg0 + (c0 >> 3), 128)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(dst + canon_off(row, kg0 + (c0 >> 3) + 1, 128)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    };
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long grow = tile * 128 + row;
      for (int m = 0; m < 2; ++m) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {                    // D1(c): hidden-1 chunk c -> A2[c & 1]
          const int b = c & 1;
          wait_full(B_ACC1FULL + b);
          wait_empty(B_A2EMPTY + b);
          fence_after_thread_sync();
          to_operand(b * 128u + ch * 64u, 64, sbias + m * 512 + c * 128 + ch * 64, sA2 + b * 32768, ch * 8);
          fence_proxy_async_smem();
          fence_before_thread_sync();
          mbar_arrive(bar + B_A2FULL + b);
          mbar_arrive(bar + B_ACC1EMPTY + b);
        }
        // D2: hidden-2 (256 columns) -> A3 over both A2 buffers
        wait_full(B_ACC2FULL);
        wait_empty(B_A2EMPTY + 0); wait_empty(B_A2EMPTY + 1);
        fence_after_thread_sync();
        to_operand(256u + ch * 128u, 128, sbias + 1024 + m * 256 + ch * 128, sA2, ch * 16);
        fence_proxy_async_smem();
        fence_before_thread_sync();
        mbar_arrive(bar + B_A2FULL + 0); mbar_arrive(bar + B_A2FULL + 1);
        mbar_arrive(bar + B_ACC2EMPTY);
        // D3: the 32 output columns (this warp: 16) -> + bias -> HBM
        wait_full(B_ACC1FULL + 0);
        fence_after_thread_sync();
        {
          float v[16];
          tmem_ld16(trow + ch * 16u, v);
          const float* b2 = sbias + 1536 + m * 32 + ch * 16;
          if (grow < a.R) {
            float* dst = (m == 0 ? a.ox : a.oy) + static_cast<size_t>(grow) * 32 + ch * 16;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j] + b2[j], v[j + 1] + b2[j + 1], v[j + 2] + b2[j + 2], v[j + 3] + b2[j + 3]);
          }
        }
        fence_before_thread_sync();
        mbar_arrive(bar + B_ACC1EMPTY + 0);
      }
    }
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

// shapes the fused MLP kernel covers (GN_DECODER_MLP=rowtile forces the row-tile GEMMs: A/B measurements)
static bool dec_mlp_fused_shape(int F, int Z, int Tp, int Tf) {
  static const bool force_rowtile = [] { const char* v = getenv("GN_DECODER_MLP"); return v && v[0] == 'r'; }();
  const int Kf = F + Z + dtc::STATE;
  return !force_rowtile && Kf <= dfu::KF_MAX && (Kf & 63) == 0 && Tp <= 16 && Tf <= 16;
}

struct DecTcLayout { size_t x_hat, feat, hid1, hid2, ox, oy, total; };
static DecTcLayout dec_tc_layout(long long R, int F, int Z, int Tp, int Tf) {
  DecTcLayout l;
  size_t off = 0;
  const bool hidden = !dec_mlp_fused_shape(F, Z, Tp, Tf);   // the fused kernel keeps both hidden layers on chip
  l.x_hat = off; off += align256(static_cast<size_t>(R) * 2 * Tp * 4);
  l.feat = off;  off += align256(static_cast<size_t>(R) * (F + Z + dtc::STATE) * 2);
  l.hid1 = off;  off += hidden ? align256(static_cast<size_t>(R) * 1024 * 2) : 0;
  l.hid2 = off;  off += hidden ? align256(static_cast<size_t>(R) * 512 * 2) : 0;
  l.ox = off;    off += align256(static_cast<size_t>(R) * ((2 * Tp + 31) & ~31) * 4);   // >= 32 columns: the fused MLP kernel's rows
  l.oy = off;    off += align256(static_cast<size_t>(R) * ((2 * Tf + 31) & ~31) * 4);
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
    // the fused MLP kernel (hidden activations stay on chip) covers feature widths up to 384 and up to 16 time steps per
    // output; other shapes run the row-tile GEMMs.  GN_DECODER_MLP=rowtile forces the latter (A/B measurements).
    const bool fused = dec_mlp_fused_shape(F, Z, Tp, Tf);
    if (fused && (!w.mlp_stream || !w.mlp_bias)) return GN_E_NULL;   // the workspace holds no hidden-layer buffers at this shape
    int ldx = n3x, ldy = n3y;
    if (fused) {
      e = cudaFuncSetAttribute(decoder_mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(dfu::SMEM_BYTES));
      if (e != cudaSuccess) return static_cast<int>(e);
      DecFusedArgs fa;
      fa.feat = feat; fa.Kf = Kf; fa.stream = static_cast<const __nv_bfloat16*>(w.mlp_stream); fa.bias = w.mlp_bias;
      fa.ox = ox; fa.oy = oy; fa.R = R;
      {
        ProfScope prof("decoder_mlp_fused", st);
        decoder_mlp_fused_kernel<<<grid, dfu::THREADS, dfu::SMEM_BYTES, st>>>(fa);
      }
      GN_LAUNCH_CHECK();
      ldx = ldy = 32;
    } else {
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
    }
    DecFinishArgs f;
    f.ox = ox; f.oy = oy; f.ldx = ldx; f.ldy = ldy; f.cur_location = cur_location;
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
