// Fused edge2node for the pairwise layer (bf16 tensor-core path).
//
// Reference (model/MS_HGNN_batch.py:259-268 with H = rel_rec + rel_send, :116-120):
//   edges_e = h_i + h_j;  ef_e = sum_t edge_feat[e,t] * (W1_t relu(W0_t edges_e + b0_t) + b1_t)
//   agg_n = sum_e H[e,n] ef_e
// Collapsed form: the first Linear acts per node, the second per node:
//   P'_n[t] = W0_t h_n + b0_t / 2                                        (GEMM 1, tensor core)
//   w[n][j][t] = edge_feat[(n,j),t] + edge_feat[(j,n),t]                  (symmetric, from smem)
//   G_n[t]  = sum_j w[n][j][t] * relu(P'_n[t] + P'_j[t])                  (SIMT, shared memory)
//   agg_n   = sum_t W1_t G_n[t] + (sum_j w[n][j][t]) b1_t                 (GEMM 2, tensor core, K = T*128)
//
// One persistent CTA per SM; a tile = SC = floor(128/N) whole scenes = SC*N <= 128 node rows.
// Per t: GEMM 1 of step t+1 runs on the tensor core while the 256 threads do the relu-sum of
// step t; P'/G never leave the SM (the unfused path moves 2 x 2.2 GB of P and G through HBM).
// W0_t / W1_t (16 KB each) are streamed from L2 with cp.async.
//
// Bound: the SIMT relu-sum (3 instructions per (n, j, t, c)): N * T * 128 * 3 thread-instructions
// per node row.  Algorithmic HBM bytes per scene: N*D*4 (h) + N*N*T*4 (edge_feat) + N*D*4 (agg).
#include <cuda_fp16.h>
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {

struct PairAggArgs {
  const float* h;            // (B*N, 64)
  const float* edge_feat;    // (B, N*N, T)
  const __nv_bfloat16* w0;   // canonical (T*128, 64): tc_agg_w0
  const __nv_bfloat16* w1;   // canonical (64, T*128): tc_agg_w1
  const float* b0;           // (T*128)
  const float* b1;           // (T, 64)
  float* agg;                // (B*N, 64)
  int B, N, T, SC;
};

namespace pagg {
constexpr int D = 64;
constexpr uint32_t OFF_A = 0;                          // h tile, bf16 canonical [128 x 64]      16 KB
constexpr int PLD = 136;                                // padded P' row (halfs): 272 B, rows 4 banks apart
constexpr uint32_t OFF_P = OFF_A + 128 * 64 * 2;       // P'_t fp16 [128][PLD]                   34 KB
constexpr uint32_t OFF_G = OFF_P + 128 * PLD * 2;      // G_t bf16 canonical [128 x 128]         32 KB
constexpr uint32_t OFF_W0 = OFF_G + 128 * 128 * 2;     // W0_t double buffer [128 x 64] x 2      32 KB
constexpr uint32_t OFF_W1 = OFF_W0 + 2 * 128 * 64 * 2; // W1_t [64 x 128]                        16 KB
constexpr uint32_t OFF_ONES = OFF_W1 + 64 * 128 * 2;   // ones operand                            4 KB
constexpr uint32_t OFF_BB = OFF_ONES + 128 * 32;       // b0_t/2 operand double buffer            8 KB
constexpr uint32_t OFF_S = OFF_BB + 2 * 128 * 32;      // S[128][16] fp32                         8 KB
constexpr uint32_t OFF_BAR = OFF_S + 128 * 16 * 4;     // mbarA, mbarB, tmem slot
constexpr uint32_t OFF_EF = OFF_BAR + 32;              // edge_feat of the tile: SC*N*N*T fp32
constexpr uint32_t TM_P = 0, TM_AGG = 256;             // TMEM columns: P' double buffer 2 x 128, agg 64
}  // namespace pagg

constexpr int PA_THREADS = 512;   // (tile row, 32-column quarter): 16 warps hide the latency of the SIMT relu-sum

__global__ void __launch_bounds__(PA_THREADS, 1)
pair_agg_tc_kernel(PairAggArgs a) {
  using namespace pagg;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = tid & 127, cq = tid >> 7;             // tile row (TMEM lane), 32-column quarter
  __half* sP = reinterpret_cast<__half*>(smem + OFF_P);
  float* sS = reinterpret_cast<float*>(smem + OFF_S);
  float* sEF = reinterpret_cast<float*>(smem + OFF_EF);
  uint64_t* mbarA = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* mbarB = mbarA + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
  const int N = a.N, T = a.T, SC = a.SC, E = N * N;
  const int NT = T * 128;

  build_ones_operand(smem + OFF_ONES, tid, PA_THREADS);
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 32) { mbar_init(mbarA, 1); mbar_init(mbarB, 1); }
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t sbase = smem_u32(smem);
  uint32_t phA = 0, phB = 0;

  // W0_t: rows [t*128, t*128+128) of the canonical (NT x 64) operand: 8 k-groups x 2 KB
  auto load_w0 = [&](int t, int buf) {
    for (int i = tid; i < 8 * 128; i += PA_THREADS) {
      const int k8 = i >> 7, n = i & 127;
      cp_async16(smem + OFF_W0 + buf * (128 * 64 * 2) + (k8 * 128 + n) * 16,
                 a.w0 + (static_cast<size_t>(k8) * NT + t * 128 + n) * 8);
    }
  };
  // W1_t: k-groups [t*16, t*16+16) of the canonical (64 x NT) operand: 16 KB contiguous
  auto load_w1 = [&](int t) {
    const __nv_bfloat16* src = a.w1 + static_cast<size_t>(t) * 16 * 64 * 8;
    for (int i = tid; i < 16 * 64; i += PA_THREADS) cp_async16(smem + OFF_W1 + i * 16, src + i * 8);
  };
  // b0_t / 2 as a [128 x 16] bias operand
  auto build_bb = [&](int t, int buf) {
    if (tid < 128) {
      const float v = 0.5f * __ldg(a.b0 + t * 128 + tid);
      __nv_bfloat16 hi = __float2bfloat16_rn(v);
      __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      const uint32_t w0 = static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&hi)) |
                          (static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&lo)) << 16);
      unsigned char* dst = smem + OFF_BB + buf * (128 * 32);
      *reinterpret_cast<uint4*>(dst + tid * 16) = make_uint4(w0, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(dst + 128 * 16 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  };
  auto issue_gemm1 = [&](int t) {     // P'_t = A W0_t^T + b0_t/2 -> TMEM columns (t&1)*128
    const int buf = t & 1;
    issue_bias(tmem_base + TM_P + buf * 128, sbase + OFF_ONES, sbase + OFF_BB + buf * (128 * 32), 128);
    issue_gemm(tmem_base + TM_P + buf * 128, sbase + OFF_A, sbase + OFF_W0 + buf * (128 * 64 * 2), 128, 64, true);
    mma_commit(mbarA);
  };

  const int rows_per_tile = SC * N;
  const int ntiles = (a.B + SC - 1) / SC;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0s = tile * SC;
    const int ns = min(SC, a.B - b0s);
    const int nv = ns * N;                                  // valid node rows of this tile
    const size_t grow0 = static_cast<size_t>(b0s) * N;
    const bool live = row < nv;
    const int sc = row / N, ni = row - sc * N;              // scene in tile, node in scene

    // ---- tile prologue: h tile -> bf16 A operand, edge_feat -> smem, W0_0, W1_0, bias operand 0 ----
    {
      const float* ef = a.edge_feat + static_cast<size_t>(b0s) * E * T;
      const int n4 = (ns * E * T) >> 2;                     // SC*E*T*4 bytes is 16-byte aligned when E*T % 4 == 0
      if (((E * T) & 3) == 0) {
        for (int i = tid; i < n4; i += PA_THREADS) cp_async16(sEF + 4 * i, ef + 4 * i);
      } else {
        for (int i = tid; i < ns * E * T; i += PA_THREADS) sEF[i] = __ldg(ef + i);
      }
      load_w0(0, 0);
      if (T > 1) load_w0(1, 1);
      load_w1(0);
      cp_async_commit();
      build_bb(0, 0);
      if (T > 1) build_bb(1, 1);
      if (tid < 128) {
        const float* src = a.h + (grow0 + row) * D;
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (live) {
            float4 x = ldg_f4(src + 8 * k8), y = ldg_f4(src + 8 * k8 + 4);
            pk = make_uint4(pack_bf16_fast(x.x, x.y), pack_bf16_fast(x.z, x.w),
                            pack_bf16_fast(y.x, y.y), pack_bf16_fast(y.z, y.w));
          }
          *reinterpret_cast<uint4*>(smem + OFF_A + canon_off(row, k8, 128)) = pk;
        }
      }
      for (int i = tid; i < 128 * 16; i += PA_THREADS) sS[i] = 0.f;
      cp_async_wait<0>();
      fence_proxy_async_smem();
      fence_before_thread_sync();
      __syncthreads();
      if (warp == 0) { fence_after_thread_sync(); if (elect_one()) issue_gemm1(0); __syncwarp(); }   // warp-uniform issue
    }

    for (int t = 0; t < T; ++t) {
      // (a) P'_t ready in TMEM
      mbar_wait(mbarA, phA); phA ^= 1;
      fence_after_thread_sync();
      // (b) drain this thread's half row as fp16x2: keep it in registers, publish it for the scene mates
      __half2 own[16];
      {
        float r[32];
        tmem_ld32(tmem_lane + TM_P + (t & 1) * 128 + cq * 32, r);
#pragma unroll
        for (int c = 0; c < 16; ++c) own[c] = __floats2half2_rn(r[2 * c], r[2 * c + 1]);
        uint4* dst = reinterpret_cast<uint4*>(sP + row * PLD + cq * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_uint4(*reinterpret_cast<uint32_t*>(&own[4 * c]), *reinterpret_cast<uint32_t*>(&own[4 * c + 1]),
                              *reinterpret_cast<uint32_t*>(&own[4 * c + 2]), *reinterpret_cast<uint32_t*>(&own[4 * c + 3]));
      }
      cp_async_wait<0>();                                    // W0_{t+1} (issued one step ago) has landed
      fence_proxy_async_smem();
      fence_before_thread_sync();
      __syncthreads();                                       // P'_t visible; W0_{t+1} / bias operand staged
      // (c) tensor core: P'_{t+1} while the SIMT part of step t runs
      if (warp == 0 && t + 1 < T) { fence_after_thread_sync(); if (elect_one()) issue_gemm1(t + 1); __syncwarp(); }
      // GEMM 2 of step t-1 has had a whole drain to finish: its operand buffers are free again
      if (t > 0) {
        mbar_wait(mbarB, phB); phB ^= 1;
        fence_after_thread_sync();
        load_w1(t);
      }
      cp_async_commit();                                     // group 1: W1_t (needed by GEMM 2 of this step)
      // W0 / bias buffers (t & 1) were last read by GEMM 1 of step t (complete): stage step t+2
      if (t + 2 < T) { load_w0(t + 2, t & 1); build_bb(t + 2, t & 1); }
      cp_async_commit();                                     // group 2: W0_{t+2} (needed one step later)
      // (d) relu-sum over the scene mates in packed fp16x2: relu(own + p) is one HFMA2.RELU, the
      //     weighted accumulation one HFMA2 (fp16 keeps 11 mantissa bits; the result feeds a bf16 operand)
      __half2 acc[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[c] = __floats2half2_rn(0.f, 0.f);
      float ssum = 0.f;
      if (live) {
        const float* efs = sEF + sc * E * T;
        const __half* prow = sP + (sc * N) * PLD + cq * 32;
        const __half2 one2 = __floats2half2_rn(1.f, 1.f);
        for (int j = 0; j < N; ++j) {
          const float w = efs[(ni * N + j) * T + t] + efs[(j * N + ni) * T + t];
          ssum += w;
          const __half2 w2 = __floats2half2_rn(w, w);
          const uint4* pj = reinterpret_cast<const uint4*>(prow + j * PLD);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 p = pj[c];
            const __half2 p0 = *reinterpret_cast<const __half2*>(&p.x), p1 = *reinterpret_cast<const __half2*>(&p.y);
            const __half2 p2 = *reinterpret_cast<const __half2*>(&p.z), p3 = *reinterpret_cast<const __half2*>(&p.w);
            acc[4 * c] = __hfma2(w2, __hfma2_relu(own[4 * c], one2, p0), acc[4 * c]);
            acc[4 * c + 1] = __hfma2(w2, __hfma2_relu(own[4 * c + 1], one2, p1), acc[4 * c + 1]);
            acc[4 * c + 2] = __hfma2(w2, __hfma2_relu(own[4 * c + 2], one2, p2), acc[4 * c + 2]);
            acc[4 * c + 3] = __hfma2(w2, __hfma2_relu(own[4 * c + 3], one2, p3), acc[4 * c + 3]);
          }
        }
        if (cq == 0) sS[row * 16 + t] = ssum;
      }
      // G_t -> bf16 A operand of GEMM 2
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float2 f0 = __half22float2(acc[4 * g]), f1 = __half22float2(acc[4 * g + 1]);
        float2 f2 = __half22float2(acc[4 * g + 2]), f3 = __half22float2(acc[4 * g + 3]);
        uint4 pk = make_uint4(pack_bf16_fast(f0.x, f0.y), pack_bf16_fast(f1.x, f1.y),
                              pack_bf16_fast(f2.x, f2.y), pack_bf16_fast(f3.x, f3.y));
        *reinterpret_cast<uint4*>(smem + OFF_G + canon_off(row, cq * 4 + g, 128)) = pk;
      }
      cp_async_wait<1>();                                    // W1_t landed (the step t+2 prefetch may still fly)
      fence_proxy_async_smem();
      fence_before_thread_sync();
      __syncthreads();                                       // also: everyone is done reading P'_t
      // (e) agg += G_t W1_t^T
      if (warp == 0) {
        fence_after_thread_sync();
        if (elect_one()) {
          issue_gemm(tmem_base + TM_AGG, sbase + OFF_G, sbase + OFF_W1, D, 128, t > 0);
          mma_commit(mbarB);
        }
        __syncwarp();
      }
    }
    // ---- epilogue: agg = acc + sum_t S[row][t] b1_t ----
    mbar_wait(mbarB, phB); phB ^= 1;
    fence_after_thread_sync();
    {
      float v[16];
      tmem_ld16(tmem_lane + TM_AGG + cq * 16, v);
      if (live) {
        for (int t = 0; t < T; ++t) {
          const float s = sS[row * 16 + t];
          const float* b1 = a.b1 + t * D + cq * 16;
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = fmaf(s, __ldg(b1 + c), v[c]);
        }
        float* dst = a.agg + (grow0 + row) * D + cq * 16;
#pragma unroll
        for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
    }
    fence_before_thread_sync();
    __syncthreads();                                         // TMEM / smem reused by the next tile
  }

  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_thread_sync();
    tmem_dealloc(tmem_base, 512);
  }
  (void)lane; (void)rows_per_tile;
}

// shared-memory footprint; the fused kernel is used when it fits and D == 64
static size_t pair_agg_smem(int N, int T) {
  const int SC = 128 / N;
  return pagg::OFF_EF + static_cast<size_t>(SC) * N * N * T * 4;
}

bool pair_agg_fits(int N, int D, int T) {
  return D == 64 && N >= 1 && N <= 128 && T <= 15 && pair_agg_smem(N, T) <= 227 * 1024;
}

int launch_pair_agg_tc(const float* h, const float* edge_feat, int B, int N, int T,
                       const gn_stage_weights* w, float* agg, cudaStream_t st) {
  if (!w->tc_agg_w0 || !w->tc_agg_w1) return GN_E_NULL;
  PairAggArgs a;
  a.h = h; a.edge_feat = edge_feat;
  a.w0 = static_cast<const __nv_bfloat16*>(w->tc_agg_w0);
  a.w1 = static_cast<const __nv_bfloat16*>(w->tc_agg_w1);
  a.b0 = w->agg_b0; a.b1 = w->agg_b1; a.agg = agg;
  a.B = B; a.N = N; a.T = T; a.SC = 128 / N;
  const size_t smem = pair_agg_smem(N, T);
  cudaError_t e = cudaFuncSetAttribute(pair_agg_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int ntiles = (B + a.SC - 1) / a.SC;
  const int grid = ntiles < GN_SM_COUNT ? ntiles : GN_SM_COUNT;
  {
    ProfScope ps__("pair_agg_tc", st);
    pair_agg_tc_kernel<<<grid, PA_THREADS, smem, st>>>(a);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
