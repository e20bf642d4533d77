// Fused edge_aggregation for the hyper layers (bf16 tensor-core path), as written in the
// reference (model/MS_HGNN_batch.py:259-265):
//   ef_e = sum_t edge_feat[e,t] * (W1_t relu(W0_t eo_e + b0_t) + b1_t),   eo = H @ h
// over tiles of 128 hyperedge rows.  The per-row scale is applied to the hidden activations, so
// the T second Linears accumulate into ONE TMEM accumulator (K = T*128):
//   per t:  G1  hidden_t = eo W0_t^T + b0_t        [128 x 64] x [64 -> 128]   (bias via the bias MMA)
//           drain: ReLU, * edge_feat[row][t], bf16 -> smem A operand
//           G2  acc += hidden_t W1_t^T             [128 x 128] x [128 -> 64]
//   epilogue: ef = acc + sum_t edge_feat[row][t] b1_t
// The (B*E, T*128) hidden tensor of the unfused path (1.8 GB at the NBA shape) never exists.
//
// Two independent 128-thread groups per CTA (tile streams), each with its own operand buffers,
// mbarriers and 256 TMEM columns; W0_t / W1_t (16 KB each) are streamed from L2 with cp.async
// one step ahead.  Bound: tensor pipe / handoff latency; algorithmic FLOPs per row:
// 2 * T * 2 * 64 * 128; algorithmic HBM bytes per row: 256 (eo) + 4T (edge_feat) + 256 (ef).
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {

struct HyperAggArgs {
  const float* eo;           // (R, 64)
  const float* edge_feat;    // (R, T)
  const __nv_bfloat16* w0;   // canonical (T*128, 64)
  const __nv_bfloat16* w1;   // canonical (64, T*128)
  const float* b0;           // (T*128)
  const float* b1;           // (T, 64)
  float* ef;                 // (R, 64)
  long long R;
  int T;
};

namespace hagg {
constexpr uint32_t OFF_ONES = 0;                        // 4 KB
constexpr uint32_t OFF_B1 = OFF_ONES + 128 * 32;        // b1 [16][64] fp32      4 KB
constexpr uint32_t OFF_BAR = OFF_B1 + 16 * 64 * 4;      // 2 x (mbarA, mbarB) + tmem slot
constexpr uint32_t OFF_GRP = OFF_BAR + 64;
constexpr uint32_t G_A = 0;                             // eo tile bf16 [128 x 64]      16 KB
constexpr uint32_t G_A1 = G_A + 128 * 64 * 2;           // hidden_t bf16 [128 x 128]    32 KB
constexpr uint32_t G_W0 = G_A1 + 128 * 128 * 2;         // W0_t [128 x 64]              16 KB
constexpr uint32_t G_W1 = G_W0 + 128 * 64 * 2;          // W1_t double buffer           32 KB
constexpr uint32_t G_BB = G_W1 + 2 * 64 * 128 * 2;      // b0_t bias operand             4 KB
constexpr uint32_t GRP_BYTES = G_BB + 128 * 32;
constexpr uint32_t SMEM_BYTES = OFF_GRP + 2 * GRP_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "hyper_agg kernel exceeds shared memory");
}  // namespace hagg

__device__ __forceinline__ void hagg_group_bar(int grp) {
  asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory");
}

template <int TT>
__global__ void __launch_bounds__(GN_THREADS, 1)
hyper_agg_tc_kernel(HyperAggArgs a) {
  using namespace hagg;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127, row = gtid;
  unsigned char* g = smem + OFF_GRP + grp * GRP_BYTES;
  float* sb1 = reinterpret_cast<float*>(smem + OFF_B1);
  uint64_t* mbarA = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + 2 * grp;
  uint64_t* mbarB = mbarA + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 32);
  const int T = TT > 0 ? TT : a.T;
  constexpr int TU = TT > 0 ? TT : 15;
  const int NT = T * 128;

  build_ones_operand(smem + OFF_ONES, tid, GN_THREADS);
  for (int i = tid; i < 16 * 64; i += GN_THREADS) sb1[i] = (i < T * 64) ? __ldg(a.b1 + i) : 0.f;
  if ((tid >> 5) == 0) tmem_alloc(tmem_slot, 512);
  if (gtid == 32) { mbar_init(mbarA, 1); mbar_init(mbarB, 1); }
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_grp = *tmem_slot + grp * 256;
  const uint32_t tmem_row = tmem_grp + (static_cast<uint32_t>((gtid >> 5) * 32) << 16);
  const uint32_t sbase = smem_u32(smem), gbase = smem_u32(g);
  uint32_t phA = 0, phB = 0;

  auto load_w0 = [&](int t) {          // rows [t*128, +128) of the canonical (NT x 64) operand
    for (int i = gtid; i < 8 * 128; i += 128) {
      const int k8 = i >> 7, n = i & 127;
      cp_async16(g + G_W0 + (k8 * 128 + n) * 16, a.w0 + (static_cast<size_t>(k8) * NT + t * 128 + n) * 8);
    }
  };
  auto load_w1 = [&](int t, int buf) { // k-groups [t*16, +16) of the canonical (64 x NT) operand: contiguous
    const __nv_bfloat16* src = a.w1 + static_cast<size_t>(t) * 16 * 64 * 8;
    for (int i = gtid; i < 16 * 64; i += 128) cp_async16(g + G_W1 + buf * (64 * 128 * 2) + i * 16, src + i * 8);
  };
  auto build_bb = [&](int t) {
    const float v = __ldg(a.b0 + t * 128 + gtid);
    __nv_bfloat16 hi = __float2bfloat16_rn(v);
    __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const uint32_t w0 = static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&hi)) |
                        (static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&lo)) << 16);
    *reinterpret_cast<uint4*>(g + G_BB + gtid * 16) = make_uint4(w0, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(g + G_BB + 128 * 16 + gtid * 16) = make_uint4(0u, 0u, 0u, 0u);
  };

  const long long ntiles = (a.R + 127) / 128;
  for (long long tile = static_cast<long long>(blockIdx.x) * 2 + grp; tile < ntiles;
       tile += static_cast<long long>(gridDim.x) * 2) {
    const long long grow = tile * 128 + row;
    const bool live = grow < a.R;
    // ---- prologue: eo row -> bf16 A operand; edge_feat row -> registers; W0_0, W1_0, bias operand 0
    load_w0(0);
    load_w1(0, 0);
    cp_async_commit();
    build_bb(0);
    float sc[TU];
#pragma unroll
    for (int t = 0; t < TU; ++t) sc[t] = (live && t < T) ? __ldg(a.edge_feat + static_cast<size_t>(grow) * T + t) : 0.f;
    {
      const float* src = a.eo + static_cast<size_t>(live ? grow : 0) * 64;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        float4 x = ldg_stream_f4(src + 8 * k8), y = ldg_stream_f4(src + 8 * k8 + 4);
        uint4 pk = make_uint4(pack_bf16_fast(x.x, x.y), pack_bf16_fast(x.z, x.w),
                              pack_bf16_fast(y.x, y.y), pack_bf16_fast(y.z, y.w));
        *reinterpret_cast<uint4*>(g + G_A + canon_off(row, k8, 128)) = pk;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    fence_before_thread_sync();
    hagg_group_bar(grp);
    if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
      fence_after_thread_sync();
      if (elect_one()) {
        issue_bias(tmem_grp, sbase + OFF_ONES, gbase + G_BB, 128);
        issue_gemm(tmem_grp, gbase + G_A, gbase + G_W0, 128, 64, true);
        mma_commit(mbarA);
      }
      __syncwarp();
    }

#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      // hidden_t ready; W0 / bias-operand buffers are free again
      mbar_wait(mbarA, phA); phA ^= 1;
      fence_after_thread_sync();
      if (t + 1 < T) { load_w0(t + 1); build_bb(t + 1); }
      // G2 of step t-1 finished: A1 and the other W1 buffer are free
      if (t > 0) { mbar_wait(mbarB, phB); phB ^= 1; fence_after_thread_sync(); }
      if (t + 1 < T) load_w1(t + 1, (t + 1) & 1);
      cp_async_commit();
      // drain: relu(hidden) * edge_feat[row][t] -> bf16 A operand of G2
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < TU; ++u) if (u == t) s = sc[u];
      {
        uint32_t r[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32_nowait(tmem_row + 32 * c, r[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(__uint_as_float(r[c][8 * q + j]), 0.f) * s;
            uint4 pk = make_uint4(pack_bf16_fast(v[0], v[1]), pack_bf16_fast(v[2], v[3]),
                                  pack_bf16_fast(v[4], v[5]), pack_bf16_fast(v[6], v[7]));
            *reinterpret_cast<uint4*>(g + G_A1 + canon_off(row, 4 * c + q, 128)) = pk;
          }
        }
      }
      cp_async_wait<0>();                               // W0_{t+1}, W1_{t+1} (and W1_t from the previous step)
      fence_proxy_async_smem();
      fence_before_thread_sync();
      hagg_group_bar(grp);
      if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
        fence_after_thread_sync();
        if (elect_one()) {
          issue_gemm(tmem_grp + 128, gbase + G_A1, gbase + G_W1 + (t & 1) * (64 * 128 * 2), 64, 128, t > 0);
          mma_commit(mbarB);
          if (t + 1 < T) {                                 // next hidden right behind it on the tensor pipe
            issue_bias(tmem_grp, sbase + OFF_ONES, gbase + G_BB, 128);
            issue_gemm(tmem_grp, gbase + G_A, gbase + G_W0, 128, 64, true);
            mma_commit(mbarA);
          }
        }
        __syncwarp();
      }
    }
    // ---- epilogue: ef = acc + sum_t edge_feat[row][t] b1_t
    mbar_wait(mbarB, phB); phB ^= 1;
    fence_after_thread_sync();
    {
      uint32_t r[2][32];
      tmem_ld32_nowait(tmem_row + 128, r[0]);
      tmem_ld32_nowait(tmem_row + 160, r[1]);
      tmem_ld_wait();
      if (live) {
        float* dst = a.ef + static_cast<size_t>(grow) * 64;
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x = __uint_as_float(r[(c + j) >> 5][(c + j) & 31]);
#pragma unroll
            for (int t = 0; t < TU; ++t) x = fmaf(sc[t], sb1[t * 64 + c + j], x);
            v[j] = x;
          }
          *reinterpret_cast<float4*>(dst + c) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    fence_before_thread_sync();
    hagg_group_bar(grp);
  }

  fence_before_thread_sync();
  __syncthreads();
  if ((tid >> 5) == 0) {
    fence_after_thread_sync();
    tmem_dealloc(*tmem_slot, 512);
  }
}

bool hyper_agg_fits(int D, int T) { return D == 64 && T >= 1 && T <= 15; }

int launch_hyper_agg_tc(const float* eo, const float* edge_feat, long long R, int T,
                        const gn_stage_weights* w, float* ef, cudaStream_t st) {
  if (!w->tc_agg_w0 || !w->tc_agg_w1) return GN_E_NULL;
  if (R <= 0) return GN_OK;
  HyperAggArgs a;
  a.eo = eo; a.edge_feat = edge_feat;
  a.w0 = static_cast<const __nv_bfloat16*>(w->tc_agg_w0);
  a.w1 = static_cast<const __nv_bfloat16*>(w->tc_agg_w1);
  a.b0 = w->agg_b0; a.b1 = w->agg_b1; a.ef = ef; a.R = R; a.T = T;
  long long ntiles = (R + 127) / 128, want = (ntiles + 1) / 2;
  const int grid = want < GN_SM_COUNT ? static_cast<int>(want) : GN_SM_COUNT;
  cudaError_t e;
  if (T == 10) {
    e = cudaFuncSetAttribute(hyper_agg_tc_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(hagg::SMEM_BYTES));
    if (e != cudaSuccess) return static_cast<int>(e);
    ProfScope ps__("hyper_agg_tc", st);
    hyper_agg_tc_kernel<10><<<grid, GN_THREADS, hagg::SMEM_BYTES, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(hyper_agg_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(hagg::SMEM_BYTES));
    if (e != cudaSuccess) return static_cast<int>(e);
    ProfScope ps__("hyper_agg_tc", st);
    hyper_agg_tc_kernel<0><<<grid, GN_THREADS, hagg::SMEM_BYTES, st>>>(a);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
