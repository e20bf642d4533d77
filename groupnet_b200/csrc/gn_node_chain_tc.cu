// Fused node-level GEMM chains of the bf16 tensor-core path (rows = B*N node rows, 128 per tile):
//
//   PRE   h (D=64) -> 256 (+b, ReLU) -> 64 (+b) = x'  -> 64 = pq          node2edge_start_mlp (:84,:125)
//                                                                          + split attention layer 0 (:80,:134)
//   POST  [agg | h] / N (128) -> 128 (+b, ReLU) -> Dout (+b) = node_feat   edge2node /N (:120,:355) + closing
//                                                                          MLP (:77,:195,:441)
//
// A chain is a list of up to 3 steps {W (bf16 canonical, resident in smem), bias (through the bias
// MMA), ReLU, optional fp32 store to HBM}; the activation of step i is the A operand of step i+1 and
// never leaves the SM.  A 256-wide step is drained in two 128-column halves that feed the next step
// as two K = 128 accumulation chunks (one 32 KB buffer).  Two independent 128-thread groups per CTA
// (own tile stream, buffers, mbarrier, 256 TMEM columns) overlap each other's SIMT and MMA phases.
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {

namespace nchain {
constexpr uint32_t OFF_ONES = 0;                         // ones operand                  4 KB
constexpr uint32_t BB_BYTES = 12 * 1024;                 // bias operands, sum of N * 32 bytes over the steps
constexpr uint32_t OFF_BB = OFF_ONES + 128 * 32;
constexpr uint32_t OFF_W = OFF_BB + BB_BYTES;            // resident weights, up to 72 KB
constexpr uint32_t W_BYTES = 72 * 1024;
constexpr uint32_t OFF_BAR = OFF_W + W_BYTES;            // 2 mbarriers + tmem slot
constexpr uint32_t OFF_GRP = OFF_BAR + 32;               // per group: bufA 32 KB | bufB 32 KB
constexpr uint32_t BUF_BYTES = 128 * 128 * 2;
constexpr uint32_t GRP_BYTES = 2 * BUF_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_GRP + 2 * GRP_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "node chain kernel exceeds shared memory");
}  // namespace nchain

__device__ __forceinline__ void nchain_group_bar(int grp) {
  asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory");
}

__global__ void __launch_bounds__(GN_THREADS, 1)
node_chain_tc_kernel(NodeChainArgs a) {
  using namespace nchain;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127, row = gtid;
  unsigned char* g = smem + OFF_GRP + grp * GRP_BYTES;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + grp;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);

  // ---- one-time setup: resident weights, bias operands ----
  uint32_t woff[3], boff[3];
  {
    uint32_t o = 0, bo = 0;
    for (int s = 0; s < a.nsteps; ++s) {
      woff[s] = o;
      boff[s] = bo;
      bo += static_cast<uint32_t>(a.step[s].N) * 32;
      const int n16 = a.step[s].N * a.step[s].K / 8;
      const uint4* src = reinterpret_cast<const uint4*>(a.step[s].W);
      for (int i = tid; i < n16; i += GN_THREADS)
        *reinterpret_cast<uint4*>(smem + OFF_W + o + 16 * i) = __ldg(src + i);
      if (a.step[s].bias != nullptr) build_bias_operand(smem + OFF_BB + boff[s], a.step[s].bias, a.step[s].N, tid, GN_THREADS);
      o += static_cast<uint32_t>(n16) * 16;
    }
    build_ones_operand(smem + OFF_ONES, tid, GN_THREADS);
  }
  if ((tid >> 5) == 0) tmem_alloc(tmem_slot, 512);
  if (gtid == 32) mbar_init(mbar, 1);
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_grp = *tmem_slot + grp * 256;
  const uint32_t tmem_row = tmem_grp + (static_cast<uint32_t>((gtid >> 5) * 32) << 16);
  const uint32_t sbase = smem_u32(smem), gbase = smem_u32(g);
  uint32_t phase = 0;
  const int K0tot = a.K0 + a.K1;

  const long long ntiles = (a.R + 127) / 128;
  for (long long tile = static_cast<long long>(blockIdx.x) * 2 + grp; tile < ntiles;
       tile += static_cast<long long>(gridDim.x) * 2) {
    const long long grow = tile * 128 + row;
    const bool live = grow < a.R;
    // ---- stage the input row (fp32, up to two concatenated sources, optional /div) into bufA ----
    {
      const long long r = live ? grow : 0;
      for (int k8 = 0; k8 < (K0tot >> 3); ++k8) {
        const int k = k8 * 8;
        const float* src = (k < a.K0) ? a.A0 + r * a.lda0 + k : a.A1 + r * a.lda1 + (k - a.K0);
        float4 x = ldg_f4(src), y = ldg_f4(src + 4);
        if (a.a_div != 0.f) {
          x.x = __fdividef(x.x, a.a_div); x.y = __fdividef(x.y, a.a_div); x.z = __fdividef(x.z, a.a_div); x.w = __fdividef(x.w, a.a_div);
          y.x = __fdividef(y.x, a.a_div); y.y = __fdividef(y.y, a.a_div); y.z = __fdividef(y.z, a.a_div); y.w = __fdividef(y.w, a.a_div);
        }
        uint4 pk = make_uint4(pack_bf16_fast(x.x, x.y), pack_bf16_fast(x.z, x.w),
                              pack_bf16_fast(y.x, y.y), pack_bf16_fast(y.z, y.w));
        *reinterpret_cast<uint4*>(g + canon_off(row, k8, 128)) = pk;
      }
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    nchain_group_bar(grp);

    int cur = 0;                       // buffer holding the current A operand (0 = bufA, 1 = bufB)
    int Kcur = K0tot;
    bool acc_pending = false;          // the current step was already started by a split predecessor
    for (int s = 0; s < a.nsteps; ++s) {
      const NodeChainStep st = a.step[s];
      const uint32_t w_addr = sbase + OFF_W + woff[s];
      const uint32_t a_addr = gbase + cur * BUF_BYTES;
      if (!acc_pending) {
        if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
          fence_after_thread_sync();
          if (elect_one()) {
            if (st.bias != nullptr) issue_bias(tmem_grp, sbase + OFF_ONES, sbase + OFF_BB + boff[s], st.N);
            issue_gemm(tmem_grp, a_addr, w_addr, st.N, Kcur, st.bias != nullptr);
            mma_commit(mbar);
          }
          __syncwarp();
        }
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_thread_sync();
      }
      acc_pending = false;
      const bool last = (s + 1 == a.nsteps);
      unsigned char* nxt = g + (cur ^ 1) * BUF_BYTES;
      if (st.N == 256 && !last) {
        // split drain: two 128-column halves feed step s+1 as two K = 128 accumulation chunks
        const NodeChainStep sn = a.step[s + 1];
        const uint32_t wn_addr = sbase + OFF_W + woff[s + 1];
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[4][32];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld32_nowait(tmem_row + hh * 128 + 32 * c, r[c]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c][8 * q + j]);
              uint4 pk = st.relu
                  ? make_uint4(pack_bf16_relu(v[0], v[1]), pack_bf16_relu(v[2], v[3]), pack_bf16_relu(v[4], v[5]), pack_bf16_relu(v[6], v[7]))
                  : make_uint4(pack_bf16_fast(v[0], v[1]), pack_bf16_fast(v[2], v[3]), pack_bf16_fast(v[4], v[5]), pack_bf16_fast(v[6], v[7]));
              *reinterpret_cast<uint4*>(nxt + canon_off(row, 4 * c + q, 128)) = pk;
            }
          fence_proxy_async_smem();
          fence_before_thread_sync();
          nchain_group_bar(grp);
          if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
            fence_after_thread_sync();
            if (elect_one()) {
              if (hh == 0 && sn.bias != nullptr) issue_bias(tmem_grp, sbase + OFF_ONES, sbase + OFF_BB + boff[s + 1], sn.N);
              // K chunk hh of W_{s+1}: k-groups [16*hh, 16*hh+16), each sn.N * 16 bytes
              issue_gemm(tmem_grp, smem_u32(nxt), wn_addr + hh * 16 * (sn.N * 16), sn.N, 128, hh == 1 || sn.bias != nullptr);
              mma_commit(mbar);
            }
            __syncwarp();
          }
          mbar_wait(mbar, phase); phase ^= 1;     // chunk consumed: the buffer is free / the result is ready
          fence_after_thread_sync();
        }
        acc_pending = true;
        // the A operand of step s+1 was consumed in place; its own output goes to the other buffer
        Kcur = 256;
        continue;
      }
      // plain drain of N <= 128 columns (or the final step): fp32 store and / or next A operand
      for (int c0 = 0; c0 < st.N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_row + c0, v);
        if (st.out != nullptr && live) {
          float* dst = st.out + grow * st.ldo + st.out_col0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (!last) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 pk = st.relu
                ? make_uint4(pack_bf16_relu(v[8 * q], v[8 * q + 1]), pack_bf16_relu(v[8 * q + 2], v[8 * q + 3]),
                             pack_bf16_relu(v[8 * q + 4], v[8 * q + 5]), pack_bf16_relu(v[8 * q + 6], v[8 * q + 7]))
                : make_uint4(pack_bf16_fast(v[8 * q], v[8 * q + 1]), pack_bf16_fast(v[8 * q + 2], v[8 * q + 3]),
                             pack_bf16_fast(v[8 * q + 4], v[8 * q + 5]), pack_bf16_fast(v[8 * q + 6], v[8 * q + 7]));
            *reinterpret_cast<uint4*>(nxt + canon_off(row, (c0 >> 3) + q, 128)) = pk;
          }
        }
      }
      fence_proxy_async_smem();
      fence_before_thread_sync();
      nchain_group_bar(grp);
      cur ^= 1;
      Kcur = st.N;
    }
  }

  fence_before_thread_sync();
  __syncthreads();
  if ((tid >> 5) == 0) {
    fence_after_thread_sync();
    tmem_dealloc(*tmem_slot, 512);
  }
}

// Requirements: N % 32 == 0 (or the tail is not stored), N <= 256, K0+K1 <= 128 and % 16 == 0, a 256-wide
// step is followed by a step with N <= 128; total weights <= 80 KB; every stored column count % 32 == 0.
int launch_node_chain_tc(const NodeChainArgs& a, const char* name, cudaStream_t st) {
  if (a.R <= 0) return GN_OK;
  if (a.nsteps < 1 || a.nsteps > 3) return GN_E_SHAPE;
  const int K0 = a.K0 + a.K1;
  if ((K0 & 15) || K0 > 128 || (a.K0 & 7)) return GN_E_SHAPE;
  size_t wbytes = 0, bbytes = 0;
  int K = K0;
  for (int s = 0; s < a.nsteps; ++s) {
    const NodeChainStep& p = a.step[s];
    if (p.K != K || (p.N & 31) || p.N > 256 || p.N < 32) return GN_E_SHAPE;
    if (p.N == 256 && (s + 1 == a.nsteps || a.step[s + 1].N > 128)) return GN_E_SHAPE;
    if (p.N > 128 && p.N != 256) return GN_E_SHAPE;
    if (p.out != nullptr && ((p.ldo & 3) || (p.out_col0 & 3))) return GN_E_ALIGN;
    wbytes += static_cast<size_t>(p.N) * p.K * 2;
    bbytes += static_cast<size_t>(p.N) * 32;
    K = p.N;
  }
  if (wbytes > nchain::W_BYTES || bbytes > nchain::BB_BYTES) return GN_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(node_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(nchain::SMEM_BYTES));
  if (e != cudaSuccess) return static_cast<int>(e);
  long long ntiles = (a.R + 127) / 128, want = (ntiles + 1) / 2;
  const int grid = want < GN_SM_COUNT ? static_cast<int>(want) : GN_SM_COUNT;
  {
    ProfScope ps__(name, st);
    node_chain_tc_kernel<<<grid, GN_THREADS, nchain::SMEM_BYTES, st>>>(a);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
