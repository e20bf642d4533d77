// Kernel (c), tensor-core path: the per-edge chain of one message-passing stage
// on tcgen05 / TMEM, over tiles of 128 edge rows.
//
//   [PAIR only] node2edge for the pairwise layer (model/MS_HGNN_batch.py:122-141):
//       per edge e=(i,j): attention logits from the split 128->32->1 MLP, softmax
//       over all N nodes, edges_e = w_i x'_i + w_j x'_j  -> bf16 A operand
//   G1  [128 x 64]  x W_init0^T  [64 -> 128]   epilogue: +b, ReLU, bf16 -> smem     (:43)
//   G2  [128 x 128] x W_init1^T  [128 -> 64]   epilogue: +b,       bf16 -> smem (z) (:43)
//   G3  [128 x 64]  x W_df0^T    [64 -> 256]   epilogue: +b, ReLU, bf16 -> smem     (:45,:47)
//   G4  [128 x 256] x W_df1^T    [256 -> 16]   epilogue: +b, Gumbel softmax over T, sigmoid,
//                                              factor * dist                        (:45-53, :446-520)
//
// One persistent CTA per SM, 256 threads = two independent 128-thread groups.
// Each group runs its own tile stream with its own operand buffers, mbarrier
// and 256 TMEM columns, so one group's SIMT phases (row staging, epilogues)
// overlap the other group's MMAs; the tensor core executes both groups'
// tcgen05.mma in issue order.  Inside a group thread t owns tile row t
// (TMEM lane t) in every phase.  All weights (72 KB bf16, canonical no-swizzle
// K-major layout) stay resident in shared memory for the life of the CTA.
//
// G3's 256-column accumulator is drained in two halves through one 32 KB
// buffer (G4 = two K=128 accumulation steps) to fit two groups in 227 KB.
//
// Roofline: tensor pipe.  Algorithmic work per edge row: 2 * (64*128 + 128*64 +
// 64*256 + 256*(T+1)) FLOP; algorithmic HBM bytes per row: 8T (dist, edge_feat)
// + PAIR: 2 * 512 / N (x', pq of its scene, amortised) | LOAD: 256 (edges).
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {

struct TcChainArgs {
  const void* w1; const void* w2; const void* w3; const void* w4;   // bf16, canonical layout
  const float* b1; const float* b2; const float* b3; const float* b4;
  const float* att_b0; const float* att_w1; const float* att_b1;    // PAIR: attention_mlp tail
  const float* edges;                 // LOAD: (R, 64) fp32
  const float* xprime; const float* pq;   // PAIR: (B*N, 64) fp32 each
  int N, E, T;
  long long R;
  const float* U; int noise_mode; unsigned long long seed; long long scene_offset; int stage_index;
  float* dist_out; float* edge_feat;
  unsigned long long* trace;          // optional: clock64() per phase, block 0, first tiles (gn_profile_set_trace)
  int tps;                            // PAIR, E >= 128: tiles per scene (scene-aligned tiling); 0 = linear 128-row tiles
};

namespace tcmlp {
constexpr uint32_t OFF_W1 = 0;                       // N=128 K=64   16 KB
constexpr uint32_t OFF_W2 = OFF_W1 + 128 * 64 * 2;   // N=64  K=128  16 KB
constexpr uint32_t OFF_W3 = OFF_W2 + 64 * 128 * 2;   // N=256 K=64   32 KB
constexpr uint32_t OFF_W4 = OFF_W3 + 256 * 64 * 2;   // N=16  K=256   8 KB
constexpr uint32_t OFF_ONES = OFF_W4 + 16 * 256 * 2; // [128 x 16] ones operand        4 KB
constexpr uint32_t OFF_BB1 = OFF_ONES + 128 * 32;    // [N x 16] bias operands: 128, 64, 256, 16 rows
constexpr uint32_t OFF_BB2 = OFF_BB1 + 128 * 32;
constexpr uint32_t OFF_BB3 = OFF_BB2 + 64 * 32;
constexpr uint32_t OFF_BB4 = OFF_BB3 + 256 * 32;
constexpr uint32_t OFF_ATT = OFF_BB4 + 16 * 32;      // attention tail: b0[32] | w1[32] | b1 (+pad)
constexpr uint32_t OFF_GRP = OFF_ATT + 80 * 4;       // per group: A0 [128 x 64] 16 KB | A1 [128 x 128] 32 KB
constexpr uint32_t A0_BYTES = 128 * 64 * 2, A1_BYTES = 128 * 128 * 2;
constexpr uint32_t GRP_BYTES = A0_BYTES + A1_BYTES;
constexpr uint32_t OFF_BAR = OFF_GRP + 2 * GRP_BYTES;   // 2 mbarriers + tmem slot
constexpr uint32_t OFF_NODE = OFF_BAR + 32;             // PAIR: per group x'[MAXN][68] | pq[MAXN][68] fp32
constexpr int MAXN = 36;                                // nodes a 128-row tile may span (see pair_fits)
constexpr int NLD = 68;                                 // padded node row (floats): rows 4 banks apart
constexpr uint32_t NODE_BYTES = 2 * MAXN * NLD * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 32;
constexpr uint32_t SMEM_BYTES_PAIR = OFF_NODE + 2 * NODE_BYTES;
static_assert(SMEM_BYTES_PAIR <= 227 * 1024, "pair chain kernel exceeds shared memory");
}  // namespace tcmlp

__device__ __forceinline__ void chain_group_bar(int grp) {
  asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory");
}

// NCH*32 accumulator columns (bias already inside, via the bias MMA) -> (ReLU) -> bf16 -> canonical
// A rows.  All NCH TMEM loads are issued before one wait; 1 LDTM + 16 CVT + 4 STS per 32 columns.
template <bool RELU, int NCH>
__device__ __forceinline__ void drain_to_smem(uint32_t tmem_addr, unsigned char* dst, int row) {
  uint32_t r[NCH][32];
#pragma unroll
  for (int c = 0; c < NCH; ++c) tc::tmem_ld32_nowait(tmem_addr + 32 * c, r[c]);
  tc::tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c][8 * g + j]);
      uint4 pk;
      if (RELU) {
        pk = make_uint4(tc::pack_bf16_relu(v[0], v[1]), tc::pack_bf16_relu(v[2], v[3]),
                        tc::pack_bf16_relu(v[4], v[5]), tc::pack_bf16_relu(v[6], v[7]));
      } else {
        pk = make_uint4(tc::pack_bf16_fast(v[0], v[1]), tc::pack_bf16_fast(v[2], v[3]),
                        tc::pack_bf16_fast(v[4], v[5]), tc::pack_bf16_fast(v[6], v[7]));
      }
      *reinterpret_cast<uint4*>(dst + tc::canon_off(row, 4 * c + g, 128)) = pk;
    }
  }
}

// TT = compile-time number of edge types (6 pairwise, 10 hyper); 0 = runtime a.T (<= 15)
#ifdef GN_ENABLE_TRACE   // make NVFLAGS+=-DGN_ENABLE_TRACE: clock64() stamps per phase (profiles/trace_chain.py)
#define GN_TRACE(pt) do { if (a.trace != nullptr && blockIdx.x == 0 && gtid == 0 && titer < 8) \
    a.trace[(grp * 8 + titer) * 16 + (pt)] = clock64(); } while (0)
#else
#define GN_TRACE(pt) do { } while (0)
#endif

template <bool PAIR, int TT, bool ALIGNED = false>
__global__ void __launch_bounds__(GN_THREADS, 1)
edge_chain_tc_kernel(TcChainArgs a) {
  using namespace tcmlp;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127;
  const int row = gtid;                                   // tile row == TMEM lane
  const float* att = reinterpret_cast<const float*>(smem + OFF_ATT);   // b0[32] | w1[32] | b1
  unsigned char* sA0 = smem + OFF_GRP + grp * GRP_BYTES;
  unsigned char* sA1 = sA0 + A0_BYTES;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + grp;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);

  // ---- one-time setup: weights, bias operands -> smem, mbarriers, TMEM ----
  {
    const uint4* src[4] = {static_cast<const uint4*>(a.w1), static_cast<const uint4*>(a.w2),
                           static_cast<const uint4*>(a.w3), static_cast<const uint4*>(a.w4)};
    const uint32_t off[4] = {OFF_W1, OFF_W2, OFF_W3, OFF_W4};
    const int n16[4] = {128 * 64 / 8, 64 * 128 / 8, 256 * 64 / 8, 16 * 256 / 8};
#pragma unroll
    for (int m = 0; m < 4; ++m)
      for (int i = tid; i < n16[m]; i += GN_THREADS)
        *reinterpret_cast<uint4*>(smem + off[m] + 16 * i) = __ldg(src[m] + i);
    build_ones_operand(smem + OFF_ONES, tid, GN_THREADS);
    build_bias_operand(smem + OFF_BB1, a.b1, 128, tid, GN_THREADS);
    build_bias_operand(smem + OFF_BB2, a.b2, 64, tid, GN_THREADS);
    build_bias_operand(smem + OFF_BB3, a.b3, 256, tid, GN_THREADS);
    build_bias_operand(smem + OFF_BB4, a.b4, 16, tid, GN_THREADS);
    if (PAIR) {
      float* attw = reinterpret_cast<float*>(smem + OFF_ATT);
      if (tid < 32) attw[tid] = __ldg(a.att_b0 + tid);
      else if (tid < 64) attw[tid] = __ldg(a.att_w1 + tid - 32);
      else if (tid == 64) attw[64] = __ldg(a.att_b1);
    }
  }
  if ((tid >> 5) == 0) tmem_alloc(tmem_slot, 512);
  if (gtid == 32) mbar_init(mbar, 1);
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_grp = *tmem_slot + grp * 256;
  const uint32_t tmem_row = tmem_grp + (static_cast<uint32_t>((gtid >> 5) * 32) << 16);
  const uint32_t sbase = smem_u32(smem), sA0_addr = smem_u32(sA0), sA1_addr = smem_u32(sA1);
  uint32_t phase = 0;
  const int T = TT > 0 ? TT : a.T;
  constexpr int TU = TT > 0 ? TT : GN_SMALL_OUT - 1;      // unroll bound

  // E >= 128 (N >= 12): tiles are aligned to scenes (ceil(E/128) per scene, the last one ragged) so a tile's node
  // block is ONE scene's N rows; with linear tiles a tile of the N = 20 fish shape would span 40 nodes
  const int tps = (PAIR && ALIGNED) ? a.tps : 0;   // compile-time 0 for the linear-tile instantiations
  const long long ntiles = tps ? (a.R / a.E) * tps : (a.R + 127) / 128;
  float* nx = reinterpret_cast<float*>(smem + OFF_NODE + (PAIR ? grp * NODE_BYTES : 0));   // x' rows
  float* np = nx + MAXN * NLD;                                                              // pq rows
  const long long total_nodes = PAIR ? (a.R / a.E) * a.N : 0;
  // cp.async the contiguous node block [b_lo*N, b_lo*N + MAXN) of tile t into this group's buffer
  auto prefetch_nodes = [&](long long t) {
    if (t < ntiles) {
      const long long node0 = (tps ? t / tps : (t * 128) / a.E) * a.N;
      const int cnt = static_cast<int>(min(static_cast<long long>(MAXN), total_nodes - node0));
      for (int i = gtid; i < cnt * 16; i += 128) {
        const int n = i >> 4, c = i & 15;
        cp_async16(nx + n * NLD + 4 * c, a.xprime + (node0 + n) * 64 + 4 * c);
        cp_async16(np + n * NLD + 4 * c, a.pq + (node0 + n) * 64 + 4 * c);
      }
    }
    cp_async_commit();
  };
  if (PAIR) prefetch_nodes(static_cast<long long>(blockIdx.x) * 2 + grp);
  int titer = -1;
  for (long long tile = static_cast<long long>(blockIdx.x) * 2 + grp; tile < ntiles;
       tile += static_cast<long long>(gridDim.x) * 2) {
    const long long tscene = tps ? tile / tps : 0;
    const int tchunk = tps ? static_cast<int>(tile - tscene * tps) : 0;
    const long long grow = tps ? tscene * a.E + tchunk * 128 + row : tile * 128 + row;
    const bool live = grow < a.R && (!tps || tchunk * 128 + row < a.E);
    ++titer;
    GN_TRACE(0);

    // ---- stage this thread's edge row as 64 bf16 of the A operand ----
    if (PAIR) {
      // the tile's node block (x', pq of the scenes it spans) was prefetched into smem
      cp_async_wait<0>();
      chain_group_bar(grp);
      const long long b_lo = tps ? tscene : (tile * 128) / a.E;
      float wi = 0.f, wj = 0.f;
      const float* xi = nx;
      const float* xj = nx;
      if (live) {
        const int N = a.N;
        const long long b = grow / a.E;
        const int e = static_cast<int>(grow - b * a.E), i = e / N, j = e - i * N;
        const int li = static_cast<int>(b - b_lo) * N + i, lj = static_cast<int>(b - b_lo) * N + j;
        const float* pi = np + li * NLD;
        const float* pj = np + lj * NLD;
        xi = nx + li * NLD;
        xj = nx + lj * NLD;
        float ai = 0.f, aj = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 32; k4 += 4) {
          const float4 ni = *reinterpret_cast<const float4*>(pi + k4);
          const float4 nj = *reinterpret_cast<const float4*>(pj + k4);
          const float4 qi = *reinterpret_cast<const float4*>(pi + 32 + k4);
          const float4 qj = *reinterpret_cast<const float4*>(pj + 32 + k4);
          const float4 b0 = *reinterpret_cast<const float4*>(att + k4);
          const float4 w1 = *reinterpret_cast<const float4*>(att + 32 + k4);
          const float p0 = qi.x + qj.x + b0.x, p1 = qi.y + qj.y + b0.y;
          const float p2 = qi.z + qj.z + b0.z, p3 = qi.w + qj.w + b0.w;
          ai = fmaf(fmaxf(ni.x + p0, 0.f), w1.x, ai); aj = fmaf(fmaxf(nj.x + p0, 0.f), w1.x, aj);
          ai = fmaf(fmaxf(ni.y + p1, 0.f), w1.y, ai); aj = fmaf(fmaxf(nj.y + p1, 0.f), w1.y, aj);
          ai = fmaf(fmaxf(ni.z + p2, 0.f), w1.z, ai); aj = fmaf(fmaxf(nj.z + p2, 0.f), w1.z, aj);
          ai = fmaf(fmaxf(ni.w + p3, 0.f), w1.w, ai); aj = fmaf(fmaxf(nj.w + p3, 0.f), w1.w, aj);
        }
        // softmax over ALL N nodes of (a * H); self loops carry incidence 2 (:124,:135-137)
        const float b1v = att[64];
        if (i == j) {
          const float si = 2.f * (ai + b1v);
          const float mx = (N > 1) ? fmaxf(si, 0.f) : si;
          const float ei = __expf(si - mx);
          wi = __fdividef(2.f * ei, ei + static_cast<float>(N - 1) * __expf(-mx));
          wj = 0.f;
        } else {
          const float si = ai + b1v, sj = aj + b1v;
          float mx = fmaxf(si, sj);
          if (N > 2) mx = fmaxf(mx, 0.f);
          const float ei = __expf(si - mx), ej = __expf(sj - mx);
          const float inv = __fdividef(1.f, ei + ej + static_cast<float>(N - 2) * __expf(-mx));
          wi = ei * inv; wj = ej * inv;
        }
      }
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        const float4 u0 = *reinterpret_cast<const float4*>(xi + 8 * k8);
        const float4 u1 = *reinterpret_cast<const float4*>(xi + 8 * k8 + 4);
        const float4 v0 = *reinterpret_cast<const float4*>(xj + 8 * k8);
        const float4 v1 = *reinterpret_cast<const float4*>(xj + 8 * k8 + 4);
        uint4 pk = make_uint4(
            pack_bf16_fast(fmaf(wi, u0.x, wj * v0.x), fmaf(wi, u0.y, wj * v0.y)),
            pack_bf16_fast(fmaf(wi, u0.z, wj * v0.z), fmaf(wi, u0.w, wj * v0.w)),
            pack_bf16_fast(fmaf(wi, u1.x, wj * v1.x), fmaf(wi, u1.y, wj * v1.y)),
            pack_bf16_fast(fmaf(wi, u1.z, wj * v1.z), fmaf(wi, u1.w, wj * v1.w)));
        *reinterpret_cast<uint4*>(sA0 + canon_off(row, k8, 128)) = pk;
      }
    } else {
      const float* src = a.edges + static_cast<size_t>(live ? grow : 0) * 64;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        float4 x = ldg_stream_f4(src + 8 * k8), y = ldg_stream_f4(src + 8 * k8 + 4);
        uint4 pk = make_uint4(pack_bf16_fast(x.x, x.y), pack_bf16_fast(x.z, x.w),
                              pack_bf16_fast(y.x, y.y), pack_bf16_fast(y.z, y.w));
        *reinterpret_cast<uint4*>(sA0 + canon_off(row, k8, 128)) = pk;
      }
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    chain_group_bar(grp);
    if (PAIR) prefetch_nodes(tile + static_cast<long long>(gridDim.x) * 2);   // overlaps the whole MMA chain
    GN_TRACE(1);

    // ---- G1: 64 -> 128 ----
    if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
      fence_after_thread_sync();
      if (elect_one()) {
        issue_bias(tmem_grp, sbase + OFF_ONES, sbase + OFF_BB1, 128);
        issue_gemm(tmem_grp, sA0_addr, sbase + OFF_W1, 128, 64, true);
        mma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase); phase ^= 1;
    GN_TRACE(2);
    fence_after_thread_sync();
    drain_to_smem<true, 4>(tmem_row, sA1, row);
    fence_proxy_async_smem();
    fence_before_thread_sync();
    chain_group_bar(grp);

    // ---- G2: 128 -> 64 (z) ----
    if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
      fence_after_thread_sync();
      if (elect_one()) {
        issue_bias(tmem_grp, sbase + OFF_ONES, sbase + OFF_BB2, 64);
        issue_gemm(tmem_grp, sA1_addr, sbase + OFF_W2, 64, 128, true);
        mma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase); phase ^= 1;
    GN_TRACE(4);
    fence_after_thread_sync();
    drain_to_smem<false, 2>(tmem_row, sA0, row);
    fence_proxy_async_smem();
    fence_before_thread_sync();
    chain_group_bar(grp);

    // ---- G3: 64 -> 256 ([distribution | factor] hidden), accumulator drained in two halves ----
    if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
      fence_after_thread_sync();
      if (elect_one()) {
        issue_bias(tmem_grp, sbase + OFF_ONES, sbase + OFF_BB3, 256);
        issue_gemm(tmem_grp, sA0_addr, sbase + OFF_W3, 256, 64, true);
        mma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase); phase ^= 1;
    GN_TRACE(6);
    fence_after_thread_sync();
    drain_to_smem<true, 4>(tmem_row, sA1, row);
    fence_proxy_async_smem();
    fence_before_thread_sync();
    chain_group_bar(grp);
    // G4a: bias + k in [0,128) -> TMEM columns [0,16) (already drained)
    if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
      fence_after_thread_sync();
      if (elect_one()) {
        issue_bias(tmem_grp, sbase + OFF_ONES, sbase + OFF_BB4, 16);
        issue_gemm(tmem_grp, sA1_addr, sbase + OFF_W4, 16, 128, true);
        mma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase); phase ^= 1;
    GN_TRACE(8);       // A1 is free again
    fence_after_thread_sync();
    drain_to_smem<true, 4>(tmem_row + 128, sA1, row);
    fence_proxy_async_smem();
    fence_before_thread_sync();
    chain_group_bar(grp);
    // G4b: k in [128,256), accumulate
    if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
      fence_after_thread_sync();
      if (elect_one()) {
        issue_gemm(tmem_grp, sA1_addr, sbase + OFF_W4 + 16 * (16 * 16), 16, 128, true);
        mma_commit(mbar);
      }
      __syncwarp();
    }
    mbar_wait(mbar, phase); phase ^= 1;
    GN_TRACE(10);
    fence_after_thread_sync();

    // ---- epilogue 4: T logits | factor logit (biases already inside) ----
    {
      float v[16];
      tmem_ld16(tmem_row, v);
      if (live) {
        float u[TU];
        if (a.noise_mode == GN_NOISE_GIVEN) {
#pragma unroll
          for (int t = 0; t < TU; ++t) u[t] = (t < T) ? __ldg(a.U + static_cast<size_t>(grow) * T + t) : 0.5f;
        } else {
          // one Philox block yields 4 consecutive elements of the (B,E,T) noise tensor
          const unsigned long long el0 =
              (static_cast<unsigned long long>(a.scene_offset) * a.E + static_cast<unsigned long long>(grow)) * T;
          const unsigned long long blk0 = el0 >> 2;
          const int lead = static_cast<int>(el0 & 3);
          // device-resident seed (graph replay): read where it is used, so the other modes keep their registers
          const unsigned long long seed = (a.noise_mode == GN_NOISE_PHILOX_DEVICE_SEED)
                                              ? __ldg(reinterpret_cast<const unsigned long long*>(a.U)) : a.seed;
#pragma unroll
          for (int t = 0; t < TU; ++t) u[t] = 0.5f;
#pragma unroll
          for (int bi = 0; bi < (TU + 3 + 3) / 4; ++bi) {
            if (bi * 4 < lead + T) {
              const uint4 r = Philox::block(blk0 + bi, static_cast<uint32_t>(a.stage_index), seed);
              const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int s = 0; s < 4; ++s) {
                const int t = bi * 4 + s - lead;
                const float uu = static_cast<float>(w[s] >> 8) * (1.0f / 16777216.0f);
#pragma unroll
                for (int tt = 0; tt < TU; ++tt)
                  if (tt >= bi * 4 - 3 && tt <= bi * 4 + 3 && tt == t) u[tt] = uu;
              }
            }
          }
        }
        // y = (logit + g) / tau, g = -log(eps - log(u + eps)), tau = 1/2; fast intrinsics (bf16 path)
        float y[TU];
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < TU; ++t)
          if (t < T) {
            const float g = -__logf(1e-10f - __logf(u[t] + 1e-10f));
            y[t] = 2.f * (v[t] + g);
            mx = fmaxf(mx, y[t]);
          }
        float den = 0.f;
#pragma unroll
        for (int t = 0; t < TU; ++t)
          if (t < T) { y[t] = __expf(y[t] - mx); den += y[t]; }
        float fl = 0.f;
#pragma unroll
        for (int o = 0; o < GN_SMALL_OUT; ++o)
          if (o == T) fl = v[o];
        const float inv = __fdividef(1.f, den);
        const float factor = __fdividef(1.f, 1.f + __expf(-fl));
        float* ef = a.edge_feat + static_cast<size_t>(grow) * T;
#pragma unroll
        for (int t = 0; t < TU; ++t)
          if (t < T) {
            const float d = y[t] * inv;
            if (a.dist_out != nullptr) a.dist_out[static_cast<size_t>(grow) * T + t] = d;
            ef[t] = factor * d;
          }
      }
    }
    GN_TRACE(11);
    // the next tile's G1 overwrites this group's TMEM columns and A0
    fence_before_thread_sync();
    chain_group_bar(grp);
  }

  fence_before_thread_sync();
  __syncthreads();
  if ((tid >> 5) == 0) {
    fence_after_thread_sync();
    tmem_dealloc(*tmem_slot, 512);
  }
}

template <bool PAIR, int TT, bool ALIGNED = false>
static int launch_chain(const TcChainArgs& a, int grid, const char* name, cudaStream_t st) {
  auto kern = edge_chain_tc_kernel<PAIR, TT, ALIGNED>;
  const uint32_t smem = PAIR ? tcmlp::SMEM_BYTES_PAIR : tcmlp::SMEM_BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  {
    ProfScope ps__(name, st);
    kern<<<grid, GN_THREADS, smem, st>>>(a);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

// A 128-row tile of the pairwise layer spans at most floor(127/E)+2 scenes; the fused node2edge
// path needs their nodes to fit the shared-memory staging buffer.
bool edge_chain_pair_fits(int N) {
  const int E = N * N;
  if (E >= 128) return N <= tcmlp::MAXN;             // scene-aligned tiles: one scene per tile
  return (127 / E + 2) * N <= tcmlp::MAXN;
}

unsigned long long* g_trace_buffer = nullptr;

int launch_edge_chain_tc(bool pair, const float* edges, const float* xprime, const float* pq,
                         int N, int E, int T, long long R, const gn_stage_weights* w,
                         const float* U, int noise_mode, unsigned long long seed, long long scene_offset,
                         int stage_index, float* dist_out, float* edge_feat, cudaStream_t st) {
  if (!w->tc_init_w0 || !w->tc_init_w1 || !w->tc_df_w0 || !w->tc_df_w1) return GN_E_NULL;
  TcChainArgs a;
  a.w1 = w->tc_init_w0; a.w2 = w->tc_init_w1; a.w3 = w->tc_df_w0; a.w4 = w->tc_df_w1;
  a.b1 = w->init_b0; a.b2 = w->init_b1; a.b3 = w->df_b0; a.b4 = w->df_b1;
  a.att_b0 = w->att_b0; a.att_w1 = w->att_w1; a.att_b1 = w->att_b1;
  a.edges = edges; a.xprime = xprime; a.pq = pq;
  a.N = N; a.E = E; a.T = T; a.R = R;
  a.U = U; a.noise_mode = noise_mode; a.seed = seed; a.scene_offset = scene_offset; a.stage_index = stage_index;
  a.dist_out = dist_out; a.edge_feat = edge_feat;
  a.trace = g_trace_buffer;
  a.tps = (pair && E >= 128) ? (E + 127) / 128 : 0;
  long long ntiles = a.tps ? (R / E) * a.tps : (R + 127) / 128, want = (ntiles + 1) / 2;
  int grid = want < GN_SM_COUNT ? static_cast<int>(want) : GN_SM_COUNT;
  if (pair) {
    if (a.tps) {
      if (T == 6) return launch_chain<true, 6, true>(a, grid, "edge_chain_pair_tc", st);
      return launch_chain<true, 0, true>(a, grid, "edge_chain_pair_tc", st);
    }
    if (T == 6) return launch_chain<true, 6>(a, grid, "edge_chain_pair_tc", st);
    return launch_chain<true, 0>(a, grid, "edge_chain_pair_tc", st);
  }
  if (T == 10) return launch_chain<false, 10>(a, grid, "edge_chain_tc", st);
  return launch_chain<false, 0>(a, grid, "edge_chain_tc", st);
}

}  // namespace gn
