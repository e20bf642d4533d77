// Fused edge2node of the pairwise layer on the fp32-grade tensor-core path (precision GN_TF32X3).
//
// Reference (model/MS_HGNN_batch.py:259-268 with H = rel_rec + rel_send, :116-120):
//   edges_e = h_i + h_j;  ef_e = sum_t edge_feat[e,t] * (W1_t relu(W0_t edges_e + b0_t) + b1_t);  agg_n = sum_e H[e,n] ef_e
// Collapsed form (both Linears act on the N node rows of a scene instead of its N^2 edges):
//   P_n[t]  = W0_t h_n + b0_t / 2                                        GEMM 1, 3xTF32 on tcgen05
//   s[n][j][t] = edge_feat[(n,j),t] + edge_feat[(j,n),t]                 symmetric; the self loop carries 2 ef (H = 2, :124)
//   G_n[t]  = sum_j s[n][j][t] * relu(P_n[t] + P_j[t])                   fp32 SIMT, symmetric pairs evaluated once
//   agg_n   = sum_t W1_t G_n[t] + (sum_j s[n][j][t]) b1_t                GEMM 2, 3xTF32; partial sums meet in fp32 registers
//
// It replaces agg_in_tf32 + edge2node_pair + agg_out_tf32 (P and G: 2 x 2.2 GB written and read back per NBA step).
//
// One persistent CTA per SM; a tile = SC = floor(128 / N) whole scenes = SC * N <= 128 node rows; the work of a tile is
// cut into U = 2T unit steps u = (t, half): 64 of the 128 hidden columns of agg_mlp[t].  Roles:
//   warps 0-11  "scene" warps (warp = scene, lane = 2 columns): the N rows of the scene's P tile sit in registers, every
//               unordered pair (i <= j) is evaluated once and added to G_i and G_j, G overwrites P in place (a (scene,
//               column pair) block of the tile belongs to exactly one thread).  This is the FMA-pipe-bound phase.
//   warps 12-15 "drain" warps, one per TMEM lane quarter (thread = tile row): WHILE the scene warps work on unit step u
//               they split G_{u-1} into tf32 hi | lo and write it to TENSOR MEMORY, where GEMM 2 reads it as its A
//               operand (TS mode), and they drain P_{u+1} (+ b0 / 2) from its accumulator into the other P tile
//               (double-buffered by unit-step parity); one 512-thread barrier per unit step.
//   all 16      stage h (hi | lo, once per tile, feeds all 2T GEMM 1s), build the symmetric weight table, and add the
//               output partial of unit step u - 2 into 16 fp32 registers each (thread = (row, 16-column slice)).
//   warp 16     weight producer: TMA bulk copies of the host-packed stream (one [64 x 64] hi | lo chunk per GEMM, in
//               issue order) into a ring of 32 KB stages.
//   warp 17     MMA issuer: GEMM 1 of step j + 2 goes out when accumulator j & 1 has been drained (p_free), GEMM 2 of
//               step j when G_j is in tensor memory (g_ready); every GEMM 2 writes a fresh accumulator (24 MMAs) —
//               the tensor core's own accumulation truncates, so long chains in one accumulator cost accuracy.
// TMEM columns: h hi|lo 0..127 | P_u 128..255 (2 x 64) | G_u hi|lo 256..383 | agg partial 384..511 (2 x 64).
//
// Bound: SIMT issue of the relu-sum, 4 instructions per (unordered pair, column): N(N+1)/2 * 128 * T * 4 per scene;
// algorithmic HBM bytes per scene: N*D*4 (h) + N*N*T*4 (edge_feat) + N*D*4 (agg).  Weights: 64 KB per unit step from L2.
#include <cstdlib>
#include <cstring>
#include "gn_tc.cuh"
#include "gn_tf32.cuh"
#include "gn_stage.h"

namespace gn {

namespace pat {
constexpr int ROW_THREADS = 512, THREADS = ROW_THREADS + 64;
constexpr int PLD = 68;                       // padded P / G row (floats): rows 4 banks apart, 128-bit row access conflict-free
constexpr uint32_t STAGE_BYTES = 32768;       // one [64 x 64] chunk, hi then lo
constexpr int NSTAGE = 3;
constexpr int SCENE_WARPS = 12;                // warps 0-11: relu-sum (warp = scene); warps 12-15: drain warps, one per TMEM lane quarter
constexpr int MAXN = 11;
constexpr uint32_t TM_A = 0, TM_P = 128, TM_G = 256, TM_AGG = 384;

struct Bars {
  uint64_t full[NSTAGE], empty[NSTAGE];
  uint64_t a_ready, g_ready, p_free, p_ready[2], agg_ready[2];
  uint32_t tmem_slot, pad;
};

struct Args {
  const float* h;            // (B*N, 64)
  const float* edge_feat;    // (B, N*N, T)
  const unsigned char* wstream;
  const float* b0;           // (T*128)
  const float* b1;           // (T, 64)
  float* agg;                // (B*N, 64)
  int B, N, T, SC, NP4;      // NP4 = N(N+1)/2 rounded up to 4
  uint32_t off_p, off_sym, off_s, off_b0, off_b1, off_pair, off_bar;
  uint32_t off_raw;           // edge_feat staging of the next tile (0: does not fit, the s table is built from global memory)
  unsigned long long* trace;  // optional (gn_profile_set_trace): clock64 stamps of block 0's first TR_TILES tiles, thread 0
};
// trace slots per tile: [0] start | [1] h staged | [2] s table written | [3] s barrier | [4] S done |
// per unit step u at 8 + 8u: start, P ready, R1 end, barrier, relu-sum end, barrier, R2 end | [120..124] epilogue
constexpr int TR_TILES = 6, TR_SLOTS = 128;

__device__ __forceinline__ void row_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// The relu-sum of one scene and one unit step: lane owns columns (2 lane, 2 lane + 1) of the 64-column half.
// Ps: the scene's N rows of the P tile (overwritten with G); sym: s[pair] for this (scene, t), pairs in (i <= j) order.
// packed fp32x2 arithmetic (FADD2 / FFMA2): one issue slot of the FMA pipe per two columns
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int N>
__device__ __forceinline__ void relu_sum(float* Ps, const float* sym, int lane) {
  unsigned long long p[N], g[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    p[i] = *reinterpret_cast<const unsigned long long*>(Ps + i * PLD + 2 * lane);
    g[i] = 0ull;
  }
  const float4* s4 = reinterpret_cast<const float4*>(sym);
  float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
  int idx = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int j = i; j < N; ++j, ++idx) {
      if ((idx & 3) == 0) cur = s4[idx >> 2];
      const float w = (idx & 3) == 0 ? cur.x : (idx & 3) == 1 ? cur.y : (idx & 3) == 2 ? cur.z : cur.w;
      float sx, sy;
      upk2(add2(p[i], p[j]), sx, sy);
      const unsigned long long r = pk2(fmaxf(sx, 0.f), fmaxf(sy, 0.f)), w2 = pk2(w, w);
      g[i] = fma2(w2, r, g[i]);
      if (j != i) g[j] = fma2(w2, r, g[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) *reinterpret_cast<unsigned long long*>(Ps + i * PLD + 2 * lane) = g[i];
}

__device__ __forceinline__ void relu_sum_n(int N, float* Ps, const float* sym, int lane) {
  switch (N) {
    case 1: relu_sum<1>(Ps, sym, lane); break;
    case 2: relu_sum<2>(Ps, sym, lane); break;
    case 3: relu_sum<3>(Ps, sym, lane); break;
    case 4: relu_sum<4>(Ps, sym, lane); break;
    case 5: relu_sum<5>(Ps, sym, lane); break;
    case 6: relu_sum<6>(Ps, sym, lane); break;
    case 7: relu_sum<7>(Ps, sym, lane); break;
    case 8: relu_sum<8>(Ps, sym, lane); break;
    case 9: relu_sum<9>(Ps, sym, lane); break;
    case 10: relu_sum<10>(Ps, sym, lane); break;
    case 11: relu_sum<11>(Ps, sym, lane); break;
    default: relu_sum<11>(Ps, sym, lane); break;
  }
}

// 18 warps: 5 on one scheduler partition (16 K registers each) caps a thread at 96 registers
// TRACE: the clock64 phase stamps (profiles/trace_pair_agg_tf32.py) are compiled in
template <bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
pair_agg_tf32_kernel(const __grid_constant__ Args a) {
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Bars* bars = reinterpret_cast<Bars*>(smem + a.off_bar);
  float* sP = reinterpret_cast<float*>(smem + a.off_p);
  float* sSym = reinterpret_cast<float*>(smem + a.off_sym);
  float* sS = reinterpret_cast<float*>(smem + a.off_s);          // [128][16] row sums of s per t
  float* sB0 = reinterpret_cast<float*>(smem + a.off_b0);        // b0 / 2
  float* sB1 = reinterpret_cast<float*>(smem + a.off_b1);
  unsigned char* sPair = smem + a.off_pair;                      // pair index -> (i, j)
  unsigned char* sPidx = sPair + 2 * (MAXN * (MAXN + 1) / 2);    // (i, j) -> pair index
  const int N = a.N, T = a.T, SC = a.SC, E = N * N, U = 2 * T;
  const int NPu = N * (N + 1) / 2, NP4 = a.NP4;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    mbar_init(&bars->a_ready, ROW_THREADS);
    mbar_init(&bars->g_ready, ROW_THREADS - SCENE_WARPS * 32); mbar_init(&bars->p_free, ROW_THREADS - SCENE_WARPS * 32);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->p_ready[i], 1); mbar_init(&bars->agg_ready[i], 1); }
  }
  if (warp == 0) tmem_alloc(&bars->tmem_slot, 512);
  for (int i = tid; i < T * 128; i += THREADS) sB0[i] = 0.5f * __ldg(a.b0 + i);
  for (int i = tid; i < T * 64; i += THREADS) sB1[i] = __ldg(a.b1 + i);
  if (tid < N) {                                                 // row i of the pair table
    int pr = tid * N - tid * (tid - 1) / 2;
    for (int j = tid; j < N; ++j, ++pr) {
      sPair[2 * pr] = static_cast<unsigned char>(tid); sPair[2 * pr + 1] = static_cast<unsigned char>(j);
      sPidx[tid * N + j] = sPidx[j * N + tid] = static_cast<unsigned char>(pr);
    }
  }
  for (int i = tid; i < SC * T * NP4; i += THREADS) sSym[i] = 0.f;   // the padding entries stay zero
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem = bars->tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  const int ntiles = (a.B + SC - 1) / SC;
  pdl_trigger();                         // programmatic dependent launch, see gn_common.cuh

  if (warp == ROW_THREADS / 32) {
    // ------------------------------------------------------------------ weight producer (warp-uniform, elect.sync)
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const unsigned char* src = a.wstream;
      for (int c = 0; c < 2 * U; ++c) {
        mbar_wait(&bars->empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&bars->full[s], STAGE_BYTES);
          bulk_g2s(sbase + s * STAGE_BYTES, src, STAGE_BYTES, &bars->full[s]);
        }
        __syncwarp();
        src += STAGE_BYTES;
        if (++s == NSTAGE) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == ROW_THREADS / 32 + 1) {
    // ------------------------------------------------------------------ MMA issuer
    int s = 0; uint32_t ph = 0, apar = 0, gpar = 0, fpar = 0;
    auto gemm = [&](uint32_t d, uint32_t a_hi, uint64_t* done) {
      mbar_wait(&bars->full[s], ph);
      fence_after_thread_sync();
      if (elect_one()) {
        const uint32_t b_hi = sbase + s * STAGE_BYTES;
        tf::issue_x3_ts(d, a_hi, a_hi + 64, b_hi, b_hi + STAGE_BYTES / 2, 64, 64, false);
        mma_commit(&bars->empty[s]);
        mma_commit(done);
      }
      __syncwarp();
      if (++s == NSTAGE) { s = 0; ph ^= 1u; }
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(&bars->a_ready, apar); apar ^= 1u;
      fence_after_thread_sync();
      gemm(tmem + TM_P, tmem + TM_A, &bars->p_ready[0]);
      gemm(tmem + TM_P + 64, tmem + TM_A, &bars->p_ready[1]);
      // GEMM 1 of unit step j + 2 goes out as soon as the drain warps have read accumulator j & 1 (p_free), GEMM 2 of unit
      // step j as soon as G_j is in tensor memory (g_ready); the drain warps signal in exactly this order
      mbar_wait(&bars->p_free, fpar); fpar ^= 1u;
      fence_after_thread_sync();
      if (U > 2) gemm(tmem + TM_P, tmem + TM_A, &bars->p_ready[0]);
      for (int st = 0; st < U; ++st) {
        if (st >= 1) {
          mbar_wait(&bars->g_ready, gpar); gpar ^= 1u;
          fence_after_thread_sync();
          gemm(tmem + TM_AGG + ((st - 1) & 1) * 64, tmem + TM_G, &bars->agg_ready[(st - 1) & 1]);
        }
        if (st + 1 < U) {
          mbar_wait(&bars->p_free, fpar); fpar ^= 1u;
          fence_after_thread_sync();
          if (st + 3 < U) gemm(tmem + TM_P + ((st + 3) & 1) * 64, tmem + TM_A, &bars->p_ready[(st + 3) & 1]);
        }
      }
      mbar_wait(&bars->g_ready, gpar); gpar ^= 1u;
      fence_after_thread_sync();
      gemm(tmem + TM_AGG + ((U - 1) & 1) * 64, tmem + TM_G, &bars->agg_ready[(U - 1) & 1]);
    }
  } else {
    // ------------------------------------------------------------------ row / scene threads
    const int q = warp & 3, sl = warp >> 2;
    const int row = q * 32 + lane;
    const int rsc = row / N, ri = row - rsc * N;           // scene in tile / node in scene of this row
    const uint32_t tmem_row = tmem + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t php = 0u, pha = 0u;             // phase bits of p_ready[b] / agg_ready[b] (bit b)
    const bool use_raw = a.off_raw != 0;
    unsigned char* sRaw = smem + a.off_raw;
    pdl_wait();                          // h and edge_feat are the predecessors' outputs
    // edge_feat of a tile -> shared memory (cp.async, 16-byte chunks from the aligned-down tile base)
    auto prefetch_raw = [&](int tile) {
      if (use_raw && tile < ntiles) {
        const int nsn = min(SC, a.B - tile * SC);
        const size_t byte0 = static_cast<size_t>(tile) * SC * E * T * 4;
        const uint32_t mis = static_cast<uint32_t>(byte0 & 15);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.edge_feat) + (byte0 - mis);
        const int nchunk = static_cast<int>((static_cast<uint32_t>(nsn) * E * T * 4 + mis + 15) >> 4);
        for (int i = tid; i < nchunk; i += ROW_THREADS) cp_async16(sRaw + 16 * i, src + 16 * i);
      }
      cp_async_commit();
    };
    auto load_h = [&](int tile, float4 (&x)[4]) {
      const bool lv = tile < ntiles && row < min(SC, a.B - tile * SC) * N;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        x[k4] = lv ? ldg_f4(a.h + (static_cast<size_t>(tile) * SC * N + row) * 64 + 16 * sl + 4 * k4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 x[4];
    prefetch_raw(blockIdx.x);
    load_h(blockIdx.x, x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int b0s = tile * SC;
      const int ns = min(SC, a.B - b0s);
      const int nv = ns * N;
      const size_t grow0 = static_cast<size_t>(b0s) * N;
      const bool live = row < nv;
      const int titer = (tile - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x);
      const bool tr = TRACE && a.trace != nullptr && blockIdx.x == 0 && tid == 0 && titer < TR_TILES;
      unsigned long long* trp = a.trace + titer * TR_SLOTS;
      if (tr) trp[0] = clock64();

      // ---- prologue: h slice -> tensor memory (hi | lo); symmetric weights s of the tile -> shared memory
      {
        uint32_t hv[16], lv[16];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          tf::split_tf32(x[k4].x, hv[4 * k4], lv[4 * k4]); tf::split_tf32(x[k4].y, hv[4 * k4 + 1], lv[4 * k4 + 1]);
          tf::split_tf32(x[k4].z, hv[4 * k4 + 2], lv[4 * k4 + 2]); tf::split_tf32(x[k4].w, hv[4 * k4 + 3], lv[4 * k4 + 3]);
        }
        tf::tmem_st16(tmem_row + TM_A + 16 * sl, hv);
        tf::tmem_st16(tmem_row + TM_A + 64 + 16 * sl, lv);
        tf::tmem_st_wait();
        fence_before_thread_sync();
        mbar_arrive(&bars->a_ready);
      }
      if (tr) trp[1] = clock64();
      if (use_raw) {
        cp_async_wait<0>();
        row_bar();                                        // every thread's copies of the tile's edge_feat have landed
        const float* rf = reinterpret_cast<const float*>(sRaw) +
                          ((static_cast<size_t>(tile) * SC * E * T * 4) & 15) / 4;
        for (int idx = tid; idx < ns * NPu; idx += ROW_THREADS) {
          const int s = idx / NPu, pr = idx - s * NPu;
          const int i = sPair[2 * pr], j = sPair[2 * pr + 1];
          const float* e0 = rf + (s * E + i * N + j) * T;
          const float* e1 = rf + (s * E + j * N + i) * T;
          float* dst = sSym + s * T * NP4 + pr;
          for (int t = 0; t < T; ++t) dst[t * NP4] = e0[t] + e1[t];
        }
      } else {
        const float* ef = a.edge_feat + static_cast<size_t>(b0s) * E * T;
        const int total = ns * NPu * T;
#pragma unroll 4
        for (int idx = tid; idx < total; idx += ROW_THREADS) {
          const int rest = idx / T, t = idx - rest * T;
          const int s = rest / NPu, pr = rest - s * NPu;
          const int i = sPair[2 * pr], j = sPair[2 * pr + 1];
          const float* e = ef + static_cast<size_t>(s) * E * T + t;
          sSym[(s * T + t) * NP4 + pr] = __ldg(e + (i * N + j) * T) + __ldg(e + (j * N + i) * T);
        }
      }
      if (tr) trp[2] = clock64();
      row_bar();                                          // the tile's s table is complete, the raw copy is consumed
      if (tr) trp[3] = clock64();
      prefetch_raw(tile + gridDim.x);
      if (live) {
        // S[row][t] = sum_j s[row][j][t]  (rank-T bias of the output; the (i, i) entry already holds 2 ef);
        // runs while GEMM 1 of the first unit step is on the tensor core
        for (int t = sl; t < T; t += 4) {
          const float* sy = sSym + (rsc * T + t) * NP4;
          float sum = 0.f;
#pragma unroll 4
          for (int j = 0; j < N; ++j) sum += sy[sPidx[ri * N + j]];      // unrolled: the loads of four terms in flight
          sS[row * 16 + t] = sum;
        }
      }
      if (tr) trp[4] = clock64();

      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      // one output partial (GEMM 2 of unit step k, accumulator k & 1) -> this thread's 16 fp32 sums
      auto add_partial = [&](int k) {
        const int b = k & 1;
        float v[16];
        mbar_wait(&bars->agg_ready[b], (pha >> b) & 1u); pha ^= 1u << b;
        fence_after_thread_sync();
        tmem_ld16(tmem_row + TM_AGG + b * 64 + 16 * sl, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += v[j];
      };
      // drain warps: P_u (+ b0 / 2) from its accumulator -> P tile u & 1 in shared memory; the accumulator is then free
      auto r1 = [&](int u) {
        const int b = u & 1, t = u >> 1;
        mbar_wait(&bars->p_ready[b], (php >> b) & 1u); php ^= 1u << b;
        fence_after_thread_sync();
        const float* bb = sB0 + t * 128 + b * 64;
        float* dst = sP + b * (128 * PLD) + row * PLD;
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          uint32_t ra[16], rb[16];
          tf::tmem_ld16_nowait(tmem_row + TM_P + b * 64 + 16 * c, ra);
          tf::tmem_ld16_nowait(tmem_row + TM_P + b * 64 + 16 * c + 16, rb);
          tmem_ld_wait();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 bv = *reinterpret_cast<const float4*>(bb + 16 * c + 4 * k4);
            *reinterpret_cast<float4*>(dst + 16 * c + 4 * k4) =
                make_float4(__uint_as_float(ra[4 * k4]) + bv.x, __uint_as_float(ra[4 * k4 + 1]) + bv.y,
                            __uint_as_float(ra[4 * k4 + 2]) + bv.z, __uint_as_float(ra[4 * k4 + 3]) + bv.w);
          }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 bv = *reinterpret_cast<const float4*>(bb + 16 * c + 16 + 4 * k4);
            *reinterpret_cast<float4*>(dst + 16 * c + 16 + 4 * k4) =
                make_float4(__uint_as_float(rb[4 * k4]) + bv.x, __uint_as_float(rb[4 * k4 + 1]) + bv.y,
                            __uint_as_float(rb[4 * k4 + 2]) + bv.z, __uint_as_float(rb[4 * k4 + 3]) + bv.w);
          }
        }
        fence_before_thread_sync();
        mbar_arrive(&bars->p_free);
      };
      // drain warps: G_u (P tile u & 1, overwritten by the relu-sum) -> tf32 hi | lo in tensor memory
      auto r2 = [&](int u) {
        const float* src = sP + (u & 1) * (128 * PLD) + row * PLD;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t hv[16], lv[16];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 g = *reinterpret_cast<const float4*>(src + 16 * c + 4 * k4);
            tf::split_tf32(g.x, hv[4 * k4], lv[4 * k4]); tf::split_tf32(g.y, hv[4 * k4 + 1], lv[4 * k4 + 1]);
            tf::split_tf32(g.z, hv[4 * k4 + 2], lv[4 * k4 + 2]); tf::split_tf32(g.w, hv[4 * k4 + 3], lv[4 * k4 + 3]);
          }
          tf::tmem_st16(tmem_row + TM_G + 16 * c, hv);
          tf::tmem_st16(tmem_row + TM_G + 64 + 16 * c, lv);
        }
        tf::tmem_st_wait();
        fence_before_thread_sync();
        mbar_arrive(&bars->g_ready);
      };
      const bool is_drain = warp >= SCENE_WARPS;
      const bool trd = TRACE && a.trace != nullptr && blockIdx.x == 0 && tid == SCENE_WARPS * 32 && titer < TR_TILES;
      if (is_drain) r1(0);
      row_bar();
      for (int u = 0; u < U; ++u) {
        const int t = u >> 1;
        unsigned long long* tru = trp + 8 + 8 * (u < 14 ? u : 13);
        if (tr) tru[0] = clock64();
        // every thread: the output partial of unit step u - 2 (its GEMM 2 had a whole unit step to finish; it also
        // guarantees that G of step u - 1 may overwrite the operand GEMM 2 of step u - 2 read)
        if (u >= 2) add_partial(u - 2);
        if (tr) tru[1] = clock64();
        if (!is_drain) {
          // scene warps: symmetric relu-sum of unit step u, G overwrites P in place
          for (int s = warp; s < ns; s += SCENE_WARPS)
            relu_sum_n(N, sP + (u & 1) * (128 * PLD) + s * N * PLD, sSym + (s * T + t) * NP4, lane);
          if (tr) tru[2] = clock64();
        } else {
          // drain warps, meanwhile: G of the previous unit step -> tensor memory, P of the next one -> shared memory
          if (u >= 1) r2(u - 1);
          if (trd) tru[4] = clock64();
          if (u + 1 < U) r1(u + 1);
          if (trd) tru[5] = clock64();
        }
        row_bar();
        if (tr) tru[3] = clock64();
      }
      // ---- tail and epilogue: last two partials, rank-T bias, store; the next tile's h rows are requested first
      {
        if (tr) trp[120] = clock64();
        load_h(tile + gridDim.x, x);          // the next tile's h rows: their DRAM latency hides behind the epilogue
        if (U >= 2) add_partial(U - 2);
        if (is_drain) r2(U - 1);
        if (tr) trp[121] = clock64();
        add_partial(U - 1);
        if (live) {
          for (int t = 0; t < T; ++t) {
            const float st = sS[row * 16 + t];
            const float* b1 = sB1 + t * 64 + 16 * sl;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const float4 bv = *reinterpret_cast<const float4*>(b1 + 4 * k4);
              acc[4 * k4] = fmaf(st, bv.x, acc[4 * k4]); acc[4 * k4 + 1] = fmaf(st, bv.y, acc[4 * k4 + 1]);
              acc[4 * k4 + 2] = fmaf(st, bv.z, acc[4 * k4 + 2]); acc[4 * k4 + 3] = fmaf(st, bv.w, acc[4 * k4 + 3]);
            }
          }
          if (tr) trp[122] = clock64();
          float* dst = a.agg + (grow0 + row) * 64 + 16 * sl;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            *reinterpret_cast<float4*>(dst + 4 * k4) = make_float4(acc[4 * k4], acc[4 * k4 + 1], acc[4 * k4 + 2], acc[4 * k4 + 3]);
        }
        if (tr) trp[123] = clock64();
        fence_before_thread_sync();
        row_bar();                                        // S / s table / P tiles are rewritten by the next tile
        if (tr) trp[124] = clock64();
      }
    }
    cp_async_wait<0>();
  }

  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_thread_sync();
    tmem_dealloc(tmem, 512);
  }
}

static uint32_t align_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

}  // namespace pat
extern unsigned long long* g_trace_buffer;      // gn_profile_set_trace (gn_edge_mlp_tc.cu)

bool pair_agg_tf32_fits(int N, int D, int T) {
  return D == 64 && N >= 1 && N <= pat::MAXN && T >= 1 && T <= 15;
}

int launch_pair_agg_tf32(const float* h, const float* edge_feat, int B, int N, int T,
                         const gn_stage_weights* w, float* agg, cudaStream_t st) {
  using namespace pat;
  if (!w->tf_pagg_w) return GN_E_NULL;
  if (B <= 0) return GN_OK;
  if (reinterpret_cast<uintptr_t>(w->tf_pagg_w) & 15) return GN_E_ALIGN;
  Args a;
  a.h = h; a.edge_feat = edge_feat; a.wstream = static_cast<const unsigned char*>(w->tf_pagg_w);
  a.b0 = w->agg_b0; a.b1 = w->agg_b1; a.agg = agg;
  a.B = B; a.N = N; a.T = T; a.SC = 128 / N;
  a.NP4 = (N * (N + 1) / 2 + 3) & ~3;
  a.trace = nullptr;
  if (g_trace_buffer != nullptr) {               // profiles/trace_pair_agg_tf32.py
    const char* only = getenv("GN_TRACE_KERNEL");
    if (only != nullptr && strcmp(only, "pair_agg_tf32") == 0) a.trace = g_trace_buffer;
  }
  uint32_t o = NSTAGE * STAGE_BYTES;
  a.off_p = o; o += 2 * 128 * PLD * 4;             // P / G tile, double-buffered by unit-step parity
  a.off_sym = o; o += static_cast<uint32_t>(a.SC) * T * a.NP4 * 4;
  a.off_s = o; o += 128 * 16 * 4;
  a.off_b0 = o; o += static_cast<uint32_t>(T) * 128 * 4;
  a.off_b1 = o; o += static_cast<uint32_t>(T) * 64 * 4;
  a.off_pair = o; o += align_up(2 * (MAXN * (MAXN + 1) / 2) + MAXN * MAXN, 16);
  a.off_bar = align_up(o, 16);
  uint32_t smem = a.off_bar + static_cast<uint32_t>(sizeof(Bars));
  if (smem > 227 * 1024) return GN_E_SHAPE;
  const uint32_t raw_bytes = align_up(static_cast<uint32_t>(a.SC) * N * N * T * 4 + 16, 16);
  a.off_raw = 0;
  if (align_up(smem, 16) + raw_bytes <= 227 * 1024 && (reinterpret_cast<uintptr_t>(edge_feat) & 15) == 0) {
    a.off_raw = align_up(smem, 16);
    smem = a.off_raw + raw_bytes;
  }
  void (*kern)(Args) = a.trace != nullptr ? pair_agg_tf32_kernel<true> : pair_agg_tf32_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int ntiles = (B + a.SC - 1) / a.SC;
  const int grid = ntiles < GN_SM_COUNT ? ntiles : GN_SM_COUNT;
  {
    ProfScope ps__("pair_agg_tf32", st);
    cudaError_t le = launch_pdl(kern, dim3(grid), dim3(THREADS), smem, st, a);
    if (le != cudaSuccess) return static_cast<int>(le);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
