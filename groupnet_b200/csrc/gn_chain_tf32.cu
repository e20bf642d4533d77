// fp32-grade tensor-core path (precision GN_TF32X3): the dense contractions of one message-passing stage as
// chains of 3xTF32 GEMMs on tcgen05 / TMEM, 1e-5 parity with the reference's fp32 addmm chain
// (model/MS_HGNN_batch.py:31-53 MLP_dict_softmax, :201-229 MLP, :247-268 edge_aggregation).
//
// ONE engine kernel runs every chain.  A tile = 128 rows (node rows or edge rows), one persistent CTA per SM:
//
//   warps 0-3  row threads: thread t owns tile row t = TMEM lane t.  They stage the tile's input rows (fp32
//              from HBM, or the fused pairwise node2edge) as split hi / lo tf32 operands in shared memory, and
//              after every GEMM drain the accumulator (tcgen05.ld), apply bias / ReLU / per-row scale in fp32,
//              and either store fp32 rows to HBM or split the activation again and write it back to TENSOR
//              MEMORY (tcgen05.st) where the next GEMM reads it as its A operand (TS mode): activations never
//              touch shared memory or HBM between the Linears of a chain;
//   warp 4     weight producer: TMA bulk copies (cp.async.bulk + mbarrier complete_tx) of the host-packed
//              weight stream (hi | lo chunks in consumption order) into a ring of 16 KB stages;
//   warp 5     MMA issuer: one thread, three kind::tf32 MMAs per K = 8 step (gn_tf32.cuh), tcgen05.commit
//              releases ring stages and publishes accumulators.
//
// The chain itself is data: a list of ops (gn_chain_tf32.cuh) built on the host by the launchers below —
// TMEM column assignment, which drains feed which op, where the issuer must wait for the row threads.
// Row threads and the issuer communicate through two rings of NBAR mbarriers (a_ready: operands written,
// acc_ready: accumulator complete); validate_program() replays the protocol on the host and rejects a program
// whose producers could run NBAR phases ahead of its consumers.
//
// Roofline: tensor pipe at 1/6 of the bf16 rate (tf32 = 1/2, three MMAs per product): 64*N/2 clk per 128 x N x 8.
#include <type_traits>
#include <cstdlib>
#include <cstring>
#include "gn_chain_tf32.cuh"
#include "gn_tf32.cuh"
#include "gn_stage.h"

namespace gn {

namespace tfe {

constexpr int ROW_THREADS = 128 * NSLICE;          // 16 row warps
constexpr int THREADS = ROW_THREADS + 64;          // + weight producer warp + MMA issuer warp

__device__ __forceinline__ void row_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

struct Bars {
  uint64_t full[8], empty[8], a_ready[NBAR], acc_ready[NBAR];
  uint64_t node_full, node_empty;        // ST_PAIR node block: filled by the producer warp's bulk copies, freed by the row threads
  uint32_t tmem_slot, pad;
};

// One pass of a thread's share of a drain: 16 accumulator columns [c0, c0 + 16) of its row (already loaded into r) ->
// fp32 epilogue -> HBM and / or the next op's A operand in tensor memory (hi | lo split).  16 columns per pass keep
// the 512 row threads inside their 96-register budget (a spill costs an L2 round trip here: the shared-memory
// carve-out leaves no L1).  The epilogue is specialised at compile time on (kind, ReLU, per-row scale / rank-T bias):
// an interpreted version cost ~2.5x the instructions.  rsb: this row's per-row scales, staged in shared memory.
template <int KIND, bool RELU, bool RS>
__device__ __forceinline__ void drain_pass(const Args& a, const Op& op, const float* aux, const float* rsb, uint32_t (&r)[16],
                                           uint32_t tmem_row, int c0, long long grow, bool live, float& carry) {
  constexpr int W = 16;
  // partial sums kept in separate accumulators (short accumulation chains in the tensor core) meet in fp32 here
  if (KIND == DR_STORE && RS) {
    for (int pi = 1; pi < op.nsum; ++pi) {
      uint32_t q[W];
      tf::tmem_ld16_nowait(tmem_row + op.acc_col + pi * op.sum_stride + c0, q);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < W; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(q[j]));
    }
  }
  if (op.bias_off >= 0) {
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(aux + op.bias_off + c0 + 4 * q);
      tf::upk2u(tf::add2(tf::pk2u(r[4 * q], r[4 * q + 1]), b.x), r[4 * q], r[4 * q + 1]);
      tf::upk2u(tf::add2(tf::pk2u(r[4 * q + 2], r[4 * q + 3]), b.y), r[4 * q + 2], r[4 * q + 3]);
    }
  }
  if (RELU) {
#pragma unroll
    for (int j = 0; j < W; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.f));
  }
  if (RS && KIND != DR_STORE) {                 // per-row scale (edge_feat_t)
    const float scale = rsb[op.rs_idx];
    const unsigned long long s2 = tf::pk2f(scale, scale);
#pragma unroll
    for (int j = 0; j < W; j += 2) tf::upk2u(tf::mul2(tf::pk2u(r[j], r[j + 1]), s2), r[j], r[j + 1]);
  }
  if (KIND == DR_DOT) {
    // even | odd columns in the two lanes of a packed accumulator; the lanes meet at the end of the pass
    unsigned long long c2 = tf::pk2f(carry, 0.f);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) {
      const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(aux + a.dot_off + c0 + 4 * q);
      c2 = tf::fma2(tf::pk2u(r[4 * q], r[4 * q + 1]), wv.x, c2);
      c2 = tf::fma2(tf::pk2u(r[4 * q + 2], r[4 * q + 3]), wv.y, c2);
    }
    float ce, co;
    tf::upk2f(c2, ce, co);
    carry = ce + co;
    return;
  }
  if (RS && KIND == DR_STORE) {                 // rank-T bias: v += sum_t rs[row][t] * bm[t][col]
    const float* bm = aux + a.bm_off + op.out_col0 + c0;
    for (int t = 0; t < a.bm_T; ++t) {
      const float st = rsb[t];
      const unsigned long long st2 = tf::pk2f(st, st);
#pragma unroll
      for (int q = 0; q < W / 4; ++q) {
        const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(bm + t * a.bm_ld + 4 * q);
        tf::upk2u(tf::fma2(st2, b.x, tf::pk2u(r[4 * q], r[4 * q + 1])), r[4 * q], r[4 * q + 1]);
        tf::upk2u(tf::fma2(st2, b.y, tf::pk2u(r[4 * q + 2], r[4 * q + 3])), r[4 * q + 2], r[4 * q + 3]);
      }
    }
  }
  if ((KIND == DR_STORE || KIND == DR_TMEM_STORE) && live) {
    float* dst = op.out + grow * op.ldo + op.out_col0 + c0;
#pragma unroll
    for (int q = 0; q < W / 4; ++q)
      *reinterpret_cast<uint4*>(dst + 4 * q) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
  }
  if (KIND == DR_TMEM || KIND == DR_TMEM_STORE) {
    uint32_t lo[W];
#pragma unroll
    for (int j = 0; j < W; ++j) tf::split_tf32(__uint_as_float(r[j]), r[j], lo[j]);
    tf::tmem_st16(tmem_row + op.dst_col + c0, r);
    tf::tmem_st16(tmem_row + op.dst_col + op.dn + c0, lo);
  }
}

// A thread's share of a drain: columns [sl*w, sl*w + w), w = dn / NSLICE = 16 or 32; both 16-column halves of a
// 32-column share are in flight from tensor memory before the first is processed.
template <int KIND, bool RELU, bool RS>
__device__ __forceinline__ void drain_slice(const Args& a, const Op& op, const float* aux, const float* rsb,
                                            uint32_t tmem_row, int sl, long long grow, bool live, float& carry) {
  const int w = op.dn / NSLICE;
  uint32_t r0[16];
  tf::tmem_ld16_nowait(tmem_row + op.drain_col + sl * w, r0);
  if (w == 32) {
    uint32_t r1[16];
    tf::tmem_ld16_nowait(tmem_row + op.drain_col + sl * w + 16, r1);
    tc::tmem_ld_wait();
    drain_pass<KIND, RELU, RS>(a, op, aux, rsb, r0, tmem_row, sl * w, grow, live, carry);
    drain_pass<KIND, RELU, RS>(a, op, aux, rsb, r1, tmem_row, sl * w + 16, grow, live, carry);
  } else {
    tc::tmem_ld_wait();
    drain_pass<KIND, RELU, RS>(a, op, aux, rsb, r0, tmem_row, sl * w, grow, live, carry);
  }
  if (KIND == DR_TMEM || KIND == DR_TMEM_STORE) tf::tmem_st_wait();
}

// x / d with r = 1 / d correctly rounded: one Newton correction gives the correctly rounded quotient (d is an agent
// count: no overflow / denormal corner cases)
__device__ __forceinline__ float div_by(float x, float d, float r) {
  const float q0 = x * r;
  return fmaf(fmaf(-q0, d, x), r, q0);
}

// DR_DOTG: this thread's share of the distribution head: lg[t] = sum over its 32 hidden units k of
// relu(acc[k] + b[k]) * W[k][t], W = [128][TP] fp32 in the smem constants (rows of TP logits, zero padded; TP is the
// even number >= T among 6, 8, 10, 12, 16 — the reference's T = 6 (pairwise) and 10 (hyper) exactly; two hidden units
// share TP / 2 16-byte reads)
template <int TP>
__device__ __forceinline__ void dot_logits(const Op& op, const float* aux, const float* w4, uint32_t tmem_row, int sl,
                                           float (&lg)[TP]) {
  static_assert(TP % 2 == 0 && TP <= 16, "rows of two hidden units are read as float4s");
  unsigned long long acc[TP / 2];          // logits (2q, 2q + 1) in the lanes of packed accumulator q
#pragma unroll
  for (int q = 0; q < TP / 2; ++q) acc[q] = 0ull;
  uint32_t r0[16], r1[16];
  tf::tmem_ld16_nowait(tmem_row + op.acc_col + 32 * sl, r0);
  tf::tmem_ld16_nowait(tmem_row + op.acc_col + 32 * sl + 16, r1);
  tc::tmem_ld_wait();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int c0 = 32 * sl + 16 * half;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const unsigned long long b2 = *reinterpret_cast<const unsigned long long*>(aux + op.bias_off + c0 + j);
      float v0, v1;
      tf::upk2f(tf::add2(half ? tf::pk2u(r1[j], r1[j + 1]) : tf::pk2u(r0[j], r0[j + 1]), b2), v0, v1);
      v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f);
      const unsigned long long v0p = tf::pk2f(v0, v0), v1p = tf::pk2f(v1, v1);
      const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(w4 + (c0 + j) * TP);
      unsigned long long wf[TP];             // TP packed pairs: unit j's TP / 2, then unit j + 1's
#pragma unroll
      for (int q = 0; q < TP / 2; ++q) { const ulonglong2 wv = wr[q]; wf[2 * q] = wv.x; wf[2 * q + 1] = wv.y; }
#pragma unroll
      for (int q = 0; q < TP / 2; ++q) acc[q] = tf::fma2(v0p, wf[q], acc[q]);
#pragma unroll
      for (int q = 0; q < TP / 2; ++q) acc[q] = tf::fma2(v1p, wf[TP / 2 + q], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < TP / 2; ++q) tf::upk2f(acc[q], lg[2 * q], lg[2 * q + 1]);
}

// drain variants the launchers use (Op::variant)
enum { DV_TMEM = 0, DV_TMEM_RELU, DV_TMEM_RELU_RS, DV_TMEM_STORE, DV_STORE, DV_STORE_BM, DV_DOT_RELU, DV_COUNT };

// PAIR: the fused pairwise node2edge staging (ST_PAIR programs) is compiled in; the ST_ROWS instance carries none of its
// code or register state (one kernel for both cost the node chains ~10 %)
// VM: bit v = drain variant v is compiled in, bit DV_COUNT = the DR_DOTG drain is (the launcher picks the smallest
// instantiated superset of the program's variants: the row warps' instruction footprint is what a program can reach)
// TRACE: the clock64 phase stamps (profiles/trace_tf32.py) are compiled in; the production instances carry none of them
template <bool PAIR, uint32_t VM, bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
chain_tf32_kernel(const __grid_constant__ Args a) {
  using namespace tc;
  // the Gumbel noise belongs to the DR_DOTG drain, the per-row scales to the RS variants: other instances carry neither
  constexpr bool HAS_DOTG = (VM >> DV_COUNT) & 1u;
  constexpr bool HAS_RS = (VM & ((1u << DV_TMEM_RELU_RS) | (1u << DV_STORE_BM))) != 0;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Bars* bars = reinterpret_cast<Bars*>(smem + a.off_bar);

  if (tid == 0) {
    for (int s = 0; s < a.nstage; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int i = 0; i < NBAR; ++i) { mbar_init(&bars->a_ready[i], ROW_THREADS); mbar_init(&bars->acc_ready[i], 1); }
    mbar_init(&bars->node_full, 1); mbar_init(&bars->node_empty, ROW_THREADS);
  }
  if (warp == 0) tmem_alloc(&bars->tmem_slot, 512);
  float* aux = reinterpret_cast<float*>(smem + a.off_aux);
  for (int i = 0; i < a.naux; ++i)
    for (int j = tid; j < a.aux_n[i]; j += THREADS) aux[a.aux_off[i] + j] = __ldg(a.aux_src[i] + j);
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem = bars->tmem_slot;
  const uint32_t sbase = smem_u32(smem);
  pdl_trigger();                         // the next kernel's CTAs may take over SMs as this grid's CTAs retire

  // The producer and issuer warps run their loops with ALL 32 lanes (warp-uniform control flow and addresses) and
  // predicate only the asynchronous instructions with elect.sync: tcgen05.mma / tcgen05.commit / cp.async.bulk take
  // their operands from the uniform datapath, and issued from a divergent single-lane branch each one costs
  // 160-290 clk of election loop and register->uniform transfers (profiles/probes/mma_probe.cu) instead of ~10.
  if (warp == ROW_THREADS / 32) {
    // ------------------------------------------------------------------ weight producer
    int s = 0; uint32_t ph = 0;
    int ci = 0;
    bool tr = false;
    unsigned long long* trp = a.trace;
    auto stream_op = [&](const Op& op) {
      const uint32_t bytes = static_cast<uint32_t>(op.N) * op.kc * 8;
      const int nch = op.K / op.kc;
      const unsigned char* src = a.wstream + op.w_off;
      for (int c = 0; c < nch; ++c, ++ci) {
        mbar_wait(&bars->empty[s], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&bars->full[s], bytes);
          bulk_g2s(sbase + a.off_ring + s * a.stage_bytes, src, bytes, &bars->full[s]);
        }
        __syncwarp();
        if (tr && ci < TR_MAXCH) trp[4 * ci] = clock64();
        src += bytes;
        if (++s == a.nstage) { s = 0; ph ^= 1u; }
      }
    };
    // ST_PAIR: the node block of a tile (pq and Y rows of the nodes its <= 128 edges touch) also comes in as bulk copies
    // from this warp — one per node row, the rows are padded in shared memory — so the row threads spend no LSU
    // instructions on it (as 16-byte cp.async from the row threads it cost ~1.7 K clk of each tile's ~16 K).  Block k + 1
    // may overwrite block k once every row thread has finished part B of block k (node_empty).
    int nb = 0;                             // next node block to issue (the CTA's nb-th tile)
    auto node_block = [&]() {
      const long long t = static_cast<long long>(blockIdx.x) + static_cast<long long>(nb) * gridDim.x;
      if (t < a.ntiles) {
        if (nb > 0) mbar_wait(&bars->node_empty, static_cast<uint32_t>(nb - 1) & 1u);
        const uint32_t uE = static_cast<uint32_t>(a.E), uN = static_cast<uint32_t>(a.N), ut = static_cast<uint32_t>(t);
        uint32_t node0, cnt;
        if (a.tps) { node0 = ut / static_cast<uint32_t>(a.tps) * uN; cnt = uN; }
        else {
          const uint32_t r_last = min(ut * 128u + 127u, static_cast<uint32_t>(a.R) - 1u);
          const uint32_t b_lo = ut * 128u / uE, b_hi = r_last / uE;
          node0 = b_lo * uN; cnt = (b_hi - b_lo + 1u) * uN;
        }
        const uint32_t np_s = sbase + a.off_node, ny_s = np_s + MAXN * NLD * 4;
        if (elect_one()) mbar_expect_tx(&bars->node_full, cnt * (64u + 128u) * 4u);
        __syncwarp();
        for (uint32_t n = 0; n < cnt; ++n) {
          if (elect_one()) {
            bulk_g2s(np_s + n * (NLD * 4), a.pq + static_cast<size_t>(node0 + n) * 64, 64 * 4, &bars->node_full);
            bulk_g2s(ny_s + n * (YLD * 4), a.ypre + static_cast<size_t>(node0 + n) * 128, 128 * 4, &bars->node_full);
          }
          __syncwarp();
        }
      }
      ++nb;
    };
    int titer = 0;
    if (PAIR) {
      pdl_wait();                           // pq / Y are the previous kernel's output
      node_block();
    }
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++titer) {
      tr = TRACE && a.trace != nullptr && blockIdx.x == 0 && titer < TR_TILES && lane == 0;
      trp = a.trace + titer * TR_SLOTS + TR_CHUNK;
      ci = 0;
      const bool last = tile + gridDim.x >= a.ntiles;
      // o == -1: the pro_op of a CTA's first tile (one call site keeps the lambda inlined and its state in registers)
      for (int o = (titer == 0 && a.pro_op >= 0) ? -1 : 0; o < a.nops; ++o) {
        if (last && o == a.skip_last_op) continue;
        stream_op(a.ops[o < 0 ? a.pro_op : o]);
        // a pipelined program consumes block k + 1 inside tile k: the CTA's second block follows the pro_op's weights
        if (PAIR && o < 0) node_block();
      }
      // here, not earlier: the wait for the row threads' part B must not hold back this tile's weights
      if (PAIR) node_block();
    }
  } else if (warp == ROW_THREADS / 32 + 1) {
    // ------------------------------------------------------------------ MMA issuer
    int s = 0; uint32_t ph = 0;
    uint32_t aw = 0, sg = 0;               // a_ready phases consumed, acc_ready phases produced
    int ci = 0;
    bool tr = false;
    unsigned long long* trp = a.trace;
    auto issue_op = [&](const Op& op, int o) {
      // weights first (they landed long ago: the producer runs a stage ahead), so that the MMAs go out as soon as the
      // row threads publish the operands
      mbar_wait(&bars->full[s], ph);
      for (int w = 0; w < op.wait_n; ++w) { mbar_wait(&bars->a_ready[aw & (NBAR - 1)], (aw / NBAR) & 1u); ++aw; }
      fence_after_thread_sync();
      if (tr) trp[3 * o] = clock64();
      const int N = op.N, K = op.K, kc = op.kc, nch = K / kc;
      const uint32_t d = tmem + op.acc_col;
      const uint32_t half = static_cast<uint32_t>(N) * kc * 4;
      for (int c = 0; c < nch; ++c, ++ci) {
        if (tr && ci < TR_MAXCH) trp[TR_CHUNK + 4 * ci + 1] = clock64();
        if (c > 0) { mbar_wait(&bars->full[s], ph); fence_after_thread_sync(); }
        if (tr && c == 0) trp[3 * o + 1] = clock64();
        if (tr && ci < TR_MAXCH) trp[TR_CHUNK + 4 * ci + 2] = clock64();
        const uint32_t b_hi = sbase + a.off_ring + s * a.stage_bytes, b_lo = b_hi + half;
        const bool acc_first = (c > 0) || (op.accumulate != 0);
        if (elect_one()) {
          if (op.a_src == A_SMEM) {
            const uint32_t a_hi = sbase + a.off_a0 + op.a_buf * a.a0_buf_bytes + static_cast<uint32_t>(c * kc / 4) * 2048u;
            tf::issue_x3_ss(d, a_hi, a_hi + a.a0_half_bytes, b_hi, b_lo, N, kc, acc_first);
          } else {
            const uint32_t ta = tmem + op.a_col + c * kc;
            tf::issue_x3_ts(d, ta, ta + K, b_hi, b_lo, N, kc, acc_first);
          }
          mma_commit(&bars->empty[s]);
        }
        __syncwarp();
        if (tr && ci < TR_MAXCH) trp[TR_CHUNK + 4 * ci + 3] = clock64();
        if (++s == a.nstage) { s = 0; ph ^= 1u; }
      }
      if (op.signal) {
        if (elect_one()) mma_commit(&bars->acc_ready[sg & (NBAR - 1)]);
        __syncwarp();
        ++sg;
      }
      if (tr) trp[3 * o + 2] = clock64();
    };
    // software-pipelined programs issue one op of a CTA's FIRST tile up front (pro_op) and skip the op that belongs to
    // the tile after the LAST one (skip_last_op)
    int titer = 0;
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++titer) {
      tr = TRACE && a.trace != nullptr && blockIdx.x == 0 && titer < TR_TILES && lane == 0;
      trp = a.trace + titer * TR_SLOTS;
      ci = 0;
      const bool last = tile + gridDim.x >= a.ntiles;
      for (int o = (titer == 0 && a.pro_op >= 0) ? -1 : 0; o < a.nops; ++o) {     // o == -1: the pro_op
        if (last && o == a.skip_last_op) continue;
        issue_op(a.ops[o < 0 ? a.pro_op : o], o < 0 ? MAX_OPS - 1 : o);
      }
    }
  } else {
    // ------------------------------------------------------------------ row threads (stage + drain)
    // 16 warps: warp w serves TMEM lane quarter q = w % 4 (the only lanes it may access) and column slice
    // sl = w / 4: every tile row is handled by NSLICE threads, each owning a quarter of the columns of every
    // staging pass and drain (4 warps per scheduler instead of 1 hide the LDTM / LDS / MUFU latencies)
    const int q = warp & 3, sl = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t tmem_row = tmem + (static_cast<uint32_t>(q * 32) << 16);
    // everything above (and the weight producer, which reads only packed weights) ran under the predecessor's tail;
    // the activations it wrote are read from here on
    pdl_wait();
    unsigned char* a0 = smem + a.off_a0;
    float* part = reinterpret_cast<float*>(smem + a.off_scr);    // [NSLICE][128][2] attention-logit partials
    float* dotp = part + NSLICE * 128 * 2;                       // [NSLICE][128]    DR_DOT partials
    float* ybuf = dotp + NSLICE * 128;                           // [128][17]        Gumbel-softmax logits
    uint32_t ar = 0, sg = 0;                 // a_ready phases produced, acc_ready phases consumed
    constexpr bool pair = PAIR;
    float* np = reinterpret_cast<float*>(smem + a.off_node);     // pq rows of the tile's node block   [MAXN][NLD]
    float* ny = np + MAXN * NLD;                                 // Y = x' W^T rows                    [MAXN][YLD]
    // pair mode: R < 2^31 (checked by the launcher), so the scene / edge index math is 32-bit unsigned
    const uint32_t uE = static_cast<uint32_t>(a.E > 0 ? a.E : 1), uN = static_cast<uint32_t>(a.N);
    const uint32_t total_nodes = pair ? static_cast<uint32_t>(a.R) / uE * uN : 0u;
    const int tps = pair ? a.tps : 0;
    unsigned long long seed = a.seed;
    if (HAS_DOTG && a.noise_mode == GN_NOISE_PHILOX_DEVICE_SEED && a.U != nullptr)
      seed = __ldg(reinterpret_cast<const unsigned long long*>(a.U));

    // ST_ROWS staging of one tile row: fp32 input columns -> split hi | lo -> canonical smem operand -> arrive on a_ready
    auto stage_rows = [&](const Op& op, long long g_row, bool lv) {
      unsigned char* hi = a0 + op.a_buf * a.a0_buf_bytes;
      unsigned char* lo = hi + a.a0_half_bytes;
      const int K = op.K, k0 = op.st_k0;
      const bool div = a.a_div != 0.f;
      const float rdiv = div ? 1.f / a.a_div : 0.f;
      // a batch of four 16-byte loads per thread goes out before the first one is consumed (a load -> split -> store loop
      // exposed one DRAM latency per iteration); K <= 64 is one batch, K = 128 two (eight in flight cost 32 registers of
      // the 96 and spilled the loop state of the row threads)
      constexpr int BATCH = 4;
#pragma unroll 1
      for (int i0 = 0; NSLICE * i0 < (K >> 2); i0 += BATCH) {
        float4 xs[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          const int k4 = sl + NSLICE * (i0 + i), k = k0 + 4 * k4;
          xs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (lv && k4 < (K >> 2))
            xs[i] = ldg_f4((k < a.k_src0) ? a.src0 + g_row * a.ld0 + k : a.src1 + g_row * a.ld1 + (k - a.k_src0));
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
          const int k4 = sl + NSLICE * (i0 + i);
          if (k4 < (K >> 2)) {
            float4 x = xs[i];
            if (div) {                       // x / d as q0 = x r, q = q0 + (x - q0 d) r: the IEEE quotient without the
              x.x = div_by(x.x, a.a_div, rdiv); x.y = div_by(x.y, a.a_div, rdiv);   // ~30-instruction div.rn sequence
              x.z = div_by(x.z, a.a_div, rdiv); x.w = div_by(x.w, a.a_div, rdiv);   // (32 per thread and tile)
            }
            uint4 h4, l4;
            tf::split_tf32(x.x, h4.x, l4.x); tf::split_tf32(x.y, h4.y, l4.y);
            tf::split_tf32(x.z, h4.z, l4.z); tf::split_tf32(x.w, h4.w, l4.w);
            *reinterpret_cast<uint4*>(hi + tf::canon_off32(row, k4, 128)) = h4;
            *reinterpret_cast<uint4*>(lo + tf::canon_off32(row, k4, 128)) = l4;
          }
        }
      }
      fence_proxy_async_smem();
      fence_before_thread_sync();
      mbar_arrive(&bars->a_ready[ar & (NBAR - 1)]); ++ar;
    };
    // programs with EV_STAGE_NEXT stage tile i + 1 inside tile i (after the last MMA that reads the staged buffer):
    // the input rows' DRAM latency hides behind the rest of the chain; the first tile is staged here
    if (a.stage_first >= 0 && static_cast<long long>(blockIdx.x) < a.ntiles) {
      const long long g0 = static_cast<long long>(blockIdx.x) * 128 + row;
      stage_rows(a.ops[a.stage_first], g0, g0 < a.R);
    }

    // ---- fused node2edge of the pairwise layer, in two parts so that a software-pipelined program can run them for
    // tile i + 1 inside the MMA waits of tile i.  Part A (shared memory only): attention over the (<= 2) members of edge
    // (i,j), softmax over ALL N nodes (non-members enter with logit 0), self loops carry incidence 2 (:124,:135-137);
    // slice sl sums 8 of the 32 attention hidden units, the partial logits meet in shared memory.  Its results — the two
    // weights and the node-block rows of the edge — stay in registers for part B.
    float s_wi = 0.f, s_wj = 0.f;
    int s_li = 0, s_lj = 0;
    uint32_t nk = 0;                       // node blocks consumed
    auto pair_stage_a = [&](long long t, bool tr, unsigned long long* trp) {
      const long long tsc = tps ? static_cast<long long>(static_cast<uint32_t>(t) / static_cast<uint32_t>(tps)) : 0;
      const int tch = tps ? static_cast<int>(t - tsc * tps) : 0;
      const long long g_row = tps ? tsc * a.E + tch * 128 + row : t * 128 + row;
      const bool lv = g_row < a.R && (!tps || tch * 128 + row < a.E);
      mbar_wait(&bars->node_full, nk & 1u);                // the producer warp's bulk copies of this tile's node block
      if (tr) trp[TR_STAGE - TR_ROWS + 0] = clock64();
      const uint32_t b_lo = tps ? static_cast<uint32_t>(tsc) : (static_cast<uint32_t>(t) * 128u) / uE;
      const int N = a.N;
      const float* att = aux + a.att_off;                  // b0[32] | w1[32] | b1
      int li = 0, lj = 0, i = 0, j = 0;
      if (lv) {
        const uint32_t b = static_cast<uint32_t>(g_row) / uE;
        const uint32_t eidx = static_cast<uint32_t>(g_row) - b * uE;
        i = static_cast<int>(eidx / uN); j = static_cast<int>(eidx) - i * N;
        li = static_cast<int>(b - b_lo) * N + i; lj = static_cast<int>(b - b_lo) * N + j;
        const float* pi = np + li * NLD;
        const float* pj = np + lj * NLD;
        float ai = 0.f, aj = 0.f;
#pragma unroll
        for (int k4 = 8 * sl; k4 < 8 * sl + 8; k4 += 4) {
          const ulonglong2 ni = *reinterpret_cast<const ulonglong2*>(pi + k4);
          const ulonglong2 nj = *reinterpret_cast<const ulonglong2*>(pj + k4);
          const ulonglong2 qi = *reinterpret_cast<const ulonglong2*>(pi + 32 + k4);
          const ulonglong2 qj = *reinterpret_cast<const ulonglong2*>(pj + 32 + k4);
          const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(att + k4);
          const float4 w1 = *reinterpret_cast<const float4*>(att + 32 + k4);
          const unsigned long long p01 = tf::add2(tf::add2(qi.x, qj.x), b0.x), p23 = tf::add2(tf::add2(qi.y, qj.y), b0.y);
          float i0, i1, i2, i3, j0, j1, j2, j3;
          tf::upk2f(tf::add2(ni.x, p01), i0, i1); tf::upk2f(tf::add2(ni.y, p23), i2, i3);
          tf::upk2f(tf::add2(nj.x, p01), j0, j1); tf::upk2f(tf::add2(nj.y, p23), j2, j3);
          ai = fmaf(fmaxf(i0, 0.f), w1.x, ai); aj = fmaf(fmaxf(j0, 0.f), w1.x, aj);
          ai = fmaf(fmaxf(i1, 0.f), w1.y, ai); aj = fmaf(fmaxf(j1, 0.f), w1.y, aj);
          ai = fmaf(fmaxf(i2, 0.f), w1.z, ai); aj = fmaf(fmaxf(j2, 0.f), w1.z, aj);
          ai = fmaf(fmaxf(i3, 0.f), w1.w, ai); aj = fmaf(fmaxf(j3, 0.f), w1.w, aj);
        }
        *reinterpret_cast<float2*>(part + (sl * 128 + row) * 2) = make_float2(ai, aj);
      }
      if (tr) trp[TR_STAGE - TR_ROWS + 1] = clock64();
      row_bar();
      if (tr) trp[TR_STAGE - TR_ROWS + 2] = clock64();
      float wi = 0.f, wj = 0.f;
      if (lv) {
        float ai = 0.f, aj = 0.f;
#pragma unroll
        for (int s4 = 0; s4 < NSLICE; ++s4) {          // fixed order: every slice gets the same sums
          const float2 p = *reinterpret_cast<const float2*>(part + (s4 * 128 + row) * 2);
          ai += p.x; aj += p.y;
        }
        const float b1v = att[64];
        if (i == j) {
          const float si = 2.f * (ai + b1v);
          const float mx = (N > 1) ? fmaxf(si, 0.f) : si;
          const float ei = expf(si - mx);
          wi = ei * __frcp_rn(ei + static_cast<float>(N - 1) * expf(-mx)) * 2.f;   // e * rcp(d): within 1 ulp of e / d
          wj = 0.f;
        } else {
          const float si = ai + b1v, sj = aj + b1v;
          float mx = fmaxf(si, sj);
          if (N > 2) mx = fmaxf(mx, 0.f);
          const float ei = expf(si - mx), ej = expf(sj - mx);
          const float rden = __frcp_rn(ei + ej + static_cast<float>(N - 2) * expf(-mx));
          wi = ei * rden; wj = ej * rden;
        }
      }
      s_wi = wi; s_wj = wj; s_li = li; s_lj = lj;
      if (tr) trp[TR_STAGE - TR_ROWS + 3] = clock64();
    };
    // Part B: hidden = relu(w_i Y_i + w_j Y_j + b), this slice's 32 of the 128 columns -> tensor memory (hi | lo) as the A
    // operand of `op`; then the node block is refilled for the tile after `t`.
    auto pair_stage_b = [&](const Op& op, long long t, bool tr, unsigned long long* trp) {
      const unsigned long long wi2 = tf::pk2f(s_wi, s_wi), wj2 = tf::pk2f(s_wj, s_wj);
      const float* yi = ny + s_li * YLD + 32 * sl;
      const float* yj = ny + s_lj * YLD + 32 * sl;
      const float* yb = aux + a.yb_off + 32 * sl;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t hv[16], lv[16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(yi + 16 * half + 4 * q4);
          const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(yj + 16 * half + 4 * q4);
          const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(yb + 16 * half + 4 * q4);
          float x0, x1, x2, x3;
          tf::upk2f(tf::fma2(wi2, u.x, tf::fma2(wj2, v.x, bb.x)), x0, x1);
          tf::upk2f(tf::fma2(wi2, u.y, tf::fma2(wj2, v.y, bb.y)), x2, x3);
          tf::split_tf32(fmaxf(x0, 0.f), hv[4 * q4], lv[4 * q4]);
          tf::split_tf32(fmaxf(x1, 0.f), hv[4 * q4 + 1], lv[4 * q4 + 1]);
          tf::split_tf32(fmaxf(x2, 0.f), hv[4 * q4 + 2], lv[4 * q4 + 2]);
          tf::split_tf32(fmaxf(x3, 0.f), hv[4 * q4 + 3], lv[4 * q4 + 3]);
        }
        tf::tmem_st16(tmem_row + op.a_col + 32 * sl + 16 * half, hv);
        tf::tmem_st16(tmem_row + op.a_col + 128 + 32 * sl + 16 * half, lv);
      }
      tf::tmem_st_wait();
      if (tr) trp[TR_STAGE - TR_ROWS + 4] = clock64();
      fence_before_thread_sync();
      mbar_arrive(&bars->a_ready[ar & (NBAR - 1)]); ++ar;     // the MMAs start while the node block is refilled
      if (tr) trp[TR_STAGE - TR_ROWS + 5] = clock64();
      mbar_arrive(&bars->node_empty); ++nk;                   // this thread is done with the node block
      if (tr) { trp[TR_STAGE - TR_ROWS + 6] = clock64(); trp[TR_STAGE - TR_ROWS + 7] = trp[TR_STAGE - TR_ROWS + 6]; }
    };
    // Gumbel noise of a tile's rows (slice sl: edge types [sl*tq, sl*tq + tq)): g = -log(eps - log(U + eps)) (:446-455)
    // -> ybuf, computed while the row threads would otherwise wait for a GEMM
    auto tile_noise = [&](long long g_row, bool lv) {
      const int T = a.T;
      if (!lv) return;
      if (a.noise_mode == GN_NOISE_GIVEN) {
        const int tq = (T + NSLICE - 1) / NSLICE, t0 = sl * tq;
        constexpr int TQ = (GN_SMALL_OUT - 1 + NSLICE - 1) / NSLICE;
#pragma unroll
        for (int jj = 0; jj < TQ; ++jj) {
          const int t = t0 + jj;
          if (jj < tq && t < T) ybuf[row * 17 + t] = gumbel_from_uniform(__ldg(a.U + static_cast<size_t>(g_row) * T + t));
        }
      } else {
        // Philox: the row's T consecutive elements span <= T / 4 + 2 counter blocks of four; each block is generated ONCE
        // (by slice = block mod 4) instead of once per element — a third of the rounds at T = 6 or 10
        const unsigned long long e0 =
            (static_cast<unsigned long long>(a.scene_offset) * a.E + static_cast<unsigned long long>(g_row)) * T;
        const unsigned long long b_first = e0 >> 2, b_last = (e0 + T - 1) >> 2;
        if (b_last - b_first <= 1) {
          // <= 2 blocks (T <= 6): a thread's critical path is what counts here (the row warps run in lock step, ~8 clk per
          // dependent instruction), so each block is generated by TWO slices and each draws two of its four elements
          const unsigned long long bk = b_first + (sl >> 1);
          if (bk <= b_last) {
            const uint4 blk = Philox::block(bk, static_cast<uint32_t>(a.stage_index), seed);
            const uint32_t w0 = (sl & 1) ? blk.z : blk.x, w1 = (sl & 1) ? blk.w : blk.y;
            const unsigned long long el = bk * 4 + 2 * (sl & 1);
            if (el >= e0 && el < e0 + T)
              ybuf[row * 17 + static_cast<int>(el - e0)] = gumbel_from_uniform(static_cast<float>(w0 >> 8) * (1.0f / 16777216.0f));
            if (el + 1 >= e0 && el + 1 < e0 + T)
              ybuf[row * 17 + static_cast<int>(el + 1 - e0)] = gumbel_from_uniform(static_cast<float>(w1 >> 8) * (1.0f / 16777216.0f));
          }
          return;
        }
        for (unsigned long long bk = b_first + sl; bk <= b_last; bk += NSLICE) {
          const uint4 blk = Philox::block(bk, static_cast<uint32_t>(a.stage_index), seed);
          const uint32_t wv[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const unsigned long long el = bk * 4 + w;
            if (el >= e0 && el < e0 + T)
              ybuf[row * 17 + static_cast<int>(el - e0)] =
                  gumbel_from_uniform(static_cast<float>(wv[w] >> 8) * (1.0f / 16777216.0f));
          }
        }
      }
    };
    // software-pipelined pairwise program: the first tile is staged up front (its G2 is the issuer's pro_op)
    if (a.pro_op >= 0 && static_cast<long long>(blockIdx.x) < a.ntiles) {
      pair_stage_a(blockIdx.x, false, a.trace);
      pair_stage_b(a.ops[a.pro_op], blockIdx.x, false, a.trace);
    }

    int titer = 0;
    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++titer) {
      const bool tr = TRACE && a.trace != nullptr && blockIdx.x == 0 && titer < TR_TILES && tid == 0;
      unsigned long long* trp = a.trace + titer * TR_SLOTS + TR_ROWS;
      const long long tscene = tps ? static_cast<long long>(static_cast<uint32_t>(tile) / static_cast<uint32_t>(tps)) : 0;
      const int tchunk = tps ? static_cast<int>(tile - tscene * tps) : 0;
      const long long grow = tps ? tscene * a.E + tchunk * 128 + row : tile * 128 + row;
      const bool live = grow < a.R && (!tps || tchunk * 128 + row < a.E);
      float carry = 0.f;
      if (HAS_DOTG && (a.pro_op >= 0 || a.stage_first >= 0) && a.edge_feat != nullptr) tile_noise(grow, live);   // ybuf is free: the previous tile ended on a barrier
      const bool has_next = tile + gridDim.x < a.ntiles;

      for (int e = 0; e < a.nev; ++e) {
        const Op& op = a.ops[a.ev_op[e]];
        long long ev_t0 = 0;
        if (tr) { ev_t0 = clock64(); trp[3 * e] = ev_t0; }
        if (a.ev_type[e] == EV_PAIR_A_NEXT || a.ev_type[e] == EV_PAIR_B_NEXT) {
          // the NEXT tile's fused node2edge, inside this tile's MMA waits (the op is skipped by the issuer on a CTA's
          // last tile, and so is the arrival)
          if (has_next) {
            if (a.ev_type[e] == EV_PAIR_A_NEXT) pair_stage_a(tile + gridDim.x, tr, trp);
            else pair_stage_b(op, tile + gridDim.x, tr, trp);
          }
          if (tr) { trp[3 * e + 1] = ev_t0; trp[3 * e + 2] = clock64(); }
          continue;
        }
        if (a.ev_type[e] == EV_STAGE_NEXT) {
          const long long nt = tile + gridDim.x;
          const long long gn = nt * 128 + row;
          stage_rows(op, gn, nt < a.ntiles && gn < a.R);
          if (tr) { trp[3 * e + 1] = ev_t0; trp[3 * e + 2] = clock64(); }
          continue;
        }
        if (a.ev_type[e] == EV_STAGE) {
          if (pair) {
            pair_stage_a(tile, tr, trp);
            pair_stage_b(op, tile, tr, trp);
          } else {
            stage_rows(op, grow, live);
          }
          if (tr) { trp[3 * e + 1] = ev_t0; trp[3 * e + 2] = clock64(); }
          if (HAS_RS && e == 0 && a.rs != nullptr) {
            // this row's per-row scales (edge_feat / S) -> shared memory, once per tile: every slice of the row needs them
            // in every drain, and without an L1 each global read is an L2 round trip
            row_bar();                                    // the previous tile's drains are done with the buffer
            for (int t = sl; t < a.rs_n; t += NSLICE) ybuf[row * 17 + t] = live ? __ldg(a.rs + grow * a.rs_ld + t) : 0.f;
            row_bar();
          }
          if (HAS_DOTG && e == 0 && a.edge_feat != nullptr) tile_noise(grow, live);   // while the first GEMM runs
          continue;
        }

        // ---- EV_DRAIN
        mbar_wait(&bars->acc_ready[sg & (NBAR - 1)], (sg / NBAR) & 1u); ++sg;
        fence_after_thread_sync();
        if (tr) trp[3 * e + 1] = clock64();
        const int kind = op.drain;
        if (((VM >> DV_COUNT) & 1u) && kind == DR_DOTG) {
          // The distribution head (128 -> T) as fp32 dot products inside this drain — no A-operand split, no MMA, no
          // accumulator round trip (the GEMM form cost a 256-column hi | lo store, a K = 128 MMA chain and a handoff) —
          // followed by the Gumbel softmax / sigmoid tail (:45-53).  Slice sl holds 32 of the 128 hidden units: its T
          // partial logits, and the factor head's partial from the DR_DOT drain, meet the other slices' through spare
          // TENSOR MEMORY columns (the row's four threads share TMEM lanes, not a warp).
          // TP logits per row, XW exchange columns per slice (the factor head's partial rides in column XW - 1)
          auto dotg = [&](auto tp_c, auto xw_c) {
            constexpr int TP = decltype(tp_c)::value, XW = decltype(xw_c)::value;
            constexpr int TMAX = TP < XW - 1 ? TP : XW - 1;
            float lg[XW];
            {
              float part[TP];
              dot_logits<TP>(op, aux, aux + a.w4_off, tmem_row, sl, part);
              uint32_t pk[XW];
#pragma unroll
              for (int j = 0; j < XW; ++j) pk[j] = j < TMAX ? __float_as_uint(part[j]) : 0u;
              pk[XW - 1] = __float_as_uint(carry);
              if constexpr (XW == 8) tf::tmem_st8(tmem_row + op.dst_col + XW * sl, pk);
              else tf::tmem_st16(tmem_row + op.dst_col + XW * sl, pk);
              tf::tmem_st_wait();
            }
            fence_before_thread_sync();
            row_bar();
            fence_after_thread_sync();
            {
              // fixed order: every slice gets the same sums
              static_assert(NSLICE == 4, "exchange reads the four slices' partials in two pairs");
              auto ldx = [&](int s4, uint32_t (&q)[XW]) {
                if constexpr (XW == 8) tf::tmem_ld8_nowait(tmem_row + op.dst_col + XW * s4, q);
                else tf::tmem_ld16_nowait(tmem_row + op.dst_col + XW * s4, q);
              };
              uint32_t qa[XW], qb[XW];
              ldx(0, qa); ldx(1, qb);
              if constexpr (XW == 8) {
                uint32_t qc[XW], qd[XW];
                ldx(2, qc); ldx(3, qd);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < XW; ++j)
                  lg[j] = ((__uint_as_float(qa[j]) + __uint_as_float(qb[j])) + __uint_as_float(qc[j])) + __uint_as_float(qd[j]);
              } else {
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < XW; ++j) lg[j] = __uint_as_float(qa[j]) + __uint_as_float(qb[j]);
                ldx(2, qa); ldx(3, qb);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < XW; ++j) lg[j] = (lg[j] + __uint_as_float(qa[j])) + __uint_as_float(qb[j]);
              }
            }
            const int T = a.T, tq = (T + NSLICE - 1) / NSLICE, t0 = sl * tq;
            if (live) {
              // y = (logit + g) / tau with the noise g left in ybuf by the staging event
              float mx = -INFINITY;
#pragma unroll
              for (int t = 0; t < TMAX; ++t) {
                if (t < T) {
                  lg[t] = (lg[t] + aux[a.gb_off + t] + ybuf[row * 17 + t]) / 0.5f;
                  mx = fmaxf(mx, lg[t]);
                }
              }
              // ex2.approx-based exponentials: 2 ulp, three orders of magnitude inside the 1e-5 bound of outputs in [0, 1]
              float den = 0.f;
#pragma unroll
              for (int t = 0; t < TMAX; ++t) {
                if (t < T) { lg[t] = __expf(lg[t] - mx); den += lg[t]; }
              }
              const float factor = __frcp_rn(1.f + __expf(-(lg[XW - 1] + aux[a.gb_off + T])));
              const float rden = __frcp_rn(den);              // e * rcp(den): within 1 ulp of e / den
#pragma unroll
              for (int t = 0; t < TMAX; ++t) {
                if (t >= t0 && t < t0 + tq && t < T) {
                  const float dd = lg[t] * rden;
                  if (a.dist_out != nullptr) a.dist_out[static_cast<size_t>(grow) * T + t] = dd;
                  a.edge_feat[static_cast<size_t>(grow) * T + t] = factor * dd;
                }
              }
            }
          };
          if (a.w4_tp == 6) dotg(std::integral_constant<int, 6>{}, std::integral_constant<int, 8>{});
          else if (a.w4_tp == 10) dotg(std::integral_constant<int, 10>{}, std::integral_constant<int, 16>{});
          else if (a.w4_tp == 8) dotg(std::integral_constant<int, 8>{}, std::integral_constant<int, 16>{});
          else if (a.w4_tp == 12) dotg(std::integral_constant<int, 12>{}, std::integral_constant<int, 16>{});
          else dotg(std::integral_constant<int, 16>{}, std::integral_constant<int, 16>{});
          fence_before_thread_sync();
          row_bar();                              // ybuf and the exchange columns are rewritten by the next tile
        } else if (kind != DR_NONE) {
          const float* rsb = ybuf + row * 17;               // per-row scales staged at tile start (programs with a.rs)
          const int v = op.variant;
#define GN_DV(V, ...) if (((VM >> (V)) & 1u) && v == (V)) drain_slice<__VA_ARGS__>(a, op, aux, rsb, tmem_row, sl, grow, live, carry)
          GN_DV(DV_TMEM, DR_TMEM, false, false);
          else GN_DV(DV_TMEM_RELU, DR_TMEM, true, false);
          else GN_DV(DV_TMEM_RELU_RS, DR_TMEM, true, true);
          else GN_DV(DV_TMEM_STORE, DR_TMEM_STORE, false, false);
          else GN_DV(DV_STORE, DR_STORE, false, false);
          else GN_DV(DV_STORE_BM, DR_STORE, false, true);
          else GN_DV(DV_DOT_RELU, DR_DOT, true, false);
#undef GN_DV
        }
        if (op.arrive) {
          fence_before_thread_sync();
          mbar_arrive(&bars->a_ready[ar & (NBAR - 1)]); ++ar;
        }
        if (tr) trp[3 * e + 2] = clock64();
      }
    }
    if (pair) cp_async_wait<0>();
  }

  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_thread_sync();
    tmem_dealloc(tmem, 512);
  }
}

// ===========================================================================================
// host side: program construction
// ===========================================================================================
struct Builder {
  Args a;
  size_t wbytes = 0;       // bytes of one tile's weight stream
  int a0_K = 0, nbuf = 1;  // staged buffer width / count
  bool node_block = false;
  uint32_t stage_bytes;
  // layout_node: reserve the ST_PAIR node block when sizing the ring (the edge chain has ONE weight stream for its
  // pairwise and hyper forms, so both use the pairwise layout)
  Builder(int a0_K_, int nbuf_, bool node_block_, bool layout_node) : a0_K(a0_K_), nbuf(nbuf_), node_block(node_block_) {
    memset(&a, 0, sizeof(a));
    a.stage_first = -1; a.pro_op = -1; a.skip_last_op = -1;
    stage_bytes = ring_stage_bytes(a0_K, nbuf, layout_node);
  }

  // largest multiple of 8 that divides K and keeps a [N x kc] hi+lo chunk inside one ring stage
  int chunk_k(int N, int K) const {
    int kc = static_cast<int>(stage_bytes / 8) / N;
    kc = kc / 8 * 8;
    if (kc > K) kc = K;
    while (kc > 8 && K % kc) kc -= 8;
    return kc;
  }
  Op& add(int a_src, int a_buf_or_col, int K, int N, int acc_col, int accumulate, int wait_n, int signal) {
    Op& o = a.ops[a.nops++];
    memset(&o, 0, sizeof(o));
    o.a_src = static_cast<short>(a_src);
    if (a_src == A_SMEM) o.a_buf = static_cast<short>(a_buf_or_col); else o.a_col = static_cast<short>(a_buf_or_col);
    o.K = static_cast<short>(K); o.N = static_cast<short>(N); o.kc = static_cast<short>(chunk_k(N, K));
    o.acc_col = static_cast<short>(acc_col); o.drain_col = -1; o.accumulate = static_cast<short>(accumulate);
    o.wait_n = static_cast<short>(wait_n); o.signal = static_cast<short>(signal);
    o.rs_idx = -1; o.bias_off = -1;
    o.w_off = static_cast<int>(wbytes);
    wbytes += static_cast<size_t>(N) * K * 8;
    return o;
  }
  // register `n` floats at `src` as smem constants; returns their float offset, or -1 when the block is full
  int aux_used = 0;
  int aux(const float* src, int n) {
    const int n4 = (n + 3) & ~3;
    if (!src || a.naux >= MAX_AUX || aux_used + n4 > AUX_FLOATS) return -1;
    a.aux_src[a.naux] = src; a.aux_n[a.naux] = static_cast<short>(n); a.aux_off[a.naux] = static_cast<short>(aux_used);
    ++a.naux;
    aux_used += n4;
    return aux_used - n4;
  }
  // bias of a drain: smem constant when it fits, else read from global memory
  void set_bias(Op& o, const float* bias, int n) {
    o.bias_off = -1; o.bias = nullptr;
    if (!bias) return;
    const int off = aux(bias, n);
    if (off >= 0) o.bias_off = static_cast<short>(off); else o.bias = bias;
  }
  void ev(int type, int op) { a.ev_type[a.nev] = static_cast<unsigned char>(type); a.ev_op[a.nev] = static_cast<unsigned char>(op); ++a.nev; }
  void drain_tmem(Op& o, int dn, int relu, const float* bias, int dst_col, int arrive) {
    o.drain = DR_TMEM; o.dn = static_cast<short>(dn); o.relu = static_cast<short>(relu); set_bias(o, bias, dn);
    o.dst_col = static_cast<short>(dst_col); o.arrive = static_cast<short>(arrive);
  }
  void drain_store(Op& o, int dn, int relu, const float* bias, float* out, long long ldo, int col0, int arrive) {
    o.drain = DR_STORE; o.dn = static_cast<short>(dn); o.relu = static_cast<short>(relu); set_bias(o, bias, dn);
    o.out = out; o.ldo = ldo; o.out_col0 = static_cast<short>(col0); o.arrive = static_cast<short>(arrive);
  }
};

// Replays the row-thread / issuer protocol: arrivals must equal waits per tile, signals must equal drains, TMEM
// ranges must fit, and neither side may get NBAR phases ahead of the other (an mbarrier ring would alias).
// the compiled drain variant of an op; -1: the combination is not built
static int drain_variant(const Op& op) {
  const bool rs = op.rs_idx >= 0;
  switch (op.drain) {
    case DR_TMEM: return op.use_bm ? -1 : (rs ? (op.relu ? DV_TMEM_RELU_RS : -1) : (op.relu ? DV_TMEM_RELU : DV_TMEM));
    case DR_TMEM_STORE: return (rs || op.relu || op.use_bm) ? -1 : DV_TMEM_STORE;
    case DR_STORE: return (rs || op.relu) ? -1 : (op.use_bm ? DV_STORE_BM : DV_STORE);
    case DR_DOT: return (rs || !op.relu || op.use_bm) ? -1 : DV_DOT_RELU;
    case DR_DOTG: return (rs || !op.relu || op.use_bm) ? -1 : 0;
    default: return 0;
  }
}

static int validate_program(Args& a) {
  if (a.nops < 1 || a.nops > MAX_OPS || a.nev < 1 || a.nev > MAX_EV) return GN_E_SHAPE;
  if (a.rs != nullptr && (a.edge_feat != nullptr || a.rs_n < 1 || a.rs_n > 16)) return GN_E_SHAPE;   // one smem buffer
  int waits = 0, signals = 0;
  for (int o = 0; o < a.nops; ++o) {
    Op& op = a.ops[o];
    if (op.drain_col < 0) op.drain_col = op.acc_col;        // the default: an op's drain reads its own accumulator
    if (op.N < 16 || op.N > 256 || (op.N & 15) || op.K < 8 || (op.K & 7) || op.kc < 8 || (op.kc & 7) || op.K % op.kc) return GN_E_SHAPE;
    if (static_cast<uint32_t>(op.N) * op.kc * 8 > a.stage_bytes) return GN_E_SHAPE;
    if (op.acc_col < 0 || op.acc_col + op.N > 512) return GN_E_SHAPE;
    if (op.a_src == A_TMEM && (op.a_col < 0 || op.a_col + 2 * op.K > 512)) return GN_E_SHAPE;
    if (op.a_src == A_SMEM && (static_cast<uint32_t>(op.K) * 128 * 4 > a.a0_half_bytes || op.K > 128)) return GN_E_SHAPE;
    if (op.signal) {
      if (op.drain == DR_DOTG) { if (op.dn != 128 || op.dst_col < 0 || op.dst_col + 16 * NSLICE > 512 || a.w4_off < 0) return GN_E_SHAPE; }
      else if (op.drain != DR_NONE && ((op.dn != 64 && op.dn != 128) || (op.drain_col == op.acc_col && op.dn > op.N))) return GN_E_SHAPE;   // NSLICE x 16 / 32
      if (op.drain_col != op.acc_col && (op.drain_col < 0 || op.drain_col + op.dn > 512 || op.nsum > 1 ||
                                         op.drain == DR_DOT || op.drain == DR_DOTG)) return GN_E_SHAPE;
      if (op.nsum > 1 && (op.sum_stride < op.dn || op.acc_col + (op.nsum - 1) * op.sum_stride + op.dn > 512)) return GN_E_SHAPE;
      if ((op.drain == DR_TMEM || op.drain == DR_TMEM_STORE) && (op.dst_col < 0 || op.dst_col + 2 * op.dn > 512)) return GN_E_SHAPE;
      if ((op.drain == DR_STORE || op.drain == DR_TMEM_STORE) &&
          (!op.out || (op.ldo & 3) || (op.out_col0 & 3) || (reinterpret_cast<uintptr_t>(op.out) & 15))) return GN_E_ALIGN;
      if (op.bias != nullptr) return GN_E_SHAPE;            // biases live in the smem constants (Builder::set_bias)
      if (op.nsum > 1 && !(op.drain == DR_STORE && op.use_bm)) return GN_E_SHAPE;
      if (op.use_bm && a.bm_off < 0) return GN_E_SHAPE;
      const int dv = drain_variant(op);
      if (dv < 0) return GN_E_SHAPE;
      a.ops[o].variant = static_cast<short>(dv);
      ++signals;
    } else if (op.arrive) return GN_E_SHAPE;
    waits += op.wait_n;
  }
  int arrivals = 0, drains = 0;
  for (int e = 0; e < a.nev; ++e) {
    const Op& op = a.ops[a.ev_op[e]];
    if (a.ev_type[e] == EV_PAIR_A_NEXT) continue;
    if (a.ev_type[e] == EV_STAGE || a.ev_type[e] == EV_STAGE_NEXT || a.ev_type[e] == EV_PAIR_B_NEXT) ++arrivals;
    else { ++drains; if (!op.signal) return GN_E_SHAPE; arrivals += op.arrive; }
  }
  if (arrivals != waits || drains != signals) return GN_E_SHAPE;
  if (a.pro_op >= 0 && (a.pro_op >= a.nops || a.skip_last_op != a.pro_op || a.stage_mode != ST_PAIR || a.stage_first >= 0 ||
                        a.ops[a.pro_op].a_src != A_TMEM || a.ops[a.pro_op].wait_n != 1 || a.rs != nullptr)) return GN_E_SHAPE;
  if (a.stage_first >= 0) {
    // per-tile side work hangs off an EV_STAGE at e == 0; the staged buffer feeds the tile's FIRST op
    if (a.stage_first != 0 || a.ops[0].a_src != A_SMEM || a.ops[0].wait_n != 1 || a.stage_mode != ST_ROWS || a.tps != 0 ||
        a.rs != nullptr) return GN_E_SHAPE;
    // a_ready arrivals are consumed in op order: the next tile's staging arrival must be the LAST arrival of this tile,
    // and it may only overwrite the buffer after the last op that reads it has completed (its drain event precedes)
    int n_next = 0, last_reader = 0;
    for (int o = 0; o < a.nops; ++o) if (a.ops[o].a_src == A_SMEM) { if (a.ops[o].a_buf != a.ops[0].a_buf) return GN_E_SHAPE; last_reader = o; }
    bool reader_drained = false;
    for (int e = 0; e < a.nev; ++e) {
      if (a.ev_type[e] == EV_STAGE) return GN_E_SHAPE;
      if (a.ev_type[e] == EV_STAGE_NEXT) { if (!reader_drained || a.ev_op[e] != 0) return GN_E_SHAPE; ++n_next; continue; }
      if (a.ev_op[e] >= last_reader) reader_drained = true;      // acc_ready signals complete in op order
      if (n_next > 0 && a.ops[a.ev_op[e]].arrive) return GN_E_SHAPE;
    }
    if (n_next != 1) return GN_E_SHAPE;
  }
  // run-ahead check over two consecutive tiles: greedy row threads against a lazy issuer and vice versa
  for (int greedy_rows = 0; greedy_rows < 2; ++greedy_rows) {
    int io = 0, ie = 0;                       // issuer op index / row event index (over 2 tiles)
    int arr = a.stage_first >= 0 ? 1 : 0;     // produced arrivals (the first tile of such a program is staged up front)
    int cons = 0, dr = 0;                     // consumed arrivals, consumed signals
    int sig = a.pro_op >= 0 ? a.ops[a.pro_op].signal : 0;   // produced signals (a pro_op was staged and issued up front)
    int pend_wait = a.ops[0].wait_n;
    const int NO = 2 * a.nops, NE = 2 * a.nev;
    auto step_rows = [&]() -> bool {
      if (ie >= NE) return false;
      const int e = ie % a.nev;
      const Op& op = a.ops[a.ev_op[e]];
      if (a.ev_type[e] == EV_PAIR_A_NEXT) { ++ie; return true; }
      if (a.ev_type[e] == EV_STAGE || a.ev_type[e] == EV_STAGE_NEXT || a.ev_type[e] == EV_PAIR_B_NEXT) { ++arr; ++ie; return true; }
      if (dr < sig) { ++dr; arr += op.arrive; ++ie; return true; }
      return false;
    };
    auto step_issuer = [&]() -> bool {
      if (io >= NO) return false;
      if (pend_wait > 0) { if (cons < arr) { ++cons; --pend_wait; return true; } return false; }
      sig += a.ops[io % a.nops].signal;
      ++io;
      if (io < NO) pend_wait = a.ops[io % a.nops].wait_n;
      return true;
    };
    for (;;) {
      bool moved = false;
      if (greedy_rows) { while (step_rows()) moved = true; if (step_issuer()) moved = true; }
      else { while (step_issuer()) moved = true; if (step_rows()) moved = true; }
      if (arr - cons >= NBAR || sig - dr >= NBAR) return GN_E_SHAPE;
      if (!moved) break;
    }
    if (io != NO || ie != NE) return GN_E_SHAPE;           // deadlock
  }
  return GN_OK;
}

}  // namespace tfe
extern unsigned long long* g_trace_buffer;      // gn_profile_set_trace (gn_edge_mlp_tc.cu)
namespace tfe {

constexpr uint32_t VM_CHAIN = (1u << DV_TMEM) | (1u << DV_TMEM_RELU) | (1u << DV_DOT_RELU) | (1u << DV_COUNT);
constexpr uint32_t VM_NODE = (1u << DV_TMEM) | (1u << DV_TMEM_RELU) | (1u << DV_TMEM_STORE) | (1u << DV_STORE);
constexpr uint32_t VM_HAGG = (1u << DV_TMEM_RELU) | (1u << DV_TMEM_RELU_RS) | (1u << DV_STORE) | (1u << DV_STORE_BM);
constexpr uint32_t VM_ALL = (2u << DV_COUNT) - 1u;

static int launch(Builder& b, long long R, long long ntiles, const unsigned char* wstream, const char* name, cudaStream_t st) {
  Args& a = b.a;
  a.trace = g_trace_buffer;
  if (a.trace != nullptr) {                      // tracing (profiles/trace_tf32.py): only the named chain writes the buffer
    const char* only = getenv("GN_TRACE_KERNEL");
    if (only != nullptr && strcmp(only, name) != 0) a.trace = nullptr;
  }
  a.R = R; a.ntiles = ntiles; a.wstream = wstream;
  a.a0_half_bytes = static_cast<uint32_t>(b.a0_K) * 128 * 4;
  a.a0_buf_bytes = 2 * a.a0_half_bytes;
  a.off_a0 = 0;
  a.off_ring = static_cast<uint32_t>(b.nbuf) * a.a0_buf_bytes;
  const uint32_t node_bytes = b.node_block ? node_block_bytes() : 0u;
  const uint32_t fixed = a.off_ring + node_bytes + FIXED_BYTES;
  a.stage_bytes = b.stage_bytes;
  if (fixed + 2 * a.stage_bytes > SMEM_BUDGET) return GN_E_SHAPE;
  int ns = static_cast<int>((SMEM_BUDGET - fixed) / a.stage_bytes);
  if (ns > 8) ns = 8;
  a.nstage = ns;
  a.off_node = a.off_ring + static_cast<uint32_t>(ns) * a.stage_bytes;
  a.off_scr = (a.off_node + node_bytes + 127u) & ~127u;
  a.off_aux = (a.off_scr + SCR_BYTES + 127u) & ~127u;
  a.off_bar = (a.off_aux + AUX_FLOATS * 4 + 127u) & ~127u;
  const uint32_t smem = a.off_bar + static_cast<uint32_t>(sizeof(Bars));
  int rc = validate_program(a);
  if (rc != GN_OK) return rc;
  if (!wstream || (reinterpret_cast<uintptr_t>(wstream) & 15)) return GN_E_NULL;
  uint32_t need = 0;
  for (int o = 0; o < a.nops; ++o)
    if (a.ops[o].signal && a.ops[o].drain != DR_NONE) need |= a.ops[o].drain == DR_DOTG ? (1u << DV_COUNT) : (1u << a.ops[o].variant);
  // the kernel instances compile the noise / per-row-scale staging in only next to the drains that consume them
  if (a.edge_feat != nullptr && !(need & (1u << DV_COUNT))) return GN_E_SHAPE;
  if (a.rs != nullptr && !(need & ((1u << DV_TMEM_RELU_RS) | (1u << DV_STORE_BM)))) return GN_E_SHAPE;
  void (*kern)(Args) = nullptr;
  const bool trc = a.trace != nullptr;
#define GN_KERN(P, M) (trc ? chain_tf32_kernel<P, M, true> : chain_tf32_kernel<P, M, false>)
  if (a.stage_mode == ST_PAIR) {
    if (need & ~VM_CHAIN) return GN_E_SHAPE;
    kern = GN_KERN(true, VM_CHAIN);
  } else if (!(need & ~VM_NODE)) kern = GN_KERN(false, VM_NODE);
  else if (!(need & ~VM_HAGG)) kern = GN_KERN(false, VM_HAGG);
  else if (!(need & ~VM_CHAIN)) kern = GN_KERN(false, VM_CHAIN);
  else kern = GN_KERN(false, VM_ALL);
#undef GN_KERN
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = ntiles < GN_SM_COUNT ? static_cast<int>(ntiles) : GN_SM_COUNT;
  {
    ProfScope ps__(name, st);
    cudaError_t le = launch_pdl(kern, dim3(grid), dim3(THREADS), smem, st, a);
    if (le != cudaSuccess) return static_cast<int>(le);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace tfe

using namespace tfe;

// ---- per-edge chain: init_MLP -> [MLP_distribution | MLP_factor] -> Gumbel softmax / sigmoid (:41-53)
//   G1 64->128 (ReLU)  G2 128->64 (z)  G3f 64->128 (factor hidden, ReLU; its 128->1 head is a dot in the drain)
//   G3d 64->128 (distribution hidden, ReLU; its 128->T head is T dots in the drain)  -> Gumbel softmax / sigmoid
// TMEM columns: acc1 0 | A1 128,256 | acc2 384 | A2 0,64 | acc3f 128 | acc3d 256 | logit exchange 448..511
bool edge_chain_tf32_fits(bool pair, int N, int T) {
  if (T < 1 || T > GN_SMALL_OUT - 1) return false;
  if (!pair) return true;
  const int E = N * N;
  if (E >= 128) return N <= MAXN;
  return (127 / E + 2) * N <= MAXN;
}

int launch_edge_chain_tf32(bool pair, const float* edges, const float* ypre, const float* pq,
                           int N, int E, int T, long long R, const gn_stage_weights* w,
                           const float* U, int noise_mode, unsigned long long seed, long long scene_offset,
                           int stage_index, float* dist_out, float* edge_feat, cudaStream_t st) {
  if (!w->tf_chain_w) return GN_E_NULL;
  if (R <= 0) return GN_OK;
  if (pair && (!ypre || !pq || R >= (1LL << 31))) return GN_E_SHAPE;
  // pair: the first Linear of init_MLP commutes with the attention-weighted gather (edges_e = w_i x'_i + w_j x'_j, :138),
  // so it ran once per NODE in the prologue (Y = x' W^T) and the staging pass emits relu(w_i Y_i + w_j Y_j + b) straight
  // into tensor memory: G1, its drain and its 64 KB of weights per tile are gone from the N^2 edge rows
  Builder b(pair ? 0 : 64, 1, pair, pair);
  if (b.stage_bytes != 65536u) return GN_E_SHAPE;          // the stream is packed for 64 KB stages in both forms
  Args& a = b.a;
  Op* g3d_p = nullptr;
  if (!pair) {
    Op& g1 = b.add(A_SMEM, 0, 64, 128, 0, 0, 1, 1);
    b.drain_tmem(g1, 128, 1, w->init_b0, 128, 1);
    Op& g2 = b.add(A_TMEM, 128, 128, 64, 384, 0, 1, 1);
    b.drain_tmem(g2, 64, 0, w->init_b1, 0, 1);
    Op& g3f = b.add(A_TMEM, 0, 64, 128, 128, 0, 1, 1);
    g3f.drain = DR_DOT; g3f.dn = 128; g3f.relu = 1; b.set_bias(g3f, w->df_b0 + 128, 128); g3f.arrive = 0;
    Op& g3d = b.add(A_TMEM, 0, 64, 128, 256, 0, 0, 1);
    // the distribution head (128 -> T) and the Gumbel softmax run inside the drain of G3d (DR_DOTG); the slices' partial
    // logits meet in columns 448..511, the only ones no accumulator or operand of this chain uses (G1 of the NEXT tile,
    // whose rows are staged early, may already be accumulating into 0..127 by then)
    g3d.drain = DR_DOTG; g3d.dn = 128; g3d.relu = 1; b.set_bias(g3d, w->df_b0, 128); g3d.dst_col = 448; g3d.arrive = 0;
    g3d_p = &g3d;
    // the next tile's rows are staged (HBM -> split -> smem) and its Gumbel noise drawn while the two head GEMMs run:
    // G1 has completed (D1 waited for it), and the staging arrival is the tile's last (D2 releases both heads)
    a.stage_first = 0;
    b.ev(EV_DRAIN, 0);
    b.ev(EV_DRAIN, 1);
    b.ev(EV_STAGE_NEXT, 0);
    b.ev(EV_DRAIN, 2);
    b.ev(EV_DRAIN, 3);
  } else {
    // Software-pipelined pairwise form.  The row threads are the bottleneck of this chain (SIMT ~14 K clk per tile
    // against ~7 K of tensor time), so the fused node2edge of tile i + 1 runs inside the MMA waits of tile i:
    //   issuer per tile:  G3f(i)              G3d(i)              G2(i+1)
    //   row threads:      D2(i)  stage A(i+1) DOTf(i) stage B(i+1) DOTG(i)      (stage A: attention, B: hidden -> TMEM)
    // G2 of a CTA's first tile is issued up front (pro_op) and skipped after its last tile.  The hidden operand of G2
    // owns columns 0..255 (free again once G2 has completed, which D2 waits for); everything else lives in 256..511:
    // acc2 384 -> A2 hi in place 384 | lo 448 -> ONE accumulator 256..383 for G3f then G3d -> logit exchange 448.
    a.yb_off = b.aux(w->init_b0, 128);
    a.att_off = b.aux(w->att_b0, 32);
    b.aux(w->att_w1, 32);
    b.aux(w->att_b1, 1);
    const size_t g2_bytes = 64 * 128 * 8;
    Op& g3f = b.add(A_TMEM, 384, 64, 128, 256, 0, 1, 1);
    g3f.drain = DR_DOT; g3f.dn = 128; g3f.relu = 1; b.set_bias(g3f, w->df_b0 + 128, 128);
    g3f.arrive = 1;                                          // the accumulator is free for G3d
    Op& g3d = b.add(A_TMEM, 384, 64, 128, 256, 0, 1, 1);
    g3d.drain = DR_DOTG; g3d.dn = 128; g3d.relu = 1; b.set_bias(g3d, w->df_b0, 128); g3d.dst_col = 448; g3d.arrive = 0;
    g3d_p = &g3d;
    Op& g2 = b.add(A_TMEM, 0, 128, 64, 384, 0, 1, 1);
    b.drain_tmem(g2, 64, 0, w->init_b1, 384, 1);
    // stream order is G2 | G3f | G3d (packing.py): the ops were added in issue order, fix their offsets
    g2.w_off = 0; g3f.w_off = static_cast<int>(g2_bytes); g3d.w_off = static_cast<int>(g2_bytes + 128 * 64 * 8);
    a.pro_op = 2; a.skip_last_op = 2;
    b.ev(EV_DRAIN, 2);            // D2(i): acc2 -> A2, releases G3f(i)
    b.ev(EV_PAIR_A_NEXT, 2);
    b.ev(EV_DRAIN, 0);            // DOTf(i), releases G3d(i)
    b.ev(EV_PAIR_B_NEXT, 2);      // hidden of tile i + 1 -> columns 0..255, releases G2(i+1)
    b.ev(EV_DRAIN, 1);            // DOTG(i)
  }
  Op& g3d = *g3d_p;
  a.stage_mode = pair ? ST_PAIR : ST_ROWS;
  a.src0 = edges; a.ld0 = 64; a.k_src0 = 64; a.src1 = nullptr; a.ld1 = 0; a.a_div = 0.f;
  a.ypre = ypre; a.pq = pq;
  a.N = N; a.E = E;
  a.tps = (pair && E >= 128) ? (E + 127) / 128 : 0;
  // the stream: init_MLP.0 (64 KB, skipped by the pair form) | the chunks of the ops above | MLP_factor.layers.1.weight
  // (128 floats) | MLP_distribution.layers.1.weight as [128][6, 8, 10, 12 or 16] fp32 (k-major rows of T logits, zero padded)
  const unsigned char* stream = static_cast<const unsigned char*>(w->tf_chain_w);
  const size_t w1_bytes = 128 * 64 * 8;
  const float* tail = reinterpret_cast<const float*>(stream + (pair ? w1_bytes : 0) + b.wbytes);
  a.dot_off = b.aux(tail, 128);
  a.w4_tp = T <= 6 ? 6 : (T <= 8 ? 8 : (T <= 10 ? 10 : (T <= 12 ? 12 : 16)));
  a.w4_off = b.aux(tail + 128, 128 * a.w4_tp);
  a.gb_off = b.aux(w->df_b1, GN_SMALL_OUT);
  if (a.dot_off < 0 || a.w4_off < 0 || a.gb_off < 0 || g3d.bias_off < 0 || (pair && (a.yb_off < 0 || a.att_off < 0))) return GN_E_SHAPE;
  a.T = T;
  a.U = U; a.noise_mode = noise_mode; a.seed = seed; a.scene_offset = scene_offset; a.stage_index = stage_index;
  a.dist_out = dist_out; a.edge_feat = edge_feat;
  const long long ntiles = a.tps ? (R / E) * a.tps : (R + 127) / 128;
  return launch(b, R, ntiles, stream + (pair ? w1_bytes : 0), pair ? "edge_chain_pair_tf32" : "edge_chain_tf32", st);
}

// ---- node prologue: x' = node2edge_start_mlp(h) (D -> 256 -> 64, :84,:125), pq = split attention layer 0 (:80,:134)
// the 256-wide hidden layer runs as two 128-column halves feeding the second Linear as two K = 128 chunks
bool node_pre_tf32_fits(int D) { return D >= 8 && D <= 128 && (D & 7) == 0; }

int launch_node_pre_tf32(const float* h, long long R, int D, const gn_stage_weights* w, float* xprime, float* pq,
                         float* ypre, cudaStream_t st) {
  if (!w->tf_pre_w) return GN_E_NULL;
  if (R <= 0) return GN_OK;
  Builder b(D, 1, false, false);
  Args& a = b.a;
  // Both halves of the 256-wide hidden layer are issued back to back (accumulators 0 and 384), so the second one runs
  // while the row threads drain the first; the second Linear's first K chunk then accumulates over the drained
  // accumulator (0..63) and SIGNALS for the second half's drain (drain_col 384): by then it has finished reading the
  // hidden operand that drain overwrites.  x' is split in place (0..127); pq and Y accumulate over the dead hidden operand.
  // TMEM columns: acc_a 0 | acc_b 384 | A_hid 128,256 | acc_x 0 -> A_x 0,64 | acc_Y 128 | acc_pq 256
  const int w0_bytes = 128 * D * 8, w1_bytes = 64 * 128 * 8;
  Op& s0a = b.add(A_SMEM, 0, D, 128, 0, 0, 1, 1);
  b.drain_tmem(s0a, 128, 1, w->node_b0, 128, 1);
  Op& s0b = b.add(A_SMEM, 0, D, 128, 384, 0, 0, 0);
  Op& s1a = b.add(A_TMEM, 128, 128, 64, 0, 0, 1, 1);
  b.drain_tmem(s1a, 128, 1, w->node_b0 + 128, 128, 1);
  s1a.drain_col = 384;
  Op& s1b = b.add(A_TMEM, 128, 128, 64, 0, 1, 1, 1);
  s1b.drain = DR_TMEM_STORE; s1b.dn = 64; s1b.relu = 0; b.set_bias(s1b, w->node_b1, 64); s1b.dst_col = 0; s1b.arrive = 1;
  s1b.out = xprime; s1b.ldo = 64; s1b.out_col0 = 0;
  Op& s2 = b.add(A_TMEM, 0, 64, 64, 256, 0, 1, 1);
  b.drain_store(s2, 64, 0, nullptr, pq, 64, 0, 0);
  // the stream is packed W0[:128] | W1[:, :128] | W0[128:] | W1[:, 128:] | pq | Y (packing.py): fix the offsets
  s0a.w_off = 0; s1a.w_off = w0_bytes; s0b.w_off = w0_bytes + w1_bytes; s1b.w_off = 2 * w0_bytes + w1_bytes;
  s2.w_off = 2 * w0_bytes + 2 * w1_bytes;
  // the staged h tile is read by ops 0 and 1: both have completed before the first drain event returns.  The next tile's
  // rows are staged after the drain of op 3: the issuer consumes a_ready arrivals in op order, and the staging arrival
  // belongs to the NEXT tile's first op, so it must follow the last arriving drain of this tile
  a.stage_first = 0;
  b.ev(EV_DRAIN, 0); b.ev(EV_DRAIN, 2); b.ev(EV_DRAIN, 3); b.ev(EV_STAGE_NEXT, 0); b.ev(EV_DRAIN, 4);
  if (ypre != nullptr) {
    // pairwise layers: Y = x' W_init0^T (no bias), the per-node half of init_MLP's first Linear (see the edge chain)
    Op& s3 = b.add(A_TMEM, 0, 64, 128, 128, 0, 0, 1);
    b.drain_store(s3, 128, 0, nullptr, ypre, 128, 0, 0);
    s3.w_off = 2 * w0_bytes + 2 * w1_bytes + 64 * 64 * 8;
    b.ev(EV_DRAIN, 5);
  }
  a.stage_mode = ST_ROWS; a.src0 = h; a.ld0 = D; a.k_src0 = D;
  return launch(b, R, (R + 127) / 128, static_cast<const unsigned char*>(w->tf_pre_w), "node_pre_tf32", st);
}

// ---- pairwise aggregation, first half of the collapse: P[:, t*128:(t+1)*128] = h W0_t^T (bias added per edge later)
bool agg_in_tf32_fits(int D, int T) { return D >= 8 && D <= 128 && (D & 7) == 0 && T >= 1 && T <= MAX_OPS; }

int launch_agg_in_tf32(const float* h, long long R, int D, int T, const gn_stage_weights* w, float* P, cudaStream_t st) {
  if (!w->tf_aggin_w) return GN_E_NULL;
  if (R <= 0) return GN_OK;
  Builder b(D, 1, false, false);
  Args& a = b.a;
  b.ev(EV_STAGE, 0);
  for (int t = 0; t < T; ++t) {
    Op& o = b.add(A_SMEM, 0, D, 128, (t & 1) * 128, 0, (t == 0 || t >= 2) ? 1 : 0, 1);
    b.drain_store(o, 128, 0, nullptr, P, static_cast<long long>(T) * 128, 0, (t + 2 < T) ? 1 : 0);
    o.out = P + t * 128;                 // column block t (out_col0 is a short: keep it 0)
    b.ev(EV_DRAIN, t);
  }
  a.stage_mode = ST_ROWS; a.src0 = h; a.ld0 = D; a.k_src0 = D;
  return launch(b, R, (R + 127) / 128, static_cast<const unsigned char*>(w->tf_aggin_w), "agg_in_tf32", st);
}

// ---- pairwise aggregation, second half: agg = G W1cat^T + S b1, K = T*128 streamed through two staged buffers
bool agg_out_tf32_fits(int D, int T) {
  return (D == 64 || D == 128) && T >= 1 && 2 * T <= MAX_OPS && 4 * T + 2 <= MAX_EV;
}

int launch_agg_out_tf32(const float* G, const float* S, long long R, int D, int T, const gn_stage_weights* w,
                        float* agg, cudaStream_t st) {
  if (!w->tf_aggout_w) return GN_E_NULL;
  if (R <= 0) return GN_OK;
  Builder b(64, 2, false, false);
  Args& a = b.a;
  const int n = 2 * T;                    // K chunks of 64
  // chunk c accumulates into accumulator c % NACC (columns 128 apart): four short chains instead of one long one,
  // summed in fp32 by the final drain (the tensor core's accumulation rounds toward zero)
  const int NACC = n < 4 ? n : 4;
  for (int c = 0; c < n; ++c) {
    Op& o = b.add(A_SMEM, c & 1, 64, D, (c % NACC) * 128, c >= NACC ? 1 : 0, 1, 1);
    o.st_k0 = static_cast<short>(64 * c);
    if (c == n - 1) {
      // the last chunk goes to accumulator 0, where the drain starts summing (any accumulator may take it)
      b.drain_store(o, D, 0, nullptr, agg, D, 0, 0);
      o.acc_col = 0; o.accumulate = 1; o.nsum = static_cast<short>(NACC); o.sum_stride = 128;
      o.use_bm = 1;
    } else {
      o.drain = DR_NONE;
    }
  }
  // rows stage two chunks ahead; a buffer is refilled once the MMAs that read it have completed
  b.ev(EV_STAGE, 0);
  if (n > 1) b.ev(EV_STAGE, 1);
  for (int c = 0; c < n; ++c) {
    b.ev(EV_DRAIN, c);
    if (c + 2 < n) b.ev(EV_STAGE, c + 2);
  }
  a.stage_mode = ST_ROWS; a.src0 = G; a.ld0 = static_cast<long long>(T) * 128; a.k_src0 = T * 128;
  a.rs = S; a.rs_ld = 16; a.rs_n = T; a.bm = w->agg_b1; a.bm_T = T; a.bm_ld = D; a.bm_off = b.aux(w->agg_b1, T * D);
  return launch(b, R, (R + 127) / 128, static_cast<const unsigned char*>(w->tf_aggout_w), "agg_out_tf32", st);
}

// (A pipelined form with 64-unit halves — two hidden accumulators, the first Linear two unit steps ahead — measured
//  SLOWER, 2.5 vs 1.9 ms: a 128 x 64 x 8 tf32 MMA holds the tensor pipe ~50 clk against ~66 for 128 x 128 x 8, so
//  halving N costs 1.5x the tensor time)
// ---- hyper edge_aggregation as written (:259-265): ef = sum_t edge_feat_t * (W1_t relu(W0_t eo + b0_t) + b1_t)
// TMEM columns: acc_hid 0 | A_hid 128,256 | acc_ef 384
bool hyper_agg_tf32_fits(int D, int T) {
  // biases (T x 128) and the rank-T output bias (T x D) must fit the smem constants
  return (D == 64 || D == 128) && T >= 1 && 2 * T <= MAX_OPS && T * 128 + T * D <= AUX_FLOATS;
}

int launch_hyper_agg_tf32(const float* eo, const float* edge_feat, long long R, int D, int T,
                          const gn_stage_weights* w, float* ef, cudaStream_t st) {
  if (!w->tf_hagg_w) return GN_E_NULL;
  if (R <= 0) return GN_OK;
  Builder b(D, 1, false, false);
  Args& a = b.a;
  b.ev(EV_STAGE, 0);
  for (int t = 0; t < T; ++t) {
    Op& g1 = b.add(A_SMEM, 0, D, 128, 0, 0, t == 0 ? 1 : 0, 1);
    b.drain_tmem(g1, 128, 1, w->agg_b0 + t * 128, 128, 1);
    g1.rs_idx = static_cast<short>(t);
    b.ev(EV_DRAIN, 2 * t);
    // two short accumulation chains (even / odd t) when they fit beside A_hid, summed in fp32 by the final drain
    const int NACC = (D <= 64 && T >= 2) ? 2 : 1;
    const bool last = t == T - 1;
    Op& g2 = b.add(A_TMEM, 128, 128, D, last ? 384 : 384 + (t % NACC) * 64, (last ? t > 0 : t >= NACC) ? 1 : 0, 1, last ? 1 : 0);
    if (last) {
      b.drain_store(g2, D, 0, nullptr, ef, D, 0, 0);
      g2.nsum = static_cast<short>(NACC); g2.sum_stride = 64;
      g2.use_bm = 1;
      b.ev(EV_DRAIN, 2 * t + 1);
    }
  }
  a.stage_mode = ST_ROWS; a.src0 = eo; a.ld0 = D; a.k_src0 = D;
  a.rs = edge_feat; a.rs_ld = T; a.rs_n = T; a.bm = w->agg_b1; a.bm_T = T; a.bm_ld = D; a.bm_off = b.aux(w->agg_b1, T * D);
  return launch(b, R, (R + 127) / 128, static_cast<const unsigned char*>(w->tf_hagg_w), "hyper_agg_tf32", st);
}

// ---- closing MLP on [agg | h] / N (:120,:355,:195,:441): 2D -> 128 (ReLU) -> Dout
bool node_post_tf32_fits(int D, int Dout, long long ld_out, const float* node_out) {
  return D >= 4 && D <= 64 && (D & 3) == 0 && (Dout == 64 || Dout == 128) && (ld_out & 3) == 0 &&
         (reinterpret_cast<uintptr_t>(node_out) & 15) == 0;
}

int launch_node_post_tf32(const float* agg, const float* h, long long R, int D, int Nagents, int Dout,
                          const gn_stage_weights* w, float* node_out, long long ld_out, cudaStream_t st) {
  if (!w->tf_post_w) return GN_E_NULL;
  if (R <= 0) return GN_OK;
  Builder b(2 * D, 1, false, false);
  Args& a = b.a;
  Op& p0 = b.add(A_SMEM, 0, 2 * D, 128, 0, 0, 1, 1);
  b.drain_tmem(p0, 128, 1, w->post_b0, 128, 1);
  Op& p1 = b.add(A_TMEM, 128, 128, Dout, 384, 0, 1, 1);
  b.drain_store(p1, Dout, 0, w->post_b1, node_out, ld_out, 0, 0);
  a.stage_first = 0;                      // [agg | h] of the next tile is staged while the second Linear runs
  b.ev(EV_DRAIN, 0); b.ev(EV_STAGE_NEXT, 0); b.ev(EV_DRAIN, 1);
  a.stage_mode = ST_ROWS; a.src0 = agg; a.ld0 = D; a.k_src0 = D; a.src1 = h; a.ld1 = D;
  a.a_div = static_cast<float>(Nagents);
  return launch(b, R, (R + 127) / 128, static_cast<const unsigned char*>(w->tf_post_w), "node_post_tf32", st);
}

}  // namespace gn
