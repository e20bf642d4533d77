// Fused hyper layer tail for h_dim = 64 (the NBA / fish shapes), E == N <= 64, Dout in {32, 64}:
//
//   eo  = H @ h                                                            (model/MS_HGNN_batch.py:263)
//   ef  = sum_t edge_feat[:, t] * (W1_t relu(W0_t eo + b0_t) + b1_t)       (:264-265, T MLPs 64->128->64)
//   agg = H^T @ ef ;  out = W1 relu(W0 [agg | h]/N + b0) + b1              (:267, :120, :195/:441)
//
// Same structure as gn_hyper_fused_tc.cu (h_dim 256) at a quarter of the footprint: 104 KB of shared
// memory and 256 TMEM columns per CTA, so TWO CTAs share an SM and one CTA's serial staging / epilogue
// phases overlap the other's MMA main loop.  Per 128-row tile (128 / N whole scenes):
//   gather   eo[128 x 64]   = Hblk[128 x 128] * hT[64 x 128]^T
//   per t    hid[128 x 128] = [eo | 1][128 x 80] * [W0_t | b0_t]^T ;  A2 = bf16(relu(hid) * edge_feat_t)
//            ef[128 x 64]  += [A2 | edge_feat_t][128 x 144] * [W1_t | b1_t]^T
//   scatter  agg[128 x 64]  = HblkT[128 x 128] * efT[64 x 128]^T
//   post     o1[128 x 128]  = [agg/N | h/N | 1][128 x 144] * [P0 | b0]^T ;  out = [relu(o1) | 1] * [P1 | b1]^T
// Warps 0-7 stage / drain, warp 8 streams the weight chunks (cp.async.bulk, one linear host-packed
// stream) through a 2-stage ring that lives in the scratch region the incidence operands use outside
// the main loop, warp 9 issues every tcgen05.mma.  It replaces hyper_agg_tc + edge2node_hyper + the
// closing node chain and the eo half of node2edge_hyper for these layers: eo, ef and agg never touch HBM.
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {
namespace hf64 {
constexpr int D = 64;
constexpr int THREADS = 320;
constexpr uint32_t OFF_EO = 0;                         // eo A operand [128 x 64]; post: closing W1 [Dout x 128]
constexpr uint32_t OFF_ONES = 16384;                   // k-groups 8,9 of the A operand: 1.0, 1.0, 0...
constexpr uint32_t OFF_A2 = 20480;                     // A2 [128 x 144]; raw incidence; inc; o1
constexpr uint32_t A2_BYTES = 128 * 144 * 2;
constexpr uint32_t OFF_SCR = OFF_A2 + A2_BYTES;        // scratch 48 KB: Hblk|hT, ring, HblkT|efT, P0|b1, out transposes
constexpr uint32_t SCR_BYTES = 49152;
constexpr uint32_t OFF_HT = OFF_SCR + 32768;
constexpr uint32_t STAGE = 20480;
constexpr uint32_t W0C = 128 * 80 * 2;                 // [W0_t | b0_t (hi, lo)]
constexpr uint32_t W1C = 64 * 144 * 2;                 // [W1_t | b1_t (hi, lo, hi)]
constexpr uint32_t P0C = 128 * 144 * 2;                // closing [W0 | b0 (hi, lo)]
constexpr uint32_t OFF_PB = OFF_SCR + P0C;             // closing b1 block [Dout x 16]
constexpr uint32_t OFF_BAR = OFF_SCR + SCR_BYTES;
enum { B_WFULL = 0, B_WEMPTY = 2, B_HFULL = 4, B_HFREE = 5, B_A2FULL = 6, B_A2FREE = 7, B_STAGE = 8, B_EOFULL = 9,
       B_EOREADY = 10, B_EFFULL = 11, B_EFTREADY = 12, B_AGGFULL = 13, B_PAREADY = 14, B_PFULL0 = 15,
       B_PFULL1 = 16, B_O1FULL = 17, B_O1READY = 18, B_OUTFULL = 19, NBAR = 20 };
constexpr uint32_t SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
static_assert(2 * (SMEM_BYTES + 1024) <= 227 * 1024, "hyper_fused64: two CTAs per SM");
static_assert(OFF_PB + 64 * 16 * 2 <= OFF_BAR, "hyper_fused64: scratch layout");
constexpr uint32_t TM_S = 0, TM_HID = 64;
}  // namespace hf64

extern unsigned long long* g_trace_buffer;

struct HyperFused64Args {
  const float* h; const float* H; const float* edge_feat; const unsigned char* wstream;
  float* node_out; long long ld_out; int Dout; int B, N, T; long long hstride;
  unsigned long long* trace;          // optional phase timeline (build with -DGN_ENABLE_TRACE)
};

namespace {
__device__ __forceinline__ void arrive(uint64_t* b) { tc::mbar_arrive(b); }
// producer-warp forms: every lane calls them, one elected lane issues (gn_tc.cuh, tcu)
using tcu::expect_tx; using tcu::bulk_g2s; using tc::drain_bar; using tc::stage_raw_H;
}  // namespace

__global__ void __launch_bounds__(hf64::THREADS, 2)
hyper_fused64_tc_kernel(HyperFused64Args a) {
  using namespace hf64;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < NBAR; ++i) {
      uint32_t cnt = 1;
      if (i == B_HFREE || i == B_A2FULL || i == B_STAGE || i == B_EOREADY || i == B_EFTREADY || i == B_PAREADY ||
          i == B_O1READY) cnt = 256;
      tc::mbar_init(bars + i, cnt);
    }
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 256);
  if (tid < 128) tc::build_ones_operand(smem + OFF_ONES, tid, 128);
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sbase = tc::smem_u32(smem);
  const int N = a.N, SC = 128 / N, T = a.T;
  const int ntiles = (a.B + SC - 1) / SC;
  const size_t post_off = static_cast<size_t>(T) * (W0C + W1C);
  const uint32_t p1_bytes = static_cast<uint32_t>(a.Dout) * 256u, pb_bytes = static_cast<uint32_t>(a.Dout) * 32u;

  if (warp == 8) {
    // ------------------------------------------------------------------ weight stream producer
    {                                                   // all 32 lanes: warp-uniform control flow, elected issue
      uint32_t ph_empty = 0x3u, ph_eofull = 0u, ph_agg = 0u;
      int stage = 0;
      auto load = [&](const unsigned char*& src, uint32_t bytes) {
        tc::mbar_wait(bars + B_WEMPTY + stage, (ph_empty >> stage) & 1u);
        ph_empty ^= 1u << stage;
        expect_tx(bars + B_WFULL + stage, bytes);
        bulk_g2s(sbase + OFF_SCR + stage * STAGE, src, bytes, bars + B_WFULL + stage);
        src += bytes;
        stage ^= 1;
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        tc::mbar_wait(bars + B_EOFULL, ph_eofull);      // Hblk / hT (scratch) consumed by the gather MMA
        ph_eofull ^= 1u;
        const unsigned char* src = a.wstream;
        for (int s = 0; s <= T; ++s) {
          if (s < T) load(src, W0C);
          if (s >= 1) load(src, W1C);
        }
        tc::mbar_wait(bars + B_AGGFULL, ph_agg);        // HblkT / efT consumed (and, before, every ring stage)
        ph_agg ^= 1u;
        const unsigned char* ps = a.wstream + post_off;
        expect_tx(bars + B_PFULL0, P0C);
        bulk_g2s(sbase + OFF_SCR, ps, P0C, bars + B_PFULL0);
        expect_tx(bars + B_PFULL1, p1_bytes + pb_bytes);
        bulk_g2s(sbase + OFF_EO, ps + P0C, p1_bytes, bars + B_PFULL1);
        bulk_g2s(sbase + OFF_PB, ps + P0C + p1_bytes, pb_bytes, bars + B_PFULL1);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    {                                                   // all 32 lanes: warp-uniform control flow, elected issue
      uint32_t ph = 1u << B_HFREE;                       // "free" barrier: the first wait passes
      int stage = 0;
      auto wait = [&](int i) { tc::mbar_wait(bars + i, (ph >> i) & 1u); ph ^= 1u << i; };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        wait(B_STAGE);
        tc::fence_after_thread_sync();
        tcu::issue_gemm(tmem + TM_S, sbase + OFF_SCR, sbase + OFF_HT, 64, 128, false);          // eo = Hblk * h
        tcu::mma_commit(bars + B_EOFULL);
        wait(B_EOREADY);
        tc::fence_after_thread_sync();
        for (int s = 0; s <= T; ++s) {
          if (s < T) {
            wait(B_HFREE);
            wait(B_WFULL + stage);
            tc::fence_after_thread_sync();
            tcu::issue_gemm(tmem + TM_HID, sbase + OFF_EO, sbase + OFF_SCR + stage * STAGE, 128, 80, false);
            tcu::mma_commit(bars + B_WEMPTY + stage);
            tcu::mma_commit(bars + B_HFULL);
            stage ^= 1;
          }
          if (s >= 1) {
            wait(B_A2FULL);
            wait(B_WFULL + stage);
            tc::fence_after_thread_sync();
            tcu::issue_gemm(tmem + TM_S, sbase + OFF_A2, sbase + OFF_SCR + stage * STAGE, 64, 144, s > 1);
            tcu::mma_commit(bars + B_WEMPTY + stage);
            tcu::mma_commit(bars + B_A2FREE);
            stage ^= 1;
          }
        }
        tcu::mma_commit(bars + B_EFFULL);
        wait(B_EFTREADY);
        tc::fence_after_thread_sync();
        tcu::issue_gemm(tmem + TM_S, sbase + OFF_SCR, sbase + OFF_HT, 64, 128, false);          // agg = HblkT * ef
        tcu::mma_commit(bars + B_AGGFULL);
        wait(B_PAREADY);
        wait(B_PFULL0);
        tc::fence_after_thread_sync();
        tcu::issue_gemm(tmem + TM_HID, sbase + OFF_A2, sbase + OFF_SCR, 128, 144, false);       // o1 (pre-ReLU)
        tcu::mma_commit(bars + B_O1FULL);
        wait(B_O1READY);
        wait(B_PFULL1);
        tc::fence_after_thread_sync();
        tcu::issue_gemm(tmem + TM_S, sbase + OFF_A2, sbase + OFF_EO, a.Dout, 128, false);
        tcu::issue_gemm(tmem + TM_S, sbase + OFF_A2 + 16 * 2048, sbase + OFF_PB, a.Dout, 16, true);
        tcu::mma_commit(bars + B_OUTFULL);
      }
    }
  } else {
    // ------------------------------------------------------------------ staging / drain warps
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;              // TMEM lane = tile row
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t ph = 1u << B_A2FREE;
    auto wait = [&](int i) { tc::mbar_wait(bars + i, (ph >> i) & 1u); ph ^= 1u << i; };
    const int r128 = tid & 127, half = tid >> 7;
    const int sc_r = r128 / N, in_r = r128 - sc_r * N;
    const int ldr = ((N + 3) & ~3) + 4;
    const float inv_n = 1.f / static_cast<float>(N);
    float* raw = reinterpret_cast<float*>(smem + OFF_A2);
    int titer = 0;
#ifdef GN_ENABLE_TRACE
#define HF64_TRACE(pt) do { if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && titer < 8) \
    a.trace[titer * 16 + (pt)] = clock64(); } while (0)
#else
#define HF64_TRACE(pt) do { } while (0)
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++titer) {
      const int b0s = tile * SC;
      HF64_TRACE(0);
      const int ns = min(SC, a.B - b0s);
      const int rows_used = ns * N;
      const bool rowvalid = row < rows_used;
      // ---- Hblk[edge row][node] (A) from a coalesced raw copy of the incidence blocks (A2 region is idle)
      stage_raw_H(raw, a.H, a.hstride, b0s, ns, N, ldr, tid);
      drain_bar();                                        // also: every warp is past its output transposes (scratch)
      {
        const bool valid = r128 < rows_used;
        const float* Hrow = raw + r128 * ldr;
#pragma unroll 4
        for (int kg = half * 8; kg < half * 8 + 8; ++kg) {
          float v[8];
          if (!valid || kg * 8 + 7 < sc_r * N || kg * 8 >= sc_r * N + N) {   // k-group outside this row's scene block
            *reinterpret_cast<uint4*>(smem + OFF_SCR + kg * 2048 + r128 * 16) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = kg * 8 + i - sc_r * N;
            v[i] = (n >= 0 && n < N) ? Hrow[n] : 0.f;
          }
          *reinterpret_cast<uint4*>(smem + OFF_SCR + kg * 2048 + r128 * 16) =
              make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                         tc::pack_bf16(v[6], v[7]));
        }
      }
      HF64_TRACE(1);
      // ---- hT[column c][node k] (B operand, 64 rows): thread = (column, 4 k-groups)
      {
        const int c = tid & 63, kq = tid >> 6;
        const float* hsrc = a.h + static_cast<size_t>(b0s) * N * D + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kg = kq * 4 + j;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int node = kg * 8 + i;
            v[i] = node < rows_used ? __ldg(hsrc + static_cast<size_t>(node) * D) : 0.f;
          }
          *reinterpret_cast<uint4*>(smem + OFF_HT + kg * 1024 + c * 16) =
              make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                         tc::pack_bf16(v[6], v[7]));
        }
      }
      tc::fence_proxy_async_smem();
      arrive(bars + B_STAGE);
      HF64_TRACE(2);
      const float* efrow = a.edge_feat + (static_cast<size_t>(b0s) * N + row) * T;
      float wnext = rowvalid ? __ldg(efrow) : 0.f;
      // ---- eo: TMEM -> bf16 A operand, this group's 32 columns
      wait(B_EOFULL);
      HF64_TRACE(3);
      tc::fence_after_thread_sync();
      {
        float v[32];
        tc::tmem_ld32(tmem + lane_addr + TM_S + g * 32, v);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(smem + OFF_EO + (g * 4 + q) * 2048 + row * 16) =
              make_uint4(tc::pack_bf16_fast(v[8 * q], v[8 * q + 1]), tc::pack_bf16_fast(v[8 * q + 2], v[8 * q + 3]),
                         tc::pack_bf16_fast(v[8 * q + 4], v[8 * q + 5]), tc::pack_bf16_fast(v[8 * q + 6], v[8 * q + 7]));
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      arrive(bars + B_EOREADY);
      HF64_TRACE(4);
      // ---- main loop: hidden columns [64g, 64g+64) of step t -> A2
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        const float w = wnext;
        if (t + 1 < T) wnext = rowvalid ? __ldg(efrow + t + 1) : 0.f;
        if (t == 1 && tile + static_cast<int>(gridDim.x) < ntiles) {      // next tile's h, H, edge_feat -> L2
          const int nb0 = (tile + gridDim.x) * SC;
          const int nns = min(SC, a.B - nb0);
          const char* hp = reinterpret_cast<const char*>(a.h + static_cast<size_t>(nb0) * N * D);
          const int hlines = (nns * N * D * 4 + 127) >> 7;
          for (int i = tid; i < hlines; i += 256) asm volatile("prefetch.global.L2 [%0];" :: "l"(hp + (static_cast<size_t>(i) << 7)));
          const int per_lines = (N * N * 4 + 127) >> 7;
          for (int i = tid; i < nns * per_lines; i += 256) {
            const int sc = i / per_lines, l = i - sc * per_lines;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(a.H + static_cast<size_t>(nb0 + sc) * a.hstride) + (static_cast<size_t>(l) << 7)));
          }
          const char* ep = reinterpret_cast<const char*>(a.edge_feat + static_cast<size_t>(nb0) * N * T);
          const int elines = (nns * N * T * 4 + 127) >> 7;
          for (int i = tid; i < elines; i += 256) asm volatile("prefetch.global.L2 [%0];" :: "l"(ep + (static_cast<size_t>(i) << 7)));
        }
        wait(B_HFULL);
        tc::fence_after_thread_sync();
        uint32_t r0[32], r1[32];
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_HID + g * 64, r0);
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_HID + g * 64 + 32, r1);
        tc::tmem_ld_wait();
        tc::fence_before_thread_sync();
        arrive(bars + B_HFREE);
        wait(B_A2FREE);
        unsigned char* a2 = smem + OFF_A2 + g * 8 * 2048;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = tc::pack_bf16_relu(__uint_as_float(r0[8 * q]) * w, __uint_as_float(r0[8 * q + 1]) * w);
          o.y = tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 2]) * w, __uint_as_float(r0[8 * q + 3]) * w);
          o.z = tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 4]) * w, __uint_as_float(r0[8 * q + 5]) * w);
          o.w = tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 6]) * w, __uint_as_float(r0[8 * q + 7]) * w);
          *reinterpret_cast<uint4*>(a2 + q * 2048 + row * 16) = o;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = tc::pack_bf16_relu(__uint_as_float(r1[8 * q]) * w, __uint_as_float(r1[8 * q + 1]) * w);
          o.y = tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 2]) * w, __uint_as_float(r1[8 * q + 3]) * w);
          o.z = tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 4]) * w, __uint_as_float(r1[8 * q + 5]) * w);
          o.w = tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 6]) * w, __uint_as_float(r1[8 * q + 7]) * w);
          *reinterpret_cast<uint4*>(a2 + (4 + q) * 2048 + row * 16) = o;
        }
        if (g == 0) {                                   // k = 128..130: edge_feat_t (hi, hi, lo) against (b1 hi, b1 lo, b1 hi)
          const __nv_bfloat16 hi = __float2bfloat16_rn(w);
          const float lo = w - __bfloat162float(hi);
          const uint32_t hi16 = static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(&hi));
          *reinterpret_cast<uint4*>(smem + OFF_A2 + 16 * 2048 + row * 16) =
              make_uint4(hi16 | (hi16 << 16), tc::pack_bf16(lo, 0.f), 0u, 0u);
          *reinterpret_cast<uint4*>(smem + OFF_A2 + 17 * 2048 + row * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        tc::fence_proxy_async_smem();
        arrive(bars + B_A2FULL);
      }
      // ---- scatter operands: HblkT (A, scratch) from a raw incidence copy (A2 region), efT (B)
      HF64_TRACE(5);
      wait(B_EFFULL);
      HF64_TRACE(6);
      tc::fence_after_thread_sync();
      stage_raw_H(raw, a.H, a.hstride, b0s, ns, N, ldr, tid);
      drain_bar();
      {
        const bool valid = r128 < rows_used;             // r128 = node row here
        const float* Hcol = raw + sc_r * N * ldr + in_r;
#pragma unroll 2
        for (int kg = half * 8; kg < half * 8 + 8; ++kg) {
          float v[8];
          if (!valid || kg * 8 + 7 < sc_r * N || kg * 8 >= sc_r * N + N) {   // k-group outside this row's scene block
            *reinterpret_cast<uint4*>(smem + OFF_SCR + kg * 2048 + r128 * 16) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int e = kg * 8 + i - sc_r * N;
            v[i] = (e >= 0 && e < N) ? Hcol[e * ldr] : 0.f;
          }
          *reinterpret_cast<uint4*>(smem + OFF_SCR + kg * 2048 + r128 * 16) =
              make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                         tc::pack_bf16(v[6], v[7]));
        }
      }
      {
        float v[32];
        tc::tmem_ld32(tmem + lane_addr + TM_S + g * 32, v);
        // rows (2m, 2m+1) share a 32-bit word of the K-major operand: pair up with the neighbour lane
        unsigned char* dst = smem + OFF_HT + (row >> 3) * 1024 + ((row & 7) >> 1) * 4 + (g * 32 + (lane & 1)) * 16;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float send = (lane & 1) ? v[i] : v[i + 1];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
          const uint32_t word = (lane & 1) ? tc::pack_bf16(recv, v[i + 1]) : tc::pack_bf16(v[i], recv);
          *reinterpret_cast<uint32_t*>(dst + i * 16) = word;
        }
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      arrive(bars + B_EFTREADY);
      HF64_TRACE(7);
      // ---- closing MLP input [agg | h] / N (+ ones) -> A2 region (the raw incidence copy there is dead)
      drain_bar();
      {
        const int r8 = lane & 7, kq = lane >> 3;          // a warp pass: 8 rows x 4 k-groups, 128 B per row
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int combo = warp * 4 + it;                // (row block 0..15) x (k-group block 0..1)
          const int r = (combo >> 1) * 8 + r8, kg = (combo & 1) * 4 + kq;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
          if (r < rows_used) {
            const float* src = a.h + (static_cast<size_t>(b0s) * N + r) * D + kg * 8;
            x = ldg_f4(src); y = ldg_f4(src + 4);
          }
          *reinterpret_cast<uint4*>(smem + OFF_A2 + (8 + kg) * 2048 + r * 16) =
              make_uint4(tc::pack_bf16(x.x * inv_n, x.y * inv_n), tc::pack_bf16(x.z * inv_n, x.w * inv_n),
                         tc::pack_bf16(y.x * inv_n, y.y * inv_n), tc::pack_bf16(y.z * inv_n, y.w * inv_n));
        }
        if (tid < 128) {
          *reinterpret_cast<uint4*>(smem + OFF_A2 + 16 * 2048 + tid * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(smem + OFF_A2 + 17 * 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      HF64_TRACE(8);
      wait(B_AGGFULL);
      HF64_TRACE(9);
      tc::fence_after_thread_sync();
      {
        float v[32];
        tc::tmem_ld32(tmem + lane_addr + TM_S + g * 32, v);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(smem + OFF_A2 + (g * 4 + q) * 2048 + row * 16) =
              make_uint4(tc::pack_bf16(v[8 * q] * inv_n, v[8 * q + 1] * inv_n),
                         tc::pack_bf16(v[8 * q + 2] * inv_n, v[8 * q + 3] * inv_n),
                         tc::pack_bf16(v[8 * q + 4] * inv_n, v[8 * q + 5] * inv_n),
                         tc::pack_bf16(v[8 * q + 6] * inv_n, v[8 * q + 7] * inv_n));
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      arrive(bars + B_PAREADY);
      HF64_TRACE(10);
      // ---- o1: ReLU -> bf16 A operand (over inc; the ones k-groups stay)
      wait(B_O1FULL);
      HF64_TRACE(11);
      tc::fence_after_thread_sync();
      {
        uint32_t r0[32], r1[32];
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_HID + g * 64, r0);
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_HID + g * 64 + 32, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(smem + OFF_A2 + (g * 8 + q) * 2048 + row * 16) = make_uint4(
              tc::pack_bf16_relu(__uint_as_float(r0[8 * q]), __uint_as_float(r0[8 * q + 1])),
              tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 2]), __uint_as_float(r0[8 * q + 3])),
              tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 4]), __uint_as_float(r0[8 * q + 5])),
              tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 6]), __uint_as_float(r0[8 * q + 7])));
          *reinterpret_cast<uint4*>(smem + OFF_A2 + (g * 8 + 4 + q) * 2048 + row * 16) = make_uint4(
              tc::pack_bf16_relu(__uint_as_float(r1[8 * q]), __uint_as_float(r1[8 * q + 1])),
              tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 2]), __uint_as_float(r1[8 * q + 3])),
              tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 4]), __uint_as_float(r1[8 * q + 5])),
              tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 6]), __uint_as_float(r1[8 * q + 7])));
        }
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      arrive(bars + B_O1READY);
      HF64_TRACE(12);
      // ---- node_feat rows: 32-column chunk g, transposed through scratch for 128-byte row segments
      wait(B_OUTFULL);
      HF64_TRACE(13);
      tc::fence_after_thread_sync();
      if (g * 32 < a.Dout) {
        float* tb = reinterpret_cast<float*>(smem + OFF_SCR) + warp * (32 * 36);
        const int rr = lane >> 3, c4 = (lane & 7) * 4;
        const int wrow0 = (warp & 3) * 32;
        float v[32];
        tc::tmem_ld32(tmem + lane_addr + TM_S + g * 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(tb + lane * 36 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rr;
          if (wrow0 + r < rows_used)
            *reinterpret_cast<float4*>(a.node_out + (static_cast<size_t>(b0s) * N + wrow0 + r) * a.ld_out + g * 32 + c4) =
                *reinterpret_cast<const float4*>(tb + r * 36 + c4);
        }
      }
      tc::fence_before_thread_sync();
      HF64_TRACE(14);
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc::tmem_dealloc(tmem, 256);
  }
}

bool hyper_fused64_fits(int N, int E, int D, int T, int Dout, long long ld_out) {
  return D == hf64::D && E == N && N >= 2 && N <= 64 && T >= 1 && T <= 15 && (Dout == 32 || Dout == 64) &&
         (ld_out & 3) == 0;
}

int launch_hyper_fused64_tc(const float* h, const float* H, const float* edge_feat, int B, int N, int T,
                            long long hstride, const gn_stage_weights* w, float* node_out, long long ld_out,
                            int Dout, cudaStream_t st) {
  if (!w->tc_hfuse_w) return GN_E_NULL;
  if (B <= 0) return GN_OK;
  HyperFused64Args a;
  a.h = h; a.H = H; a.edge_feat = edge_feat;
  a.wstream = static_cast<const unsigned char*>(w->tc_hfuse_w);
  a.node_out = node_out; a.ld_out = ld_out; a.Dout = Dout; a.B = B; a.N = N; a.T = T; a.hstride = hstride;
  a.trace = g_trace_buffer;
  const int SC = 128 / N;
  const int ntiles = (B + SC - 1) / SC;
  const int grid = ntiles < 2 * GN_SM_COUNT ? ntiles : 2 * GN_SM_COUNT;
  cudaError_t e = cudaFuncSetAttribute(hyper_fused64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(hf64::SMEM_BYTES));
  if (e != cudaSuccess) return static_cast<int>(e);
  { ProfScope ps__("hyper_fused64_tc", st);
    hyper_fused64_tc_kernel<<<grid, hf64::THREADS, hf64::SMEM_BYTES, st>>>(a); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
