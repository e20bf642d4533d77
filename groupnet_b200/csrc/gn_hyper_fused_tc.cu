// Fused hyper edge_aggregation for wide layers (h_dim = 256, E == N <= 64): the crowd shape.
//
//   eo   = H @ h                                    (model/MS_HGNN_batch.py:263, edges <- nodes)
//   ef   = sum_t edge_feat[:, t] * (W1_t relu(W0_t eo + b0_t) + b1_t)      (:264-265, T MLPs 256->128->256)
//   agg  = H^T @ ef                                 (:267, nodes <- edges)
//
// One persistent CTA per SM walks 128-row tiles (SC = 128 / N whole scenes: edge rows and node rows
// of a tile belong to the same scenes, so both incidence products are tile-local).  Everything runs
// on tcgen05 with fp32 accumulators in TMEM; the only HBM traffic is h, H, edge_feat in and agg out
// (the generic path wrote and re-read eo, a (B*E, T*128) bf16 hidden tensor and ef).
//
//   gather   eo[128 x 256]   = Hblk[128 x 128] * hT[256 x 128]^T        (block-diagonal incidence)
//   per t    hid[128 x 128]  = [eo | 1][128 x 272] * [W0_t | b0_t]^T     (bias through the MMA; two K chunks)
//            ef[128 x 256]  += A2_j[128 x 64(+16)] * [W1_tj (| b1_t)]^T  A2_j = bf16(relu(hid[:, 64j:64j+64]) * edge_feat_t)
//   scatter  agg[128 x 256]  = HblkT[128 x 128] * efT[256 x 128]^T
//
// Warp roles (320 threads): warps 0-3 / 4-7 drain the two hidden halves (TMEM -> bf16 A2 operand),
// warp 8 streams the weight chunks with cp.async.bulk (TMA bulk copy, mbarrier complete_tx) through a
// 3-stage ring in the order the MMAs consume them (host-packed as ONE linear stream, packing.py),
// warp 9 issues every tcgen05.mma.  G1 of step t+1 is issued before G2 of step t, so the drains
// overlap the tensor pipe.  Per 128-row tile the weight stream is 1.41 MB from L2: the kernel is
// L2->SM bandwidth bound (~42 B/clk/SM), not tensor-pipe bound; see DESIGN.md.
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {
namespace hf {
constexpr int D = 256;
constexpr int THREADS = 320;
constexpr uint32_t W0A_CHUNK = 128 * 128 * 2;    // W0_t[:, 0:128]
constexpr uint32_t W0B_CHUNK = 128 * 144 * 2;    // W0_t[:, 128:256] | 16 bias k-columns
constexpr uint32_t W1A_CHUNK = 256 * 80 * 2;     // 256 outputs x (64 hidden + 16 b1 k-columns)
constexpr uint32_t W1B_CHUNK = 256 * 64 * 2;
constexpr uint32_t STAGE = 40960;
constexpr int NSTAGE = 3;
constexpr uint32_t OFF_EO = 0;                   // eo A operand [128 x 256]; epilogue: efT B operand [256 x 128]
constexpr uint32_t OFF_ONES = 65536;             // k-groups 32,33 of the A operand: 1.0, 1.0, 0...
constexpr uint32_t OFF_A2 = OFF_ONES + 4096;     // A2_0 [128 x 80] | A2_1 [128 x 64]; staging: Hblk; epilogue: HblkT
constexpr uint32_t A2_0_BYTES = 128 * 80 * 2;
constexpr uint32_t OFF_A2_1 = OFF_A2 + A2_0_BYTES;
constexpr uint32_t OFF_RING = OFF_A2 + A2_0_BYTES + 128 * 64 * 2;     // 3 stages; staging: hT [256 x 128]
constexpr uint32_t OFF_BAR = OFF_RING + NSTAGE * STAGE;
enum { B_WFULL = 0, B_WEMPTY = 3, B_HFULL = 6, B_HFREE = 10, B_A2FULL = 14, B_A2FREE = 16, B_STAGE = 18,
       B_EOFULL = 19, B_EOREADY = 20, B_EFFULL = 21, B_EFTREADY = 22, B_AGGFULL = 23,
       B_PXFULL = 24, B_PXEMPTY = 25, B_PAREADY = 26, B_O1FULL = 27, B_O1READY = 28, B_OUTFULL = 29, NBAR = 30 };
constexpr uint32_t SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "hyper_fused_tc: shared memory budget");
constexpr uint32_t TM_EF = 0, TM_HB = 256, TM_AGG = 256, TM_O1 = 0, TM_OUT = 256;
// closing MLP (post) weight chunks, streamed after the aggregation chunks: buffers alternate between ring
// stage 2 ("Y") and the A2 region ("X"); the [agg | h]/N operand sits in the eo region and ring stages 0-1
constexpr uint32_t P0_CHUNK = 128 * 128 * 2;            // post_w0[:, 128c : 128c+128]
constexpr uint32_t P0B_CHUNK = P0_CHUNK + 128 * 16 * 2; // last one carries the bias block
constexpr uint32_t OFF_X = OFF_A2, OFF_Y = OFF_RING + 2 * STAGE, OFF_AH = OFF_RING;
static_assert(P0B_CHUNK <= A2_0_BYTES + 128 * 64 * 2, "post chunk must fit the A2 region");
}  // namespace hf

extern unsigned long long* g_trace_buffer;

struct HyperFusedArgs {
  const float* h; const float* H; const float* edge_feat; const unsigned char* wstream;
  float* agg; int B, N, T; long long hstride;
  float* node_out; long long ld_out; int Dout; int post; size_t post_off;
  unsigned long long* trace;          // optional: clock64() per phase, block 0, first tiles (gn_profile_set_trace)
};

// producer-warp forms: every lane calls them, one elected lane issues (gn_tc.cuh, tcu)
using tc::mbar_arrive; using tcu::bulk_g2s; using tc::drain_bar; using tc::stage_raw_H;
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { tcu::expect_tx(b, bytes); }

__global__ void __launch_bounds__(hf::THREADS, 1)
hyper_fused_tc_kernel(HyperFusedArgs a) {
  using namespace hf;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < NBAR; ++i) {
      uint32_t cnt = 1;
      if (i >= B_HFREE && i < B_HFREE + 4) cnt = 256;
      if (i >= B_A2FULL && i < B_A2FULL + 2) cnt = 128;
      if (i == B_STAGE || i == B_EOREADY || i == B_EFTREADY || i == B_PAREADY || i == B_O1READY) cnt = 256;
      tc::mbar_init(bars + i, cnt);
    }
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
  if (tid < 256) tc::build_ones_operand(smem + OFF_ONES, tid, 256);
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sbase = tc::smem_u32(smem);
  const int N = a.N, SC = 128 / N, T = a.T;
  const int ntiles = (a.B + SC - 1) / SC;

  if (warp == 8) {
    // ------------------------------------------------------------------ weight stream producer
    {                                                   // all 32 lanes: warp-uniform control flow, elected issue
      uint32_t ph_empty = 0x7u, ph_eofull = 0u, ph_px = 1u, ph_agg = 0u;
      int stage = 0;
      auto load = [&](const unsigned char*& src, uint32_t bytes) {
        tc::mbar_wait(bars + B_WEMPTY + stage, (ph_empty >> stage) & 1u);
        ph_empty ^= 1u << stage;
        mbar_expect_tx(bars + B_WFULL + stage, bytes);
        bulk_g2s(sbase + OFF_RING + stage * STAGE, src, bytes, bars + B_WFULL + stage);
        src += bytes;
        stage = stage == NSTAGE - 1 ? 0 : stage + 1;
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned char* src = a.wstream;
        stage = NSTAGE - 1;                              // every tile starts in the stage hT does not cover
        load(src, W0A_CHUNK);
        tc::mbar_wait(bars + B_EOFULL, ph_eofull);      // hT (ring stages 0-1) consumed by the gather MMA
        ph_eofull ^= 1u;
        load(src, W0B_CHUNK);
        for (int s = 1; s <= T; ++s) {
          if (s < T) { load(src, W0A_CHUNK); load(src, W0B_CHUNK); }
          load(src, W1A_CHUNK); load(src, W1B_CHUNK);
        }
        if (a.post) {
          const unsigned char* ps = a.wstream + a.post_off;
          const uint32_t c4 = static_cast<uint32_t>(a.Dout) * 160u, c5 = static_cast<uint32_t>(a.Dout) * 128u;
          auto load_y = [&](uint32_t bytes) { stage = NSTAGE - 1; load(ps, bytes); };
          auto load_x = [&](uint32_t bytes) {
            tc::mbar_wait(bars + B_PXEMPTY, ph_px); ph_px ^= 1u;
            mbar_expect_tx(bars + B_PXFULL, bytes);
            bulk_g2s(sbase + OFF_X, ps, bytes, bars + B_PXFULL);
            ps += bytes;
          };
          load_y(P0_CHUNK);
          tc::mbar_wait(bars + B_AGGFULL, ph_agg); ph_agg ^= 1u;   // HblkT (A2 region) consumed
          load_x(P0_CHUNK);
          load_y(P0_CHUNK);
          load_x(P0B_CHUNK);
          load_y(c4);
          load_x(c5);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    {                                                   // all 32 lanes: warp-uniform control flow, elected issue
      uint32_t ph = 0u;                                  // bit i = parity to wait for on barrier i
      ph |= 0xFu << B_HFREE;                             // "free" barriers: the first wait passes
      int stage = 0;
      auto wait = [&](int i) { tc::mbar_wait(bars + i, (ph >> i) & 1u); ph ^= 1u << i; };
      long long wacc[4] = {0, 0, 0, 0};
      int ititer = 0;
      auto wait_t = [&](int i, int slot) {
#ifdef GN_ENABLE_TRACE
        if (a.trace != nullptr) { const long long t0 = clock64(); wait(i); wacc[slot] += clock64() - t0; return; }
#endif
        wait(i);
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ititer) {
#ifdef GN_ENABLE_TRACE
        if (a.trace != nullptr && blockIdx.x == 0 && ititer > 0 && ititer <= 8) {
          for (int k = 0; k < 4; ++k) { a.trace[128 + (ititer - 1) * 4 + k] = wacc[k]; wacc[k] = 0; }
        }
#endif
        wait(B_STAGE);
        tc::fence_after_thread_sync();
        tcu::issue_gemm(tmem + TM_EF, sbase + OFF_A2, sbase + OFF_RING, 256, 128, false);     // eo = Hblk * h
        tcu::mma_commit(bars + B_EOFULL);
        wait(B_EOREADY);
        tc::fence_after_thread_sync();
        stage = NSTAGE - 1;
        for (int s = 0; s <= T; ++s) {
          if (s < T) {                                   // hid[128 x 128] = [eo | 1] * [W0_s | b0_s]^T, two K chunks
            const int p = s & 1;
            wait_t(B_HFREE + p, 0);
            wait_t(B_WFULL + stage, 1);
            tc::fence_after_thread_sync();
            tcu::issue_gemm(tmem + TM_HB + p * 128, sbase + OFF_EO, sbase + OFF_RING + stage * STAGE, 128, 128, false);
            tcu::mma_commit(bars + B_WEMPTY + stage);
            stage = stage == NSTAGE - 1 ? 0 : stage + 1;
            wait_t(B_WFULL + stage, 1);
            tc::fence_after_thread_sync();
            tcu::issue_gemm(tmem + TM_HB + p * 128, sbase + OFF_EO + 16 * 2048, sbase + OFF_RING + stage * STAGE,
                           128, 144, true);
            tcu::mma_commit(bars + B_WEMPTY + stage);
            tcu::mma_commit(bars + B_HFULL + p);
            stage = stage == NSTAGE - 1 ? 0 : stage + 1;
          }
          if (s >= 1) {
            for (int j = 0; j < 2; ++j) {
              wait_t(B_A2FULL + j, 2);
              wait_t(B_WFULL + stage, 3);
              tc::fence_after_thread_sync();
              tcu::issue_gemm(tmem + TM_EF, sbase + (j ? OFF_A2_1 : OFF_A2), sbase + OFF_RING + stage * STAGE,
                             256, j ? 64 : 80, !(s == 1 && j == 0));
              tcu::mma_commit(bars + B_WEMPTY + stage);
              tcu::mma_commit(bars + B_A2FREE + j);
              stage = stage == NSTAGE - 1 ? 0 : stage + 1;
            }
          }
        }
        tcu::mma_commit(bars + B_EFFULL);
        wait(B_EFTREADY);
        tc::fence_after_thread_sync();
        tcu::issue_gemm(tmem + TM_AGG, sbase + OFF_A2, sbase + OFF_EO, 256, 128, false);       // agg = HblkT * ef
        tcu::mma_commit(bars + B_AGGFULL);
        if (a.post) {
          // o1 = relu([agg | h]/N W0^T + b0): four K chunks of 128 (A: eo region, then ring stages 0-1)
          wait(B_PAREADY);
          wait(B_WFULL + 2);
          tc::fence_after_thread_sync();
          tcu::issue_gemm(tmem + TM_O1, sbase + OFF_EO, sbase + OFF_Y, 128, 128, false);
          tcu::mma_commit(bars + B_WEMPTY + 2);
          wait(B_PXFULL);
          tc::fence_after_thread_sync();
          tcu::issue_gemm(tmem + TM_O1, sbase + OFF_EO + 16 * 2048, sbase + OFF_X, 128, 128, true);
          tcu::mma_commit(bars + B_PXEMPTY);
          wait(B_WFULL + 2);
          tc::fence_after_thread_sync();
          tcu::issue_gemm(tmem + TM_O1, sbase + OFF_AH, sbase + OFF_Y, 128, 128, true);
          tcu::mma_commit(bars + B_WEMPTY + 2);
          wait(B_PXFULL);
          tc::fence_after_thread_sync();
          tcu::issue_gemm(tmem + TM_O1, sbase + OFF_AH + 16 * 2048, sbase + OFF_X, 128, 128, true);
          tcu::issue_gemm(tmem + TM_O1, sbase + OFF_ONES, sbase + OFF_X + P0_CHUNK, 128, 16, true);
          tcu::mma_commit(bars + B_PXEMPTY);
          tcu::mma_commit(bars + B_O1FULL);
          // out = o1 W1^T + b1: two K chunks of 64
          wait(B_O1READY);
          wait(B_WFULL + 2);
          tc::fence_after_thread_sync();
          tcu::issue_gemm(tmem + TM_OUT, sbase + OFF_EO, sbase + OFF_Y, a.Dout, 64, false);
          tcu::issue_gemm(tmem + TM_OUT, sbase + OFF_ONES, sbase + OFF_Y + static_cast<uint32_t>(a.Dout) * 128u,
                         a.Dout, 16, true);
          tcu::mma_commit(bars + B_WEMPTY + 2);
          wait(B_PXFULL);
          tc::fence_after_thread_sync();
          tcu::issue_gemm(tmem + TM_OUT, sbase + OFF_EO + 8 * 2048, sbase + OFF_X, a.Dout, 64, true);
          tcu::mma_commit(bars + B_PXEMPTY);
          tcu::mma_commit(bars + B_OUTFULL);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ staging / drain warps
    const int g = warp >> 2;                             // hidden half / column half
    const int row = (warp & 3) * 32 + lane;              // TMEM lane = tile row
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t ph = 0u;
    ph |= 0x3u << B_A2FREE;
    auto wait = [&](int i) { tc::mbar_wait(bars + i, (ph >> i) & 1u); ph ^= 1u << i; };
    const int r128 = tid & 127, half = tid >> 7;
    const int sc_r = r128 / N, in_r = r128 - sc_r * N;   // scene-in-tile and index-in-scene of row r128
    const int ldr = ((N + 3) & ~3) + 4;                  // row stride of the raw incidence staging
    int titer = 0;
#ifdef GN_ENABLE_TRACE
#define HF_TRACE(pt) do { if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && titer < 8) \
    a.trace[titer * 16 + (pt)] = clock64(); } while (0)
#else
#define HF_TRACE(pt) do { } while (0)
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++titer) {
      const int b0s = tile * SC;
      HF_TRACE(0);
      const int ns = min(SC, a.B - b0s);
      const int rows_used = ns * N;
      // ---- Hblk[edge row][node column] (A operand, block diagonal): raw incidence rows staged coalesced in
      //      the (idle) eo region, then each thread packs the k-groups of its row
      {
        float* raw = reinterpret_cast<float*>(smem + OFF_EO);
        drain_bar();                                      // the previous tile's output transposes used this region
        stage_raw_H(raw, a.H, a.hstride, b0s, ns, N, ldr, tid);
        drain_bar();
        const bool valid = r128 < rows_used;
        const float* Hrow = raw + r128 * ldr;
#pragma unroll 4
        for (int kg = half * 8; kg < half * 8 + 8; ++kg) {
          float v[8];
          const int n0 = kg * 8 - sc_r * N;
          if (!valid || n0 + 7 < 0 || n0 >= N) {           // k-group outside this row's scene block
            *reinterpret_cast<uint4*>(smem + OFF_A2 + kg * 2048 + r128 * 16) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
          if (n0 >= 0 && n0 + 8 <= N && (N & 3) == 0) {
            const float4 x = *reinterpret_cast<const float4*>(Hrow + n0), y = *reinterpret_cast<const float4*>(Hrow + n0 + 4);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int n = n0 + i;
              v[i] = (n >= 0 && n < N) ? Hrow[n] : 0.f;
            }
          }
          *reinterpret_cast<uint4*>(smem + OFF_A2 + kg * 2048 + r128 * 16) =
              make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                         tc::pack_bf16(v[6], v[7]));
        }
      }
      HF_TRACE(1);
      // ---- hT[column c][node k] (B operand): thread = column, 8 nodes per 16-byte store
      {
        const float* hsrc = a.h + static_cast<size_t>(b0s) * N * D + tid;
#pragma unroll 4
        for (int kg = 0; kg < 16; ++kg) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int node = kg * 8 + i;
            v[i] = (node < rows_used) ? __ldg(hsrc + static_cast<size_t>(node) * D) : 0.f;
          }
          *reinterpret_cast<uint4*>(smem + OFF_RING + kg * 4096 + tid * 16) =
              make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                         tc::pack_bf16(v[6], v[7]));
        }
      }
      tc::fence_proxy_async_smem();
      mbar_arrive(bars + B_STAGE);
      HF_TRACE(2);
      const bool rowvalid = row < rows_used;
      const float* efrow = a.edge_feat + (static_cast<size_t>(b0s) * N + row) * T;
      float wnext = rowvalid ? __ldg(efrow) : 0.f;
      // ---- eo: TMEM -> bf16 A operand, this group's 128 columns
      wait(B_EOFULL);
      HF_TRACE(3);
      tc::fence_after_thread_sync();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float v[32];
        tc::tmem_ld32(tmem + lane_addr + TM_EF + g * 128 + cc * 32, v);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(smem + OFF_EO + (g * 16 + cc * 4 + q) * 2048 + row * 16) =
              make_uint4(tc::pack_bf16_fast(v[8 * q], v[8 * q + 1]), tc::pack_bf16_fast(v[8 * q + 2], v[8 * q + 3]),
                         tc::pack_bf16_fast(v[8 * q + 4], v[8 * q + 5]), tc::pack_bf16_fast(v[8 * q + 6], v[8 * q + 7]));
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      mbar_arrive(bars + B_EOREADY);
      HF_TRACE(4);
      // ---- main loop: hidden half g of step t -> A2_g
      unsigned char* a2 = smem + (g ? OFF_A2_1 : OFF_A2);
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        const int p = t & 1;
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
        wait(B_HFULL + p);
        tc::fence_after_thread_sync();
        uint32_t r0[32], r1[32];
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_HB + p * 128 + g * 64, r0);
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_HB + p * 128 + g * 64 + 32, r1);
        tc::tmem_ld_wait();
        tc::fence_before_thread_sync();
        mbar_arrive(bars + B_HFREE + p);
        wait(B_A2FREE + g);
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
        if (g == 0) {                                   // k = 64..66: edge_feat_t (hi, hi, lo) against (b1 hi, b1 lo, b1 hi)
          const __nv_bfloat16 hi = __float2bfloat16_rn(w);
          const float lo = w - __bfloat162float(hi);
          const uint32_t hi16 = static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(&hi));
          *reinterpret_cast<uint4*>(a2 + 8 * 2048 + row * 16) =
              make_uint4(hi16 | (hi16 << 16), tc::pack_bf16(lo, 0.f), 0u, 0u);
          *reinterpret_cast<uint4*>(a2 + 9 * 2048 + row * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        tc::fence_proxy_async_smem();
        mbar_arrive(bars + B_A2FULL + g);
      }
      // ---- epilogue: HblkT (A), efT (B), then agg out
      HF_TRACE(5);
      wait(B_EFFULL);
      HF_TRACE(6);
      tc::fence_after_thread_sync();
      {
        float* raw = reinterpret_cast<float*>(smem + OFF_AH);     // ring stages 0-1 are idle since ef_full
        stage_raw_H(raw, a.H, a.hstride, b0s, ns, N, ldr, tid);
        drain_bar();
        const bool valid = r128 < rows_used;             // r128 = node row here
        const float* Hcol = raw + sc_r * N * ldr + in_r;
#pragma unroll 2
        for (int kg = half * 8; kg < half * 8 + 8; ++kg) {
          float v[8];
          if (!valid || kg * 8 + 7 < sc_r * N || kg * 8 >= sc_r * N + N) {   // k-group outside this row's scene block
            *reinterpret_cast<uint4*>(smem + OFF_A2 + kg * 2048 + r128 * 16) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int e = kg * 8 + i - sc_r * N;
            v[i] = (e >= 0 && e < N) ? Hcol[e * ldr] : 0.f;
          }
          *reinterpret_cast<uint4*>(smem + OFF_A2 + kg * 2048 + r128 * 16) =
              make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]), tc::pack_bf16(v[4], v[5]),
                         tc::pack_bf16(v[6], v[7]));
        }
      }
      HF_TRACE(7);
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float v[32];
        tc::tmem_ld32(tmem + lane_addr + TM_EF + g * 128 + cc * 32, v);
        // rows (2m, 2m+1) share a 32-bit word of the K-major operand: exchange with the neighbour lane so each
        // lane stores full words (even lane: even columns, odd lane: odd columns)
        unsigned char* dst = smem + OFF_EO + (row >> 3) * 4096 + ((row & 7) >> 1) * 4 +
                             (g * 128 + cc * 32 + (lane & 1)) * 16;
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
      mbar_arrive(bars + B_EFTREADY);
      HF_TRACE(8);
      if (a.post) {
        const float inv_n = 1.f / static_cast<float>(N);
        // h rows / N -> A operand columns 256..511 (ring stages 0-1; the raw incidence copy there is dead once
        // every thread has built its HblkT rows).  A warp pass covers 8 rows x 4 k-groups: 128 contiguous
        // bytes per row from L2.
        drain_bar();
        {
          const int r8 = lane & 7, kq = lane >> 3;
#pragma unroll 8
          for (int it = 0; it < 16; ++it) {
            const int combo = warp * 16 + it;             // (row block 0..15) x (k-group block 0..7)
            const int r = (combo >> 3) * 8 + r8, kg = (combo & 7) * 4 + kq;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
            if (r < rows_used) {
              const float* src = a.h + (static_cast<size_t>(b0s) * N + r) * D + kg * 8;
              x = ldg_f4(src); y = ldg_f4(src + 4);
            }
            *reinterpret_cast<uint4*>(smem + OFF_AH + kg * 2048 + r * 16) =
                make_uint4(tc::pack_bf16(x.x * inv_n, x.y * inv_n), tc::pack_bf16(x.z * inv_n, x.w * inv_n),
                           tc::pack_bf16(y.x * inv_n, y.y * inv_n), tc::pack_bf16(y.z * inv_n, y.w * inv_n));
          }
        }
        HF_TRACE(9);
        wait(B_AGGFULL);
        HF_TRACE(10);
        tc::fence_after_thread_sync();
        // agg / N -> A operand columns 0..255 (eo region; efT was consumed by the scatter MMA)
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          float v[32];
          tc::tmem_ld32(tmem + lane_addr + TM_AGG + g * 128 + cc * 32, v);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(smem + OFF_EO + (g * 16 + cc * 4 + q) * 2048 + row * 16) =
                make_uint4(tc::pack_bf16(v[8 * q] * inv_n, v[8 * q + 1] * inv_n),
                           tc::pack_bf16(v[8 * q + 2] * inv_n, v[8 * q + 3] * inv_n),
                           tc::pack_bf16(v[8 * q + 4] * inv_n, v[8 * q + 5] * inv_n),
                           tc::pack_bf16(v[8 * q + 6] * inv_n, v[8 * q + 7] * inv_n));
        }
        tc::fence_proxy_async_smem();
        tc::fence_before_thread_sync();
        mbar_arrive(bars + B_PAREADY);
        HF_TRACE(11);
        // o1: relu -> bf16 A operand [128 x 128] (eo region), this group's 64 columns
        wait(B_O1FULL);
        HF_TRACE(12);
        tc::fence_after_thread_sync();
        {
          uint32_t r0[32], r1[32];
          tc::tmem_ld32_nowait(tmem + lane_addr + TM_O1 + g * 64, r0);
          tc::tmem_ld32_nowait(tmem + lane_addr + TM_O1 + g * 64 + 32, r1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<uint4*>(smem + OFF_EO + (g * 8 + q) * 2048 + row * 16) = make_uint4(
                tc::pack_bf16_relu(__uint_as_float(r0[8 * q]), __uint_as_float(r0[8 * q + 1])),
                tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 2]), __uint_as_float(r0[8 * q + 3])),
                tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 4]), __uint_as_float(r0[8 * q + 5])),
                tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 6]), __uint_as_float(r0[8 * q + 7])));
            *reinterpret_cast<uint4*>(smem + OFF_EO + (g * 8 + 4 + q) * 2048 + row * 16) = make_uint4(
                tc::pack_bf16_relu(__uint_as_float(r1[8 * q]), __uint_as_float(r1[8 * q + 1])),
                tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 2]), __uint_as_float(r1[8 * q + 3])),
                tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 4]), __uint_as_float(r1[8 * q + 5])),
                tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 6]), __uint_as_float(r1[8 * q + 7])));
          }
        }
        tc::fence_proxy_async_smem();
        tc::fence_before_thread_sync();
        mbar_arrive(bars + B_O1READY);
        HF_TRACE(13);
        // node_feat rows: 32-column chunks alternate between the two groups
        wait(B_OUTFULL);
        HF_TRACE(14);
        tc::fence_after_thread_sync();
        // transpose each 32 x 32 block through shared memory (the eo region is dead): a warp store
        // instruction then covers 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes
        {
          float* tb = reinterpret_cast<float*>(smem + OFF_EO) + warp * (32 * 36);
          const int rr = lane >> 3, c4 = (lane & 7) * 4;
          const int wrow0 = (warp & 3) * 32;
#pragma unroll 1
          for (int cc = g; cc * 32 < a.Dout; cc += 2) {
            float v[32];
            tc::tmem_ld32(tmem + lane_addr + TM_OUT + cc * 32, v);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(tb + lane * 36 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + rr;
              if (wrow0 + r < rows_used)
                *reinterpret_cast<float4*>(a.node_out + (static_cast<size_t>(b0s) * N + wrow0 + r) * a.ld_out + cc * 32 + c4) =
                    *reinterpret_cast<const float4*>(tb + r * 36 + c4);
            }
          }
        }
        tc::fence_before_thread_sync();
        HF_TRACE(15);
        continue;
      }
      wait(B_AGGFULL);
      tc::fence_after_thread_sync();
      float* out = a.agg + (static_cast<size_t>(b0s) * N + row) * D + g * 128;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float v[32];
        tc::tmem_ld32(tmem + lane_addr + TM_AGG + g * 128 + cc * 32, v);
        if (rowvalid) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(out + cc * 32 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      tc::fence_before_thread_sync();
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc::tmem_dealloc(tmem, 512);
  }
}

bool hyper_fused_fits(int N, int E, int D, int T) {
  return D == hf::D && E == N && N >= 2 && N <= 64 && T >= 1 && T <= 15;
}

bool hyper_fused_post_fits(int Dout, long long ld_out) {
  return Dout >= 32 && Dout <= 256 && (Dout & 31) == 0 && (ld_out & 3) == 0;
}

int launch_hyper_fused_tc(const float* h, const float* H, const float* edge_feat, int B, int N, int T,
                          long long hstride, const gn_stage_weights* w, float* agg,
                          float* node_out, long long ld_out, int Dout, cudaStream_t st) {
  if (!w->tc_hfuse_w) return GN_E_NULL;
  if (B <= 0) return GN_OK;
  HyperFusedArgs a;
  a.h = h; a.H = H; a.edge_feat = edge_feat;
  a.wstream = static_cast<const unsigned char*>(w->tc_hfuse_w);
  a.agg = agg; a.B = B; a.N = N; a.T = T; a.hstride = hstride;
  a.node_out = node_out; a.ld_out = ld_out; a.Dout = Dout;
  a.post = node_out != nullptr ? 1 : 0;
  a.post_off = static_cast<size_t>(T) * (hf::W0A_CHUNK + hf::W0B_CHUNK + hf::W1A_CHUNK + hf::W1B_CHUNK);
  a.trace = g_trace_buffer;
  const int SC = 128 / N;
  const int ntiles = (B + SC - 1) / SC;
  const int grid = ntiles < GN_SM_COUNT ? ntiles : GN_SM_COUNT;
  cudaError_t e = cudaFuncSetAttribute(hyper_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(hf::SMEM_BYTES));
  if (e != cudaSuccess) return static_cast<int>(e);
  { ProfScope ps__("hyper_fused_tc", st);
    hyper_fused_tc_kernel<<<grid, hf::THREADS, hf::SMEM_BYTES, st>>>(a); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
