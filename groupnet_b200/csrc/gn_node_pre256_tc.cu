// Node-level prologue for wide layers (h_dim = 256) on tcgen05, one fused chain per 128-row tile:
//
//   hid = relu(h W0^T + b0)      256 -> 256     node2edge_start_mlp.layers.0  (model/MS_HGNN_batch.py:84,:127)
//   x'  = hid W1^T + b1          256 -> 64      node2edge_start_mlp.layers.1
//   pq  = x' [Wp | Wq]^T          64 -> 64      attention_mlp.layers.0 split into its two 64-column halves (:80,:133)
//
// The generic path ran three tc_linear launches with hid (bf16) and x' round-tripping through HBM.  Here
// only h is read and x', pq are written.  Warp-specialised like gn_hyper_fused_tc.cu (320 threads):
//   warps 0-3  drain: TMEM -> relu/bf16 -> next A operand (in place over the h tile), x'/pq -> global
//   warps 4-7  load the NEXT tile's h rows (fp32 -> bf16 canonical operand) into the other A buffer
//   warp 8     streams the weight chunks with cp.async.bulk through a 2-stage ring (one linear stream)
//   warp 9     issues every tcgen05.mma
// Biases ride through the MMA (ones[128 x 16] x [bias_hi, bias_lo]).  Weight stream per tile: 178 KB
// from L2; h tile: 128 KB from HBM -> the kernel is bounded by HBM/L2 bandwidth, not the tensor pipe.
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {
namespace np2 {
constexpr int THREADS = 320;
constexpr uint32_t A_BYTES = 128 * 256 * 2;
constexpr uint32_t OFF_A = 0;                          // two A buffers
constexpr uint32_t OFF_ONES = 2 * A_BYTES;
constexpr uint32_t OFF_RING = OFF_ONES + 4096;
constexpr uint32_t STAGE = 40960;
constexpr uint32_t OFF_BAR = OFF_RING + 2 * STAGE;
constexpr uint32_t W0K = 256 * 64 * 2;                 // W0[:, 64c : 64c+64]
constexpr uint32_t W0KB = W0K + 256 * 16 * 2;          // last K chunk + bias block
constexpr uint32_t W1B = 64 * 272 * 2;                 // W1 + bias block
constexpr uint32_t WPQ = 64 * 64 * 2;
enum { B_WFULL = 0, B_WEMPTY = 2, B_AFULL = 4, B_AFREE = 6, B_HIDFULL = 8, B_HIDREADY = 9, B_XFULL = 10,
       B_XREADY = 11, B_PQFULL = 12, NBAR = 13 };
constexpr uint32_t SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "node_pre256: shared memory budget");
constexpr uint32_t TM_HID = 0, TM_X = 256, TM_PQ = 320;
}  // namespace np2

struct NodePre256Args {
  const float* h; const unsigned char* wstream; float* xprime; float* pq; long long R;
};

__device__ __forceinline__ void np_arrive(uint64_t* b) { tc::mbar_arrive(b); }
__device__ __forceinline__ void np_expect_tx(uint64_t* b, uint32_t bytes) { tcu::expect_tx(b, bytes); }
__device__ __forceinline__ void np_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  tcu::bulk_g2s(dst, src, bytes, bar);
}

__global__ void __launch_bounds__(np2::THREADS, 1)
node_pre256_tc_kernel(NodePre256Args a) {
  using namespace np2;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + NBAR * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < NBAR; ++i) {
      uint32_t cnt = 1;
      if (i == B_AFULL || i == B_AFULL + 1 || i == B_HIDREADY || i == B_XREADY) cnt = 128;
      tc::mbar_init(bars + i, cnt);
    }
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, 512);
  if (tid < 128) tc::build_ones_operand(smem + OFF_ONES, tid, 128);
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t sbase = tc::smem_u32(smem);
  const long long ntiles = (a.R + 127) / 128;

  if (warp == 8) {
    // ------------------------------------------------------------------ weight stream producer
    {                                                   // all 32 lanes: warp-uniform control flow, elected issue
      uint32_t ph_empty = 0x3u;
      int stage = 0;
      auto load = [&](const unsigned char*& src, uint32_t bytes) {
        tc::mbar_wait(bars + B_WEMPTY + stage, (ph_empty >> stage) & 1u);
        ph_empty ^= 1u << stage;
        np_expect_tx(bars + B_WFULL + stage, bytes);
        np_bulk_g2s(sbase + OFF_RING + stage * STAGE, src, bytes, bars + B_WFULL + stage);
        src += bytes;
        stage ^= 1;
      };
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const unsigned char* src = a.wstream;
        load(src, W0K); load(src, W0K); load(src, W0K); load(src, W0KB);
        load(src, W1B);
        load(src, WPQ);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    {                                                   // all 32 lanes: warp-uniform control flow, elected issue
      uint32_t ph = 0u;
      int stage = 0, buf = 0;
      auto wait = [&](int i) { tc::mbar_wait(bars + i, (ph >> i) & 1u); ph ^= 1u << i; };
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const uint32_t abuf = sbase + OFF_A + buf * A_BYTES;
        wait(B_AFULL + buf);
        tc::fence_after_thread_sync();
        for (int kc = 0; kc < 4; ++kc) {
          wait(B_WFULL + stage);
          tc::fence_after_thread_sync();
          const uint32_t wb = sbase + OFF_RING + stage * STAGE;
          tcu::issue_gemm(tmem + TM_HID, abuf + kc * 8 * 2048, wb, 256, 64, kc > 0);
          if (kc == 3) tcu::issue_gemm(tmem + TM_HID, sbase + OFF_ONES, wb + W0K, 256, 16, true);
          tcu::mma_commit(bars + B_WEMPTY + stage);
          stage ^= 1;
        }
        tcu::mma_commit(bars + B_HIDFULL);
        wait(B_HIDREADY);
        wait(B_WFULL + stage);
        tc::fence_after_thread_sync();
        {
          const uint32_t wb = sbase + OFF_RING + stage * STAGE;
          tcu::issue_gemm(tmem + TM_X, abuf, wb, 64, 256, false);
          tcu::issue_gemm(tmem + TM_X, sbase + OFF_ONES, wb + 64 * 256 * 2, 64, 16, true);
          tcu::mma_commit(bars + B_WEMPTY + stage);
          stage ^= 1;
        }
        tcu::mma_commit(bars + B_XFULL);
        wait(B_XREADY);
        wait(B_WFULL + stage);
        tc::fence_after_thread_sync();
        tcu::issue_gemm(tmem + TM_PQ, abuf, sbase + OFF_RING + stage * STAGE, 64, 64, false);
        tcu::mma_commit(bars + B_WEMPTY + stage);
        stage ^= 1;
        tcu::mma_commit(bars + B_PQFULL);
        tcu::mma_commit(bars + B_AFREE + buf);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ h tile loaders (128 threads)
    const int lw = warp - 4;
    const int r8 = lane & 7, kq = lane >> 3;
    uint32_t ph_free = 0x3u;
    int buf = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
      tc::mbar_wait(bars + B_AFREE + buf, (ph_free >> buf) & 1u);
      ph_free ^= 1u << buf;
      unsigned char* A = smem + OFF_A + buf * A_BYTES;
      const long long row0 = tile * 128;
      // a warp pass covers 8 rows x 4 k-groups: 128 contiguous bytes per row from HBM, 16-byte stores
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {
        const int combo = lw * 32 + it;               // 0..127 = (row block 0..15) x (k-group block 0..7)
        const int r = (combo >> 3) * 8 + r8, kg = (combo & 7) * 4 + kq;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
        if (row0 + r < a.R) {
          const float* src = a.h + static_cast<size_t>(row0 + r) * 256 + kg * 8;
          x = ldg_f4(src); y = ldg_f4(src + 4);
        }
        *reinterpret_cast<uint4*>(A + kg * 2048 + r * 16) =
            make_uint4(tc::pack_bf16(x.x, x.y), tc::pack_bf16(x.z, x.w), tc::pack_bf16(y.x, y.y), tc::pack_bf16(y.z, y.w));
      }
      tc::fence_proxy_async_smem();
      np_arrive(bars + B_AFULL + buf);
    }
  } else {
    // ------------------------------------------------------------------ drain warps (TMEM lanes 0..127)
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t ph = 0u;
    auto wait = [&](int i) { tc::mbar_wait(bars + i, (ph >> i) & 1u); ph ^= 1u << i; };
    int buf = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
      unsigned char* A = smem + OFF_A + buf * A_BYTES;
      const long long grow = tile * 128 + row;
      const bool live = grow < a.R;
      // hid: relu -> bf16 A operand, in place over the h tile
      wait(B_HIDFULL);
      tc::fence_after_thread_sync();
#pragma unroll 1
      for (int cc = 0; cc < 8; cc += 2) {
        uint32_t r0[32], r1[32];
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_HID + cc * 32, r0);
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_HID + cc * 32 + 32, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(A + (cc * 4 + q) * 2048 + row * 16) = make_uint4(
              tc::pack_bf16_relu(__uint_as_float(r0[8 * q]), __uint_as_float(r0[8 * q + 1])),
              tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 2]), __uint_as_float(r0[8 * q + 3])),
              tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 4]), __uint_as_float(r0[8 * q + 5])),
              tc::pack_bf16_relu(__uint_as_float(r0[8 * q + 6]), __uint_as_float(r0[8 * q + 7])));
          *reinterpret_cast<uint4*>(A + (cc * 4 + 4 + q) * 2048 + row * 16) = make_uint4(
              tc::pack_bf16_relu(__uint_as_float(r1[8 * q]), __uint_as_float(r1[8 * q + 1])),
              tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 2]), __uint_as_float(r1[8 * q + 3])),
              tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 4]), __uint_as_float(r1[8 * q + 5])),
              tc::pack_bf16_relu(__uint_as_float(r1[8 * q + 6]), __uint_as_float(r1[8 * q + 7])));
        }
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      np_arrive(bars + B_HIDREADY);
      // x': fp32 row to HBM + bf16 operand (k-groups 0..7 of the same buffer)
      wait(B_XFULL);
      tc::fence_after_thread_sync();
      {
        uint32_t r0[32], r1[32];
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_X, r0);
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_X + 32, r1);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<uint4*>(A + q * 2048 + row * 16) = make_uint4(
              tc::pack_bf16_fast(__uint_as_float(r0[8 * q]), __uint_as_float(r0[8 * q + 1])),
              tc::pack_bf16_fast(__uint_as_float(r0[8 * q + 2]), __uint_as_float(r0[8 * q + 3])),
              tc::pack_bf16_fast(__uint_as_float(r0[8 * q + 4]), __uint_as_float(r0[8 * q + 5])),
              tc::pack_bf16_fast(__uint_as_float(r0[8 * q + 6]), __uint_as_float(r0[8 * q + 7])));
          *reinterpret_cast<uint4*>(A + (4 + q) * 2048 + row * 16) = make_uint4(
              tc::pack_bf16_fast(__uint_as_float(r1[8 * q]), __uint_as_float(r1[8 * q + 1])),
              tc::pack_bf16_fast(__uint_as_float(r1[8 * q + 2]), __uint_as_float(r1[8 * q + 3])),
              tc::pack_bf16_fast(__uint_as_float(r1[8 * q + 4]), __uint_as_float(r1[8 * q + 5])),
              tc::pack_bf16_fast(__uint_as_float(r1[8 * q + 6]), __uint_as_float(r1[8 * q + 7])));
        }
        tc::fence_proxy_async_smem();
        tc::fence_before_thread_sync();
        np_arrive(bars + B_XREADY);
        if (live) {
          float* dst = a.xprime + static_cast<size_t>(grow) * 64;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            *reinterpret_cast<uint4*>(dst + 4 * q) = make_uint4(r0[4 * q], r0[4 * q + 1], r0[4 * q + 2], r0[4 * q + 3]);
            *reinterpret_cast<uint4*>(dst + 32 + 4 * q) = make_uint4(r1[4 * q], r1[4 * q + 1], r1[4 * q + 2], r1[4 * q + 3]);
          }
        }
      }
      // pq
      wait(B_PQFULL);
      tc::fence_after_thread_sync();
      {
        uint32_t r0[32], r1[32];
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_PQ, r0);
        tc::tmem_ld32_nowait(tmem + lane_addr + TM_PQ + 32, r1);
        tc::tmem_ld_wait();
        tc::fence_before_thread_sync();
        if (live) {
          float* dst = a.pq + static_cast<size_t>(grow) * 64;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            *reinterpret_cast<uint4*>(dst + 4 * q) = make_uint4(r0[4 * q], r0[4 * q + 1], r0[4 * q + 2], r0[4 * q + 3]);
            *reinterpret_cast<uint4*>(dst + 32 + 4 * q) = make_uint4(r1[4 * q], r1[4 * q + 1], r1[4 * q + 2], r1[4 * q + 3]);
          }
        }
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    tc::tmem_dealloc(tmem, 512);
  }
}

bool node_pre256_fits(int D) { return D == 256; }

int launch_node_pre256_tc(const float* h, long long R, const gn_stage_weights* w, float* xprime, float* pq,
                          cudaStream_t st) {
  if (!w->tc_npre_w) return GN_E_NULL;
  if (R <= 0) return GN_OK;
  NodePre256Args a;
  a.h = h; a.wstream = static_cast<const unsigned char*>(w->tc_npre_w); a.xprime = xprime; a.pq = pq; a.R = R;
  const long long ntiles = (R + 127) / 128;
  const int grid = ntiles < GN_SM_COUNT ? static_cast<int>(ntiles) : GN_SM_COUNT;
  cudaError_t e = cudaFuncSetAttribute(node_pre256_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(np2::SMEM_BYTES));
  if (e != cudaSuccess) return static_cast<int>(e);
  { ProfScope ps__("node_pre256_tc", st);
    node_pre256_tc_kernel<<<grid, np2::THREADS, np2::SMEM_BYTES, st>>>(a); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
