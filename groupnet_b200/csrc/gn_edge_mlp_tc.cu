// Kernel (c), tensor-core path: MLP_dict_softmax (model/MS_HGNN_batch.py:31-53)
// + Gumbel softmax (:446-520) over tiles of 128 edge rows as a chain of four
// tcgen05.mma GEMMs with fp32 accumulators in TMEM:
//
//   G1  [128 x 64]  x W_init0^T  [64 -> 128]   epilogue: +b, ReLU, bf16 -> smem
//   G2  [128 x 128] x W_init1^T  [128 -> 64]   epilogue: +b,       bf16 -> smem   (z)
//   G3  [128 x 64]  x W_df0^T    [64 -> 256]   epilogue: +b, ReLU, bf16 -> smem   (dist | factor hidden)
//   G4  [128 x 256] x W_df1^T    [256 -> 16]   epilogue: +b, Gumbel softmax over T, sigmoid
//
// One persistent CTA per SM (grid = 148).  All weights (72 KB bf16) stay in
// shared memory for the life of the CTA; activations never leave the SM
// between the four GEMMs.  One thread issues the MMAs (tcgen05.mma is a
// single-thread instruction); completion is signalled through
// tcgen05.commit -> mbarrier; the 8 warps then drain the accumulator with
// tcgen05.ld (warp w owns TMEM lanes 32*(w%4).., i.e. tile rows, and one half
// of the columns), apply the epilogue and write the next A operand.
//
// Roofline: tensor pipe.  Algorithmic work per 128-row tile: 128 * 32,768 MAC
// (+ 128 * 4,096 for the padded N = 16 head) = 9.4 MFLOP; algorithmic HBM
// bytes per row: 256 (edges) + 8 T (dist, edge_feat).
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {

struct TcMlpWeights {
  const void* w1; const void* w2; const void* w3; const void* w4;   // bf16, canonical layout
  const float* b1; const float* b2; const float* b3; const float* b4;
};

namespace tcmlp {
constexpr int TM = 128;
constexpr uint32_t OFF_W1 = 0;                       // N=128 K=64   16 KB
constexpr uint32_t OFF_W2 = OFF_W1 + 128 * 64 * 2;   // N=64  K=128  16 KB
constexpr uint32_t OFF_W3 = OFF_W2 + 64 * 128 * 2;   // N=256 K=64   32 KB
constexpr uint32_t OFF_W4 = OFF_W3 + 256 * 64 * 2;   // N=16  K=256   8 KB
constexpr uint32_t OFF_A0 = OFF_W4 + 16 * 256 * 2;   // [128 x 64]  edges, later z
constexpr uint32_t OFF_A1 = OFF_A0 + TM * 64 * 2;    // [128 x 256] hidden of G1 (128 cols) / G3 (256 cols)
constexpr uint32_t OFF_B = OFF_A1 + TM * 256 * 2;    // biases: 128 + 64 + 256 + 16 floats
constexpr uint32_t OFF_BAR = OFF_B + 464 * 4;        // mbarrier (8 B) + tmem base (4 B)
constexpr uint32_t SMEM_BYTES = OFF_BAR + 16;
constexpr uint32_t TMEM_COLS = 256;
}  // namespace tcmlp

__global__ void __launch_bounds__(GN_THREADS, 1)
edge_mlp_tc_kernel(const float* __restrict__ edges, long long R, int T, int E, TcMlpWeights W,
                   const float* __restrict__ U, int noise_mode, unsigned long long seed,
                   long long scene_offset, int stage_index,
                   float* __restrict__ dist_out, float* __restrict__ edge_feat) {
  using namespace tcmlp;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;           // TMEM lane quarter, column half
  const int row = q * 32 + lane;                    // tile row owned in every epilogue
  float* bias = reinterpret_cast<float*>(smem + OFF_B);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8);

  // ---- one-time setup: weights + biases -> smem, mbarrier, TMEM ----
  {
    const uint4* src[4] = {static_cast<const uint4*>(W.w1), static_cast<const uint4*>(W.w2),
                           static_cast<const uint4*>(W.w3), static_cast<const uint4*>(W.w4)};
    const uint32_t off[4] = {OFF_W1, OFF_W2, OFF_W3, OFF_W4};
    const int n16[4] = {128 * 64 / 8, 64 * 128 / 8, 256 * 64 / 8, 16 * 256 / 8};
#pragma unroll
    for (int m = 0; m < 4; ++m)
      for (int i = tid; i < n16[m]; i += GN_THREADS)
        *reinterpret_cast<uint4*>(smem + off[m] + 16 * i) = __ldg(src[m] + i);
    for (int i = tid; i < 128; i += GN_THREADS) bias[i] = __ldg(W.b1 + i);
    for (int i = tid; i < 64; i += GN_THREADS) bias[128 + i] = __ldg(W.b2 + i);
    for (int i = tid; i < 256; i += GN_THREADS) bias[192 + i] = __ldg(W.b3 + i);
    for (int i = tid; i < 16; i += GN_THREADS) bias[448 + i] = __ldg(W.b4 + i);
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  if (tid == 32) mbar_init(mbar, 1);
  fence_proxy_async_smem();
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
  const uint32_t sbase = smem_u32(smem);
  uint32_t phase = 0;

  const long long ntiles = (R + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    const int nrows = static_cast<int>(min(static_cast<long long>(TM), R - row0));

    // ---- stage the edge tile as bf16 A operand: task = (row, k-group of 8) ----
#pragma unroll
    for (int it = 0; it < (TM * 8) / GN_THREADS; ++it) {
      const int task = it * GN_THREADS + tid;
      const int r = task & (TM - 1), k8 = task >> 7;
      uint4 pk = make_uint4(0u, 0u, 0u, 0u);
      if (r < nrows) {
        const float* src = edges + static_cast<size_t>(row0 + r) * GN_ATT_DIM + k8 * 8;
        float4 a = ldg_stream_f4(src), b = ldg_stream_f4(src + 4);
        pk.x = pack_bf16(a.x, a.y); pk.y = pack_bf16(a.z, a.w);
        pk.z = pack_bf16(b.x, b.y); pk.w = pack_bf16(b.z, b.w);
      }
      *reinterpret_cast<uint4*>(smem + OFF_A0 + canon_off(r, k8, TM)) = pk;
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();

    // ---- G1: 64 -> 128 ----
    if (tid == 0) {
      fence_after_thread_sync();
      issue_gemm(tmem_base, sbase + OFF_A0, sbase + OFF_W1, 128, 64, false);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_thread_sync();
#pragma unroll
    for (int cc = 0; cc < 64; cc += 32) {
      const int c0 = hf * 64 + cc;
      float v[32];
      tmem_ld32(tmem_row + c0, v);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(v[8 * g + j] + bias[c0 + 8 * g + j], 0.f);
        uint4 pk = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                              pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        *reinterpret_cast<uint4*>(smem + OFF_A1 + canon_off(row, (c0 >> 3) + g, TM)) = pk;
      }
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();

    // ---- G2: 128 -> 64 (z) ----
    if (tid == 0) {
      fence_after_thread_sync();
      issue_gemm(tmem_base, sbase + OFF_A1, sbase + OFF_W2, 64, 128, false);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_thread_sync();
    {
      const int c0 = hf * 32;
      float v[32];
      tmem_ld32(tmem_row + c0, v);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[8 * g + j] + bias[128 + c0 + 8 * g + j];
        uint4 pk = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                              pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        *reinterpret_cast<uint4*>(smem + OFF_A0 + canon_off(row, (c0 >> 3) + g, TM)) = pk;
      }
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();

    // ---- G3: 64 -> 256 ([distribution | factor] hidden) ----
    if (tid == 0) {
      fence_after_thread_sync();
      issue_gemm(tmem_base, sbase + OFF_A0, sbase + OFF_W3, 256, 64, false);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_thread_sync();
#pragma unroll
    for (int cc = 0; cc < 128; cc += 32) {
      const int c0 = hf * 128 + cc;
      float v[32];
      tmem_ld32(tmem_row + c0, v);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(v[8 * g + j] + bias[192 + c0 + 8 * g + j], 0.f);
        uint4 pk = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                              pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        *reinterpret_cast<uint4*>(smem + OFF_A1 + canon_off(row, (c0 >> 3) + g, TM)) = pk;
      }
    }
    fence_proxy_async_smem();
    fence_before_thread_sync();
    __syncthreads();

    // ---- G4: 256 -> 16 (T logits | factor logit | zero padding) ----
    if (tid == 0) {
      fence_after_thread_sync();
      issue_gemm(tmem_base, sbase + OFF_A1, sbase + OFF_W4, 16, 256, false);
      mma_commit(mbar);
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_thread_sync();
    if (hf == 0) {
      float v[16];
      tmem_ld16(tmem_row, v);
      if (row < nrows) {
        const long long grow = row0 + row;
        float y[GN_SMALL_OUT - 1];
        float mx = -INFINITY;
#pragma unroll
        for (int t = 0; t < GN_SMALL_OUT - 1; ++t) {
          if (t < T) {
            float u;
            if (noise_mode == GN_NOISE_GIVEN) {
              u = __ldg(U + static_cast<size_t>(grow) * T + t);
            } else {
              unsigned long long el = (static_cast<unsigned long long>(scene_offset) * E + grow) * T + t;
              u = Philox::uniform(el, static_cast<uint32_t>(stage_index), seed);
            }
            y[t] = (v[t] + bias[448 + t] + gumbel_from_uniform(u)) / 0.5f;
            mx = fmaxf(mx, y[t]);
          }
        }
        float den = 0.f;
#pragma unroll
        for (int t = 0; t < GN_SMALL_OUT - 1; ++t)
          if (t < T) { y[t] = expf(y[t] - mx); den += y[t]; }
        float fl = 0.f;
#pragma unroll
        for (int o = 0; o < GN_SMALL_OUT; ++o)
          if (o == T) fl = v[o] + bias[448 + o];
        const float factor = 1.f / (1.f + expf(-fl));
#pragma unroll
        for (int t = 0; t < GN_SMALL_OUT - 1; ++t)
          if (t < T) {
            float d = y[t] / den;
            if (dist_out != nullptr) dist_out[static_cast<size_t>(grow) * T + t] = d;
            edge_feat[static_cast<size_t>(grow) * T + t] = factor * d;
          }
      }
    }
    // the next tile's G1 overwrites TMEM columns and A0: order it after this epilogue
    fence_before_thread_sync();
    __syncthreads();
  }

  fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    fence_after_thread_sync();
    tmem_dealloc(tmem_base, tcmlp::TMEM_COLS);
  }
}

int launch_edge_mlp_tc(const float* edges, long long R, int T, int E, const gn_stage_weights* w,
                       const float* U, int noise_mode, unsigned long long seed, long long scene_offset,
                       int stage_index, float* dist_out, float* edge_feat, cudaStream_t st) {
  if (!w->tc_init_w0 || !w->tc_init_w1 || !w->tc_df_w0 || !w->tc_df_w1) return GN_E_NULL;
  TcMlpWeights W{w->tc_init_w0, w->tc_init_w1, w->tc_df_w0, w->tc_df_w1,
                 w->init_b0, w->init_b1, w->df_b0, w->df_b1};
  cudaError_t e = cudaFuncSetAttribute(edge_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(tcmlp::SMEM_BYTES));
  if (e != cudaSuccess) return static_cast<int>(e);
  long long ntiles = (R + tcmlp::TM - 1) / tcmlp::TM;
  int grid = ntiles < GN_SM_COUNT ? static_cast<int>(ntiles) : GN_SM_COUNT;
  {
    ProfScope ps__("edge_mlp_tc", st);
    edge_mlp_tc_kernel<<<grid, GN_THREADS, tcmlp::SMEM_BYTES, st>>>(
        edges, R, T, E, W, U, noise_mode, seed, scene_offset, stage_index, dist_out, edge_feat);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
