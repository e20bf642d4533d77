// Generic tcgen05 row-tile linear:  out[R x N] = epi( A[R x K] * W[n0:n0+N, :]^T )
//
// Used for every GEMM of the bf16 path that is not inside the fused per-edge
// chain (gn_edge_mlp_tc.cu): node2edge_start_mlp, the attention projections,
// the collapsed aggregation GEMMs and the closing MLP (MS_HGNN_batch.py:84,
// :80, :247-268, :77,:195).
//
// Structure: persistent CTAs of 256 threads = two independent 128-thread
// groups.  Each group owns its own tile stream (tile = 128 rows), its own
// shared-memory operand buffers, mbarrier and 256 TMEM columns, and runs
//   for each 128-wide K chunk:  stage A (fp32|bf16 row-major -> bf16 canonical),
//                               stage W (flat cp.async of the pre-arranged chunk),
//                               one thread issues the chunk's tcgen05.mma's
//   epilogue: tcgen05.ld -> bias / ReLU / per-row scale / rank-T bias -> global
// so one group's loads and epilogue overlap the other group's MMAs (the tensor
// core executes both groups' instructions in issue order).
//
// A may be the concatenation of two row-major sources (the [agg | h] / N input
// of the closing MLP, :267,:120) and may be divided by a scalar on the fly.
#include "gn_tc.cuh"
#include "gn_stage.h"

namespace gn {

namespace tclin {
constexpr int KCH = 128;                               // K chunk
constexpr uint32_t A_BYTES = 128 * KCH * 2;            // 32 KB
constexpr uint32_t W_BYTES = 256 * KCH * 2;            // 64 KB
constexpr uint32_t GRP_BYTES = A_BYTES + W_BYTES;
constexpr uint32_t OFF_BIAS = 2 * GRP_BYTES;           // bias[256] floats
constexpr uint32_t OFF_BMAT = OFF_BIAS + 256 * 4;      // bias_mat[16][256] floats
constexpr uint32_t OFF_BAR = OFF_BMAT + 16 * 256 * 4;  // 2 MMA mbarriers | tmem slot | 2 weight-copy mbarriers
constexpr uint32_t SMEM_BYTES = OFF_BAR + 48;
enum { MODE_PLAIN = 0, MODE_ROWSCALE = 1, MODE_BIASMAT = 2 };
}  // namespace tclin

__device__ __forceinline__ void group_bar(int grp) {
  asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory");
}

// Epilogue of one 32-column chunk held in v[] (compile-time modes, everything unrolled).
template <bool RELU, bool OUTF32, int MODE, int NC>
__device__ __forceinline__ void epilogue_chunk(const TcLinArgs& a, float (&v)[32], int c0, long long grow,
                                               const float* sbias, const float* sbmat, const float (&rs)[16],
                                               float scale) {
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float x = v[j] + sbias[c0 + j];
    if (RELU) x = fmaxf(x, 0.f);
    if (MODE == tclin::MODE_ROWSCALE) x *= scale;
    if (MODE == tclin::MODE_BIASMAT) {
#pragma unroll
      for (int t = 0; t < 16; ++t) x = fmaf(rs[t], sbmat[t * 256 + c0 + j], x);
    }
    v[j] = x;
  }
  if (OUTF32) {
    float* dst = static_cast<float*>(a.out) + static_cast<size_t>(grow) * a.ldo + a.out_col0 + c0;
#pragma unroll
    for (int j = 0; j < NC; j += 4)
      *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  } else {
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(grow) * a.ldo + a.out_col0 + c0;
#pragma unroll
    for (int j = 0; j < NC; j += 8)
      *reinterpret_cast<uint4*>(dst + j) = make_uint4(tc::pack_bf16(v[j], v[j + 1]), tc::pack_bf16(v[j + 2], v[j + 3]),
                                                      tc::pack_bf16(v[j + 4], v[j + 5]), tc::pack_bf16(v[j + 6], v[j + 7]));
  }
}

template <bool RELU, bool OUTF32, int MODE>
__global__ void __launch_bounds__(GN_THREADS, 1)
tc_linear_kernel(TcLinArgs a) {
  using namespace tclin;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127;
  const int q = (gtid >> 5), lane = tid & 31, row = q * 32 + lane;
  unsigned char* sA = smem + grp * GRP_BYTES;
  unsigned char* sW = sA + A_BYTES;
  float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* sbmat = reinterpret_cast<float*>(smem + OFF_BMAT);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + grp;
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR + 24) + grp;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
  const int K = a.K0 + a.K1, N = a.N;

  for (int i = tid; i < 256; i += GN_THREADS)
    sbias[i] = (a.bias != nullptr && i < N) ? __ldg(a.bias + a.n0 + i) : 0.f;
  if (MODE == MODE_BIASMAT) {
    for (int i = tid; i < 16 * 256; i += GN_THREADS) {
      int t = i >> 8, c = i & 255;
      sbmat[i] = (t < a.bm_T && c < N) ? __ldg(a.bias_mat + t * a.bm_ld + a.n0 + c) : 0.f;
    }
  }
  if ((tid >> 5) == 0) tmem_alloc(tmem_slot, 512);
  if (gtid == 32) { mbar_init(mbar, 1); mbar_init(wbar, 1); }
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_grp = *tmem_slot + grp * 256;
  const uint32_t tmem_row = tmem_grp + (static_cast<uint32_t>(q * 32) << 16);
  const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW);
  uint32_t phase = 0, wphase = 0;
  const long long ntiles = (a.R + 127) / 128;

  for (long long tile = static_cast<long long>(blockIdx.x) * 2 + grp; tile < ntiles;
       tile += static_cast<long long>(gridDim.x) * 2) {
    const long long row0 = tile * 128;
    const int nrows = static_cast<int>(min(128LL, a.R - row0));
    for (int kc0 = 0; kc0 < K; kc0 += KCH) {
      const int kcw = min(KCH, K - kc0), nk8 = kcw >> 3;
      if (kc0 > 0) { mbar_wait(mbar, phase); phase ^= 1; }   // previous chunk's MMAs released the buffers
      // ---- W chunk: rows [n0, n0+N) of k-groups [kc0/8, kc0/8 + nk8): N*16 contiguous bytes per k-group, one bulk copy
      // each (issued by the group's first warp, completion counted in bytes on wbar; 32 cp.async + index arithmetic per
      // thread before)
      if (gtid < 32) {
        tcu::expect_tx(wbar, static_cast<uint32_t>(nk8) * N * 16);
        for (int k8 = 0; k8 < nk8; ++k8)
          tcu::bulk_g2s(sW_addr + static_cast<uint32_t>(k8) * N * 16,
                        a.W + (static_cast<size_t>((kc0 >> 3) + k8) * a.Ntot + a.n0) * 8, static_cast<uint32_t>(N) * 16, wbar);
      }
      // ---- A chunk: task = (row, k-group)
      const bool from_bf16 = !a.a0_is_f32;
      if (from_bf16 && a.K1 == 0 && a.a_div == 0.f) {
        // a bf16 row-major source needs no conversion: its 16-byte k-groups go straight to their place in the canonical
        // operand by cp.async, all of a chunk's loads in flight at once (through registers the loop below exposes one
        // DRAM round trip per batch of four k-groups: 4 per chunk, ~10 us of a 13 us chunk on the decoder MLPs)
        const int r = gtid;
        if (r < nrows) {
          const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(a.A0) + static_cast<size_t>(row0 + r) * a.lda0 + kc0;
#pragma unroll 4
          for (int k8 = 0; k8 < nk8; ++k8) cp_async16(sA + canon_off(r, k8, 128), src + 8 * k8);
        } else {
          for (int k8 = 0; k8 < nk8; ++k8) *reinterpret_cast<uint4*>(sA + canon_off(r, k8, 128)) = make_uint4(0u, 0u, 0u, 0u);
        }
        cp_async_commit();
      } else
#pragma unroll 4
      for (int k8 = 0; k8 < nk8; ++k8) {
        const int r = gtid;
        const int k = kc0 + k8 * 8;
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (r < nrows) {
          if (k < a.K0 && from_bf16) {
            pk = __ldg(reinterpret_cast<const uint4*>(
                static_cast<const __nv_bfloat16*>(a.A0) + static_cast<size_t>(row0 + r) * a.lda0 + k));
          } else {
            const float* src = (k < a.K0)
                ? static_cast<const float*>(a.A0) + static_cast<size_t>(row0 + r) * a.lda0 + k
                : a.A1 + static_cast<size_t>(row0 + r) * a.lda1 + (k - a.K0);
            float4 x = ldg_f4(src), y = ldg_f4(src + 4);
            if (a.a_div != 0.f) {
              const float inv = a.a_div;
              x.x = __fdividef(x.x, inv); x.y = __fdividef(x.y, inv); x.z = __fdividef(x.z, inv); x.w = __fdividef(x.w, inv);
              y.x = __fdividef(y.x, inv); y.y = __fdividef(y.y, inv); y.z = __fdividef(y.z, inv); y.w = __fdividef(y.w, inv);
            }
            pk = make_uint4(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w), pack_bf16(y.x, y.y), pack_bf16(y.z, y.w));
          }
        }
        *reinterpret_cast<uint4*>(sA + canon_off(r, k8, 128)) = pk;
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      fence_before_thread_sync();
      group_bar(grp);
      if (gtid < 32) {                       // warp-uniform issue: one elected lane, operands stay uniform
        mbar_wait(wbar, wphase);             // the weight chunk has landed
        fence_after_thread_sync();
        if (elect_one()) {
          issue_gemm(tmem_grp, sA_addr, sW_addr, N, kcw, kc0 > 0);
          mma_commit(mbar);
        }
        __syncwarp();
      }
      wphase ^= 1;
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_thread_sync();

    // ---- epilogue: this thread owns tile row `row`
    const bool live = row < nrows;
    const long long grow = row0 + row;
    float rs[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) rs[t] = 0.f;
    if (MODE == MODE_BIASMAT && live) {
#pragma unroll
      for (int t = 0; t < 16; ++t)
        if (t < a.bm_T) rs[t] = __ldg(a.rowscale + static_cast<size_t>(grow) * a.rs_ld + t);
    }
    const int nfull = N & ~31;
    for (int c0 = 0; c0 < nfull; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_row + c0, v);
      float scale = 1.f;
      if (MODE == MODE_ROWSCALE && live)
        scale = __ldg(a.rowscale + static_cast<size_t>(grow) * a.rs_ld + ((a.n0 + c0) >> a.rs_shift));
      if (live) epilogue_chunk<RELU, OUTF32, MODE, 32>(a, v, c0, grow, sbias, sbmat, rs, scale);
    }
    if (N & 16) {                                    // 16-column tail
      float w[16], v[32];
      tmem_ld16(tmem_row + nfull, w);
#pragma unroll
      for (int j = 0; j < 16; ++j) { v[j] = w[j]; v[16 + j] = 0.f; }
      float scale = 1.f;
      if (MODE == MODE_ROWSCALE && live)
        scale = __ldg(a.rowscale + static_cast<size_t>(grow) * a.rs_ld + ((a.n0 + nfull) >> a.rs_shift));
      if (live) epilogue_chunk<RELU, OUTF32, MODE, 16>(a, v, nfull, grow, sbias, sbmat, rs, scale);
    }
    // next tile's first MMA overwrites this group's TMEM columns and operand buffers
    fence_before_thread_sync();
    group_bar(grp);
  }

  fence_before_thread_sync();
  __syncthreads();
  if ((tid >> 5) == 0) {
    fence_after_thread_sync();
    tmem_dealloc(*tmem_slot, 512);
  }
}

// host-side launcher; returns GN_OK / error.  Requirements: K0 % 8 == 0, K % 16 == 0,
// N % 16 == 0, 16 <= N <= 256, 16-byte aligned row starts; a per-row scale must be
// constant over each 32-column chunk (rs_shift >= 5).
template <bool RELU, bool OUTF32, int MODE>
static int launch_one(const TcLinArgs& a, int grid, const char* name, cudaStream_t st) {
  auto kern = tc_linear_kernel<RELU, OUTF32, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(tclin::SMEM_BYTES));
  if (e != cudaSuccess) return static_cast<int>(e);
  {
    ProfScope ps__(name, st);
    kern<<<grid, GN_THREADS, tclin::SMEM_BYTES, st>>>(a);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

int launch_tc_linear(const TcLinArgs& a, const char* name, cudaStream_t st) {
  const int K = a.K0 + a.K1;
  if (a.R <= 0) return GN_OK;
  if ((a.K0 & 7) || (K & 15) || (a.N & 15) || a.N < 16 || a.N > 256) return GN_E_SHAPE;
  if ((a.lda0 & 3) || (a.K1 && (a.lda1 & 3)) || (a.ldo & 3) || (a.out_col0 & 3)) return GN_E_ALIGN;
  if (!a.a0_is_f32 && (a.lda0 & 7)) return GN_E_ALIGN;
  if (!a.out_is_f32 && ((a.ldo & 7) || (a.out_col0 & 7))) return GN_E_ALIGN;
  const int mode = a.bias_mat ? tclin::MODE_BIASMAT : (a.rowscale ? tclin::MODE_ROWSCALE : tclin::MODE_PLAIN);
  if (mode == tclin::MODE_ROWSCALE && a.rs_shift < 5) return GN_E_SHAPE;
  if (mode == tclin::MODE_BIASMAT && (a.bm_T > 16 || !a.rowscale)) return GN_E_SHAPE;
  if (a.out_div != 0.f) return GN_E_SHAPE;            // reserved
  long long ntiles = (a.R + 127) / 128;
  long long want = (ntiles + 1) / 2;
  int grid = want < GN_SM_COUNT ? static_cast<int>(want) : GN_SM_COUNT;
  const int key = (a.relu ? 1 : 0) | (a.out_is_f32 ? 2 : 0) | (mode << 2);
  switch (key) {
    case 0:  return launch_one<false, false, 0>(a, grid, name, st);
    case 1:  return launch_one<true, false, 0>(a, grid, name, st);
    case 2:  return launch_one<false, true, 0>(a, grid, name, st);
    case 3:  return launch_one<true, true, 0>(a, grid, name, st);
    case 5:  return launch_one<true, false, 1>(a, grid, name, st);
    case 10: return launch_one<false, true, 2>(a, grid, name, st);
    default: return GN_E_SHAPE;
  }
}

}  // namespace gn
