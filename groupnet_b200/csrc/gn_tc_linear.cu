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
//   for each 64-wide K chunk:   stage A (fp32 row-major -> bf16 canonical through registers, or bf16 row-major
//                               by cp.async), stage W (bulk copies of the pre-arranged chunk, mbarrier complete_tx),
//                               one thread issues the chunk's tcgen05.mma's; two stages per group, so chunk g + 1
//                               loads while chunk g multiplies and the next tile's first chunk under the epilogue
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
constexpr int KCH = 64;                                // K chunk = one pipeline stage
constexpr int NST = 2;                                 // stages per group: chunk g + 1 loads while chunk g multiplies
constexpr uint32_t A_BYTES = 128 * KCH * 2;            // 16 KB
constexpr uint32_t W_BYTES = 256 * KCH * 2;            // 32 KB
constexpr uint32_t STG_BYTES = A_BYTES + W_BYTES;
constexpr uint32_t GRP_BYTES = NST * STG_BYTES;        // 96 KB
constexpr uint32_t OFF_BIAS = 2 * GRP_BYTES;           // bias[256] floats
constexpr uint32_t OFF_BMAT = OFF_BIAS + 256 * 4;      // bias_mat[16][256] floats
constexpr uint32_t OFF_BAR = OFF_BMAT + 16 * 256 * 4;  // MMA mbarriers [grp][stage] | weight-copy mbarriers [grp][stage] | tmem slot
constexpr uint32_t SMEM_BYTES = OFF_BAR + 80;
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
  unsigned char* sG = smem + grp * GRP_BYTES;
  float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* sbmat = reinterpret_cast<float*>(smem + OFF_BMAT);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + grp * NST;        // [stage]: its MMAs have completed
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR + 32) + grp * NST;   // [stage]: its weight chunk has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);
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
  if (gtid == 32) {
    for (int st = 0; st < NST; ++st) { mbar_init(mbar + st, 1); mbar_init(wbar + st, 1); }
  }
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_grp = *tmem_slot + grp * 256;
  const uint32_t tmem_row = tmem_grp + (static_cast<uint32_t>(q * 32) << 16);
  const long long ntiles = (a.R + 127) / 128;
  const int nch = (K + KCH - 1) / KCH;
  const long long tstep = static_cast<long long>(gridDim.x) * 2;
  const long long tile0 = static_cast<long long>(blockIdx.x) * 2 + grp;
  const bool from_bf16 = !a.a0_is_f32;
  const bool a_async = from_bf16 && a.K1 == 0 && a.a_div == 0.f;

  // The group's chunks form one sequence over its tiles (g = running index, stage = g & 1).  produce(g) loads chunk g
  // into its stage once the MMAs of chunk g - 2 have released it; consume(g) waits for the loads and issues the MMAs.
  // produce(g + 1) runs BEFORE consume(g): a chunk's loads are in flight under the previous chunk's wait and MMAs, and
  // the first chunk of the next tile under this tile's epilogue.
  auto produce = [&](long long tile, int c, uint32_t g) {
    const uint32_t st = g & 1u;
    if (g >= static_cast<uint32_t>(NST)) mbar_wait(mbar + st, ((g >> 1) - 1u) & 1u);
    unsigned char* sA = sG + st * STG_BYTES;
    const uint32_t sW_addr = smem_u32(sA + A_BYTES);
    const long long row0 = tile * 128;
    const int nrows = static_cast<int>(min(128LL, a.R - row0));
    const int kc0 = c * KCH, kcw = min(KCH, K - kc0), nk8 = kcw >> 3;
    // ---- W chunk: rows [n0, n0+N) of k-groups [kc0/8, kc0/8 + nk8): N*16 contiguous bytes per k-group, one bulk copy
    // each (issued by the group's first warp, completion counted in bytes on the stage's wbar)
    if (gtid < 32) {
      tcu::expect_tx(wbar + st, static_cast<uint32_t>(nk8) * N * 16);
      for (int k8 = 0; k8 < nk8; ++k8)
        tcu::bulk_g2s(sW_addr + static_cast<uint32_t>(k8) * N * 16,
                      a.W + (static_cast<size_t>((kc0 >> 3) + k8) * a.Ntot + a.n0) * 8, static_cast<uint32_t>(N) * 16, wbar + st);
    }
    // ---- A chunk: task = (row, k-group)
    const int r = gtid;
    if (a_async) {
      // a bf16 row-major source needs no conversion: its 16-byte k-groups go straight to their place in the canonical
      // operand by cp.async, all of a chunk's loads in flight at once
      if (r < nrows) {
        const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(a.A0) + static_cast<size_t>(row0 + r) * a.lda0 + kc0;
#pragma unroll 4
        for (int k8 = 0; k8 < nk8; ++k8) cp_async16(sA + canon_off(r, k8, 128), src + 8 * k8);
      } else {
        for (int k8 = 0; k8 < nk8; ++k8) *reinterpret_cast<uint4*>(sA + canon_off(r, k8, 128)) = make_uint4(0u, 0u, 0u, 0u);
      }
    } else {
#pragma unroll 4
      for (int k8 = 0; k8 < nk8; ++k8) {
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
    }
    cp_async_commit();                       // one group per chunk, empty on the conversion path
  };
  auto consume = [&](int c, uint32_t g, bool next_in_flight) {
    const uint32_t st = g & 1u;
    if (next_in_flight) cp_async_wait<1>(); else cp_async_wait<0>();
    fence_proxy_async_smem();
    fence_before_thread_sync();
    group_bar(grp);
    if (gtid < 32) {                         // warp-uniform issue: one elected lane, operands stay uniform
      mbar_wait(wbar + st, (g >> 1) & 1u);   // the weight chunk has landed
      fence_after_thread_sync();
      if (elect_one()) {
        const uint32_t sA_addr = smem_u32(sG + st * STG_BYTES);
        issue_gemm(tmem_grp, sA_addr, sA_addr + A_BYTES, N, min(KCH, K - c * KCH), c > 0);
        mma_commit(mbar + st);
      }
      __syncwarp();
    }
  };

  uint32_t g = 0;
  if (tile0 < ntiles) produce(tile0, 0, 0u);
  for (long long tile = tile0; tile < ntiles; tile += tstep) {
    const long long row0 = tile * 128;
    const int nrows = static_cast<int>(min(128LL, a.R - row0));
    for (int c = 0; c < nch; ++c, ++g) {
      long long nt = tile;
      int nc = c + 1;
      if (nc == nch) { nt = tile + tstep; nc = 0; }
      const bool more = nt < ntiles;
      if (more) produce(nt, nc, g + 1u);
      consume(c, g, more);
    }
    mbar_wait(mbar + ((g - 1u) & 1u), ((g - 1u) >> 1) & 1u);   // the tile's last chunk: the accumulator is complete
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
    // next tile's first MMA overwrites this group's TMEM columns
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
