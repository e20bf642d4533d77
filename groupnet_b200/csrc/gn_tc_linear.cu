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
constexpr uint32_t OFF_BAR = 2 * GRP_BYTES;            // 2 mbarriers + tmem slot
constexpr uint32_t SMEM_BYTES = OFF_BAR + 32;
}  // namespace tclin

__device__ __forceinline__ void group_bar(int grp) {
  asm volatile("bar.sync %0, 128;" :: "r"(grp + 1) : "memory");
}

__global__ void __launch_bounds__(GN_THREADS, 1)
tc_linear_kernel(TcLinArgs a) {
  using namespace tclin;
  using namespace tc;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127;
  const int q = (gtid >> 5), lane = tid & 31, row = q * 32 + lane;
  unsigned char* sA = smem + grp * GRP_BYTES;
  unsigned char* sW = sA + A_BYTES;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_BAR) + grp;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);

  if ((tid >> 5) == 0) tmem_alloc(tmem_slot, 512);
  if (gtid == 32) mbar_init(mbar, 1);
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_grp = *tmem_slot + grp * 256;
  const uint32_t tmem_row = tmem_grp + (static_cast<uint32_t>(q * 32) << 16);
  const uint32_t sA_addr = smem_u32(sA), sW_addr = smem_u32(sW);
  uint32_t phase = 0;
  const int K = a.K0 + a.K1, N = a.N;
  const long long ntiles = (a.R + 127) / 128;

  for (long long tile = static_cast<long long>(blockIdx.x) * 2 + grp; tile < ntiles;
       tile += static_cast<long long>(gridDim.x) * 2) {
    const long long row0 = tile * 128;
    const int nrows = static_cast<int>(min(128LL, a.R - row0));
    for (int kc0 = 0; kc0 < K; kc0 += KCH) {
      const int kcw = min(KCH, K - kc0), nk8 = kcw >> 3;
      if (kc0 > 0) { mbar_wait(mbar, phase); phase ^= 1; }   // previous chunk's MMAs released the buffers
      // ---- W chunk: rows [n0, n0+N) of k-groups [kc0/8, kc0/8 + nk8): N*16 contiguous bytes per k-group
      {
        const int per = N;                                   // 16-byte units per k-group
        for (int i = gtid; i < nk8 * per; i += 128) {
          int k8 = i / per, n = i - k8 * per;
          const __nv_bfloat16* src = a.W + (static_cast<size_t>((kc0 >> 3) + k8) * a.Ntot + a.n0 + n) * 8;
          cp_async16(sW + (static_cast<size_t>(k8) * N + n) * 16, src);
        }
        cp_async_commit();
      }
      // ---- A chunk: task = (row, k-group)
      for (int task = gtid; task < nk8 * 128; task += 128) {
        const int r = task & 127, k8 = task >> 7;
        const int k = kc0 + k8 * 8;
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (r < nrows) {
          if (k < a.K0 && !a.a0_is_f32) {
            pk = __ldg(reinterpret_cast<const uint4*>(
                static_cast<const __nv_bfloat16*>(a.A0) + static_cast<size_t>(row0 + r) * a.lda0 + k));
          } else {
            const float* src = (k < a.K0)
                ? static_cast<const float*>(a.A0) + static_cast<size_t>(row0 + r) * a.lda0 + k
                : a.A1 + static_cast<size_t>(row0 + r) * a.lda1 + (k - a.K0);
            float4 x = ldg_f4(src), y = ldg_f4(src + 4);
            if (a.a_div != 0.f) {
              x.x = __fdiv_rn(x.x, a.a_div); x.y = __fdiv_rn(x.y, a.a_div);
              x.z = __fdiv_rn(x.z, a.a_div); x.w = __fdiv_rn(x.w, a.a_div);
              y.x = __fdiv_rn(y.x, a.a_div); y.y = __fdiv_rn(y.y, a.a_div);
              y.z = __fdiv_rn(y.z, a.a_div); y.w = __fdiv_rn(y.w, a.a_div);
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
      if (gtid == 0) {
        fence_after_thread_sync();
        issue_gemm(tmem_grp, sA_addr, sW_addr, N, kcw, kc0 > 0);
        mma_commit(mbar);
      }
    }
    mbar_wait(mbar, phase); phase ^= 1;
    fence_after_thread_sync();

    // ---- epilogue: this thread owns tile row `row`
    const bool live = row < nrows;
    const long long grow = row0 + row;
    float rs[16];
    if (a.bias_mat != nullptr) {
#pragma unroll
      for (int t = 0; t < 16; ++t)
        rs[t] = (live && t < a.bm_T) ? __ldg(a.rowscale + static_cast<size_t>(grow) * a.rs_ld + t) : 0.f;
    }
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      if (N - c0 >= 32) {
        tmem_ld32(tmem_row + c0, v);
      } else {                                    // N % 32 == 16 tail
        float w[16];
        tmem_ld16(tmem_row + c0, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) { v[j] = w[j]; v[16 + j] = 0.f; }
      }
      const int ncol = min(32, N - c0);
      if (live) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < ncol) {
            const int gc = a.n0 + c0 + j;
            float x = v[j];
            if (a.bias != nullptr) x += __ldg(a.bias + gc);
            if (a.relu) x = fmaxf(x, 0.f);
            if (a.rowscale != nullptr && a.bias_mat == nullptr)
              x *= __ldg(a.rowscale + static_cast<size_t>(grow) * a.rs_ld + (gc >> a.rs_shift));
            if (a.bias_mat != nullptr) {
#pragma unroll
              for (int t = 0; t < 16; ++t)
                if (t < a.bm_T) x = fmaf(rs[t], __ldg(a.bias_mat + t * a.bm_ld + gc), x);
            }
            if (a.out_div != 0.f) x = __fdiv_rn(x, a.out_div);
            v[j] = x;
          }
        }
        if (a.out_is_f32) {
          float* dst = static_cast<float*>(a.out) + static_cast<size_t>(grow) * a.ldo + a.out_col0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (j < ncol) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(grow) * a.ldo + a.out_col0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            if (j < ncol)
              *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                              pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
        }
      }
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
// N % 16 == 0, 16 <= N <= 256, 16-byte aligned row starts.
int launch_tc_linear(const TcLinArgs& a, const char* name, cudaStream_t st) {
  const int K = a.K0 + a.K1;
  if (a.R <= 0) return GN_OK;
  if ((a.K0 & 7) || (K & 15) || (a.N & 15) || a.N < 16 || a.N > 256) return GN_E_SHAPE;
  if ((a.lda0 & 3) || (a.K1 && (a.lda1 & 3)) || (a.ldo & 3) || (a.out_col0 & 3)) return GN_E_ALIGN;
  if (!a.a0_is_f32 && (a.lda0 & 7)) return GN_E_ALIGN;
  if (!a.out_is_f32 && ((a.ldo & 7) || (a.out_col0 & 7))) return GN_E_ALIGN;
  cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(tclin::SMEM_BYTES));
  if (e != cudaSuccess) return static_cast<int>(e);
  long long ntiles = (a.R + 127) / 128;
  long long want = (ntiles + 1) / 2;
  int grid = want < GN_SM_COUNT ? static_cast<int>(want) : GN_SM_COUNT;
  {
    ProfScope ps__(name, st);
    tc_linear_kernel<<<grid, GN_THREADS, tclin::SMEM_BYTES, st>>>(a);
  }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

}  // namespace gn
