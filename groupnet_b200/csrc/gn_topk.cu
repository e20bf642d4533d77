// Kernel (a): fused agent feature-correlation + per-agent top-k hyperedge
// selection for every scale + incidence emission.
//
// Replaces, for S scales in ONE pass over x:
//   q = F.normalize(x, p=2, dim=2); corr = q @ q^T     model/GroupNet_nba.py:284-286
//   idx = topk(corr, k); H = zeros.scatter(2, idx, 1)  model/MS_HGNN_batch.py:382-385
//   H = ones(B,1,N) when scale == N                     model/MS_HGNN_batch.py:375-377
//
// Layout: a CTA owns a group of SG whole scenes at a time (grid-stride over
// groups).  x rows of the group are contiguous in HBM and are staged once in
// shared memory with 128-bit loads; the Gram matrix, the sort and every H_s are
// produced from shared memory; each H_s block of the group is contiguous in
// HBM and is written with coalesced stores.  corr never touches HBM unless
// the caller asks for it.
//
// Sort: one (sub-)warp per correlation row, a bitonic network over P lanes
// (P = N rounded up to a power of two; 2 keys per lane when P == 64) on
// (value desc, agent index asc) keys exchanged with warp shuffles.  One sort
// serves every scale: rank < k_s <=> member of the scale-s hyperedge.
#include "gn_common.cuh"

namespace gn {

struct TopkArgs {
  int S;
  int k[GN_MAX_SCALES];        // clamped to >= 1; ignored when ones[s]
  int ones[GN_MAX_SCALES];     // scale == N: single all-ones hyperedge
  float* H[GN_MAX_SCALES];
  long long stride[GN_MAX_SCALES];
};

__device__ __forceinline__ bool key_before(float va, int ia, float vb, int ib) {
  return (va > vb) || (va == vb && ia < ib);
}

// Bitonic sort of P keys held one per lane in P-lane segments of a warp.
template <int P>
__device__ __forceinline__ void bitonic_sort_lanes(float& v, int& idx, int sub) {
#pragma unroll
  for (int k = 2; k <= P; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      float pv = __shfl_xor_sync(0xffffffffu, v, j);
      int pi = __shfl_xor_sync(0xffffffffu, idx, j);
      bool up = (sub & k) == 0;
      bool lower = (sub & j) == 0;
      bool mine_first = key_before(v, idx, pv, pi);
      bool keep = (lower == up) ? mine_first : !mine_first;
      if (!keep) { v = pv; idx = pi; }
    }
  }
}

// 64 keys per row: position p = lane + 32*r, r in {0,1}.
__device__ __forceinline__ void bitonic_sort_64(float (&v)[2], int (&idx)[2], int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {
        // partner is the other register of the same lane; k == 64 here: ascending order
        bool first0 = key_before(v[0], idx[0], v[1], idx[1]);
        if (!first0) {
          float tv = v[0]; v[0] = v[1]; v[1] = tv;
          int ti = idx[0]; idx[0] = idx[1]; idx[1] = ti;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          int pos = lane + 32 * r;
          float pv = __shfl_xor_sync(0xffffffffu, v[r], j);
          int pi = __shfl_xor_sync(0xffffffffu, idx[r], j);
          bool up = (pos & k) == 0;
          bool lower = (pos & j) == 0;
          bool mine_first = key_before(v[r], idx[r], pv, pi);
          bool keep = (lower == up) ? mine_first : !mine_first;
          if (!keep) { v[r] = pv; idx[r] = pi; }
        }
      }
    }
  }
}

// Dynamic shared memory plan (floats unless noted):
//   xs   [SG*N][D+4]      staged features (absent when FROM_CORR)
//   inv  [SG*N]           1 / max(||x||, 1e-12)
//   cs   [SG*N][N+1]      correlation rows
//   rk   [SG*N][NP4] u8   rank of agent n in row r (NP4 = N rounded up to 4)
template <int P, bool FROM_CORR>
__global__ void __launch_bounds__(GN_THREADS)
corr_topk_kernel(const float* __restrict__ x, const float* __restrict__ corr_in,
                 int B, int N, int D, int SG, TopkArgs a, float* __restrict__ corr_out) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = GN_THREADS / 32;
  const int ldx = D + 4, ldc = N + 1, np4 = (N + 3) & ~3;
  const int rows_max = SG * N;
  float* xs = smem;
  float* inv = xs + (FROM_CORR ? 0 : rows_max * ldx);
  float* cs = inv + ((rows_max + 3) & ~3);
  unsigned char* rk = reinterpret_cast<unsigned char*>(cs + ((rows_max * ldc + 3) & ~3));

  const int ngroups = (B + SG - 1) / SG;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int b0 = grp * SG;
    const int ns = min(SG, B - b0);
    const int rows = ns * N;

    if (!FROM_CORR) {
      // ---- 1. stage x rows (contiguous block of rows*D floats), 128-bit loads
      const float* src = x + static_cast<size_t>(b0) * N * D;
      const int d4 = D >> 2;
      for (int i = tid; i < rows * d4; i += GN_THREADS) {
        int r = i / d4, c = i - r * d4;
        float4 v = ldg_stream_f4(src + static_cast<size_t>(r) * D + 4 * c);
        *reinterpret_cast<float4*>(xs + r * ldx + 4 * c) = v;
      }
      __syncthreads();
      // ---- 2. row norms, one warp per row; normalise in place with a true
      //         division like F.normalize (x / max(||x||, eps))
      for (int r = warp; r < rows; r += NWARP) {
        float ss = 0.f;
        for (int c = lane; c < d4; c += 32) {
          float4 v = *reinterpret_cast<const float4*>(xs + r * ldx + 4 * c);
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        float den = fmaxf(sqrtf(ss), 1e-12f);
        for (int c = lane; c < d4; c += 32) {
          float4 v = *reinterpret_cast<float4*>(xs + r * ldx + 4 * c);
          v.x = __fdiv_rn(v.x, den); v.y = __fdiv_rn(v.y, den);
          v.z = __fdiv_rn(v.z, den); v.w = __fdiv_rn(v.w, den);
          *reinterpret_cast<float4*>(xs + r * ldx + 4 * c) = v;
        }
      }
      __syncthreads();
      // ---- 3. Gram matrix: 4x4 register blocks over the upper block triangle,
      //         K split over 2 adjacent lanes when D % 8 == 0
      const int nb = (N + 3) >> 2;
      const int ntri = nb * (nb + 1) / 2;
      // K split over ks adjacent lanes: pick the split that minimises passes x slice length
      // (N = 64: 136 blocks -> ks = 8 gives 5 passes of 8 k-steps instead of 2 passes of 32)
      int ks = 1;
      {
        int best = 1 << 30;
        for (int cand = 1; cand <= 8; cand <<= 1) {
          if (D % (4 * cand)) break;
          const int cost = ((ns * ntri * cand + GN_THREADS - 1) / GN_THREADS) * (D / (4 * cand));
          if (cost < best) { best = cost; ks = cand; }
        }
      }
      const int ntask = ns * ntri * ks;
      for (int base = 0; base < ntask; base += GN_THREADS) {
        int task = base + tid;
        bool valid = task < ntask;
        int tt = valid ? task : 0;
        int half = tt % ks; tt /= ks;
        int tri = tt % ntri; int g = tt / ntri;
        // decode tri -> (bi <= bj)
        int bi = 0, rem = tri;
        while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
        int bj = bi + rem;
        const float* qa[4]; const float* qb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          int ra = min(4 * bi + u, N - 1), rb = min(4 * bj + u, N - 1);
          qa[u] = xs + (g * N + ra) * ldx;
          qb[u] = xs + (g * N + rb) * ldx;
        }
        float acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int w = 0; w < 4; ++w) acc[u][w] = 0.f;
        // interleaved K slices: the ks lanes of a block read consecutive float4s (no bank conflicts)
        for (int k = 4 * half; k < D; k += 4 * ks) {
          float4 av[4], bv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            av[u] = *reinterpret_cast<const float4*>(qa[u] + k);
            bv[u] = *reinterpret_cast<const float4*>(qb[u] + k);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              acc[u][w] = fmaf(av[u].x, bv[w].x, acc[u][w]);
              acc[u][w] = fmaf(av[u].y, bv[w].y, acc[u][w]);
              acc[u][w] = fmaf(av[u].z, bv[w].z, acc[u][w]);
              acc[u][w] = fmaf(av[u].w, bv[w].w, acc[u][w]);
            }
        }
        for (int o = 1; o < ks; o <<= 1) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w)
              acc[u][w] += __shfl_xor_sync(0xffffffffu, acc[u][w], o);
        }
        if (valid && half == 0) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              int i = 4 * bi + u, j = 4 * bj + w;
              if (i < N && j < N) {
                cs[(g * N + i) * ldc + j] = acc[u][w];
                if (bi != bj) cs[(g * N + j) * ldc + i] = acc[u][w];
              }
            }
        }
      }
      __syncthreads();
    } else {
      // corr rows supplied by the caller: (ns*N*N) contiguous floats
      const float* src = corr_in + static_cast<size_t>(b0) * N * N;
      for (int i = tid; i < rows * N; i += GN_THREADS) {
        int r = i / N, c = i - r * N;
        cs[r * ldc + c] = __ldg(src + i);
      }
      __syncthreads();
    }

    // ---- optional corr output (coalesced, contiguous block)
    if (!FROM_CORR && corr_out != nullptr) {
      float* dst = corr_out + static_cast<size_t>(b0) * N * N;
      for (int i = tid; i < rows * N; i += GN_THREADS) {
        int r = i / N, c = i - r * N;
        dst[i] = cs[r * ldc + c];
      }
    }

    // ---- 4. sort every row once; rank[r][agent] = position in the descending order
    if (P <= 32) {
      constexpr int PL = (P <= 32) ? P : 32;
      constexpr int RPW = 32 / PL;                 // rows per warp pass
      const int sub = lane % PL, seg = lane / PL;
      for (int rb = warp * RPW; rb < rows; rb += NWARP * RPW) {
        int r = rb + seg;
        bool rv = r < rows;
        float v = -INFINITY; int idx = 0x7fffffff;
        if (rv && sub < N) { v = cs[r * ldc + sub]; idx = sub; }
        bitonic_sort_lanes<PL>(v, idx, sub);
        if (rv && idx < N) rk[r * np4 + idx] = static_cast<unsigned char>(sub);
      }
    } else {
      for (int r = warp; r < rows; r += NWARP) {
        float v[2]; int idx[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          int p = lane + 32 * q;
          if (p < N) { v[q] = cs[r * ldc + p]; idx[q] = p; }
          else { v[q] = -INFINITY; idx[q] = 0x7fffffff; }
        }
        bitonic_sort_64(v, idx, lane);
#pragma unroll
        for (int q = 0; q < 2; ++q)
          if (idx[q] < N) rk[r * np4 + idx[q]] = static_cast<unsigned char>(lane + 32 * q);
      }
    }
    __syncthreads();

    // ---- 5. emit every H_s.  Within a scene the (E,N) block is contiguous.
    for (int s = 0; s < a.S; ++s) {
      float* Hs = a.H[s];
      const long long st = a.stride[s];
      if (a.ones[s]) {
        for (int i = tid; i < ns * N; i += GN_THREADS) {
          int g = i / N, n = i - g * N;
          Hs[static_cast<long long>(b0 + g) * st + n] = 1.0f;
        }
      } else {
        // one warp per correlation row, lanes over the agents: consecutive rows are contiguous in H, so the
        // stores stay coalesced and the loop has no integer divisions (they were 25 % of this kernel's samples)
        const int k = a.k[s];
        if (N >= 32) {
          for (int r = warp; r < rows; r += NWARP) {
            const int g = r / N, e = r - g * N;
            const unsigned char* rr = rk + r * np4;
            float* dst = Hs + static_cast<long long>(b0 + g) * st + static_cast<long long>(e) * N;
            for (int n = lane; n < N; n += 32) dst[n] = rr[n] < k ? 1.0f : 0.0f;
          }
        } else {                                         // short rows: a flat index keeps every lane busy
          const int per = N * N;
          for (int i = tid; i < ns * per; i += GN_THREADS) {
            int g = i / per, o = i - g * per;
            int e = o / N, n = o - e * N;
            Hs[static_cast<long long>(b0 + g) * st + o] = rk[(g * N + e) * np4 + n] < k ? 1.0f : 0.0f;
          }
        }
      }
    }
    __syncthreads();   // smem is reused by the next group
  }
}

// ---------------------------------------------------------------------------
// Small-N fast path (N <= 16: NBA N=11, fish N=8).  At the HBM roofline an SM has ~145 cycles per
// NBA scene, i.e. ~580 warp-instructions: a shuffle network per 11-element row does not fit.  Here
//   * one THREAD per feature row loads it (16 independent 128-bit loads), stages it raw in shared
//     memory and keeps 1 / max(||x||, 1e-12);
//   * the Gram matrix is computed on the raw rows in 4x4 register blocks over the upper block
//     triangle and scaled by inv_i * inv_j (differs from normalise-then-dot by ~1e-7);
//   * one THREAD per correlation row holds the row in registers and runs the all-pairs comparator
//     network (NP(NP-1)/2 compare-and-count steps, (value desc, index asc) order): rank[n];
//   * the same thread emits its row of every H_s (rank[n] < k_s): consecutive threads write
//     consecutive rows, so each H_s block of the group is written contiguously.
// ---------------------------------------------------------------------------
template <int NP, bool FROM_CORR>
__global__ void __launch_bounds__(GN_THREADS, 2)
corr_topk_small_kernel(const float* __restrict__ x, const float* __restrict__ corr_in,
                       int B, int N, int D, int SG, TopkArgs a, float* __restrict__ corr_out) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int ldx = D + 4, ldc = N | 1;                    // odd row stride: conflict-free row-per-thread reads
  const int rows_max = SG * N;
  // x rows are staged with ONE TMA bulk copy per row (cp.async.bulk -> padded shared-memory row, mbarrier
  // complete_tx).  The per-16-byte cp.async loop this replaces cost a quarter of the kernel's issue slots
  // (round-1 ncu source view); double-buffering the groups was tried and lost (smaller groups, 0.17 -> 0.19 ms).
  float* xs0 = smem;
  float* inv = xs0 + (FROM_CORR ? 0 : rows_max * ldx);
  float* cs = inv + ((rows_max + 3) & ~3);
  __shared__ __align__(8) unsigned long long xbar[2];
  const int ngroups = (B + SG - 1) / SG;
  auto issue_rows = [&](int grp, int buf) {
    const int b0 = grp * SG;
    const int rows = min(SG, B - b0) * N;
    const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&xbar[buf]));
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(rows * D * 4) : "memory");
    for (int r = tid; r < rows; r += GN_THREADS) {
      const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(xs0 + (buf * rows_max + r) * ldx));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   :: "r"(dst), "l"(x + (static_cast<size_t>(b0) * N + r) * D), "r"(D * 4), "r"(bar) : "memory");
    }
  };
  unsigned xphase = 0u;                                  // bit b = parity to wait for on xbar[b]
  if (!FROM_CORR) {
    if (tid == 0) {
      for (int b = 0; b < 2; ++b)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(static_cast<unsigned>(__cvta_generic_to_shared(&xbar[b]))) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  const int buf = 0;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int b0 = grp * SG;
    const int ns = min(SG, B - b0);
    const int rows = ns * N;
    float* xs = xs0 + buf * rows_max * ldx;
    if (!FROM_CORR) {
      __syncthreads();                                   // previous group's readers (cs, xs) are done
      const int d4 = D >> 2;
      // ---- 1. stage this group's rows (the other resident CTA of the SM overlaps its compute with this wait)
      issue_rows(grp, buf);
      {
        const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&xbar[buf]));
        const unsigned par = (xphase >> buf) & 1u;
        unsigned done = 0;
        for (unsigned spin = 0; spin < (1u << 24) && !done; ++spin)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(done) : "r"(bar), "r"(par) : "memory");
        if (!done) __trap();
        xphase ^= 1u << buf;
      }
      if (NP <= 12 && (D & 7) == 0) {
        // ---- 2a. raw Gram in k-slices: 8 adjacent lanes per scene, lane t owns float2 columns
        //          {t, t+8, ...}; every x element is read from shared memory exactly once.  The
        //          squared norms are the Gram diagonal.  Partial sums are reduced over the 8 lanes.
        constexpr int NPAIR = NP * (NP + 1) / 2;
        const int d2 = D >> 1;
        const int ntask = ns * 8;
        for (int base = 0; base < ntask; base += GN_THREADS) {
          const int task = base + tid;
          const bool valid = task < ntask;
          const int g = (valid ? task : 0) >> 3, t = task & 7;
          const float* xg = xs + g * N * ldx;
          float acc[NPAIR];
#pragma unroll
          for (int p = 0; p < NPAIR; ++p) acc[p] = 0.f;
#pragma unroll 2
          for (int c = t; c < d2; c += 8) {
            float2 v[NP];
#pragma unroll
            for (int r = 0; r < NP; ++r)
              v[r] = (r < N) ? *reinterpret_cast<const float2*>(xg + r * ldx + 2 * c) : make_float2(0.f, 0.f);
            int p = 0;
#pragma unroll
            for (int i = 0; i < NP; ++i)
#pragma unroll
              for (int j = i; j < NP; ++j, ++p) {
                acc[p] = fmaf(v[i].x, v[j].x, acc[p]);
                acc[p] = fmaf(v[i].y, v[j].y, acc[p]);
              }
          }
#pragma unroll
          for (int p = 0; p < NPAIR; ++p) {
            acc[p] += __shfl_xor_sync(0xffffffffu, acc[p], 1);
            acc[p] += __shfl_xor_sync(0xffffffffu, acc[p], 2);
            acc[p] += __shfl_xor_sync(0xffffffffu, acc[p], 4);
          }
          if (valid) {
            // every lane holds all sums; lane t writes the pairs with p % 8 == t (mirrored)
            float invn[NP];
            int p = 0;
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              invn[i] = 1.0f / fmaxf(sqrtf(acc[p]), 1e-12f);
              p += NP - i;
            }
            p = 0;
#pragma unroll
            for (int i = 0; i < NP; ++i)
#pragma unroll
              for (int j = i; j < NP; ++j, ++p) {
                if ((p & 7) == t && i < N && j < N) {
                  const float c = acc[p] * (invn[i] * invn[j]);
                  cs[(g * N + i) * ldc + j] = c;
                  cs[(g * N + j) * ldc + i] = c;
                }
              }
          }
        }
      } else {
        // ---- 2b. generic: row norms (one thread per row from smem), then 4x4 register blocks
        for (int r = tid; r < rows; r += GN_THREADS) {
          const float* src = xs + r * ldx;
          float ss = 0.f;
          for (int c = 0; c < d4; ++c) {
            float4 v = *reinterpret_cast<const float4*>(src + 4 * c);
            ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
          }
          inv[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        }
        __syncthreads();
        const int nb = (N + 3) >> 2;
        const int ntri = nb * (nb + 1) / 2;
        const int ks = (D % 8 == 0) ? 2 : 1;
        const int ntask = ns * ntri * ks;
        for (int base = 0; base < ntask; base += GN_THREADS) {
          int task = base + tid;
          bool valid = task < ntask;
          int tt = valid ? task : 0;
          int half = tt % ks; tt /= ks;
          int tri = tt % ntri; int g = tt / ntri;
          int bi = 0, rem = tri;
          while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
          int bj = bi + rem;
          const float* qa[4]; const float* qb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            qa[u] = xs + (g * N + min(4 * bi + u, N - 1)) * ldx;
            qb[u] = xs + (g * N + min(4 * bj + u, N - 1)) * ldx;
          }
          float acc[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w) acc[u][w] = 0.f;
          const int kbeg = half * (D / ks), kend = kbeg + D / ks;
          for (int k = kbeg; k < kend; k += 4) {
            float4 av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              av[u] = *reinterpret_cast<const float4*>(qa[u] + k);
              bv[u] = *reinterpret_cast<const float4*>(qb[u] + k);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                acc[u][w] = fmaf(av[u].x, bv[w].x, acc[u][w]);
                acc[u][w] = fmaf(av[u].y, bv[w].y, acc[u][w]);
                acc[u][w] = fmaf(av[u].z, bv[w].z, acc[u][w]);
                acc[u][w] = fmaf(av[u].w, bv[w].w, acc[u][w]);
              }
          }
          if (ks == 2) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int w = 0; w < 4; ++w) acc[u][w] += __shfl_xor_sync(0xffffffffu, acc[u][w], 1);
          }
          if (valid && half == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                int i = 4 * bi + u, j = 4 * bj + w;
                if (i < N && j < N) {
                  float c = acc[u][w] * (inv[g * N + i] * inv[g * N + j]);   // commutative: exact symmetry
                  cs[(g * N + i) * ldc + j] = c;
                  if (bi != bj) cs[(g * N + j) * ldc + i] = c;
                }
              }
          }
        }
      }
      __syncthreads();
      if (corr_out != nullptr) {
        float* dst = corr_out + static_cast<size_t>(b0) * N * N;
        for (int i = tid; i < rows * N; i += GN_THREADS) {
          int r = i / N, c = i - r * N;
          dst[i] = cs[r * ldc + c];
        }
      }
    }
    // ---- rank + emit: one thread per correlation row
    for (int r = tid; r < rows; r += GN_THREADS) {
      float v[NP];
      int rank[NP];
      if (FROM_CORR) {
        const float* src = corr_in + (static_cast<size_t>(b0) * N + r) * N;
#pragma unroll
        for (int n = 0; n < NP; ++n) v[n] = (n < N) ? __ldg(src + n) : -INFINITY;
      } else {
#pragma unroll
        for (int n = 0; n < NP; ++n) v[n] = (n < N) ? cs[r * ldc + n] : -INFINITY;
      }
#pragma unroll
      for (int n = 0; n < NP; ++n) rank[n] = 0;
#pragma unroll
      for (int j = 1; j < NP; ++j)
#pragma unroll
        for (int m = 0; m < j; ++m) {
          // m < j: m precedes j iff v[m] >= v[j] (ties to the lower index)
          bool mfirst = v[m] >= v[j];
          rank[j] += mfirst ? 1 : 0;
          rank[m] += mfirst ? 0 : 1;
        }
      const int sc = r / N, e = r - sc * N;
      for (int s = 0; s < a.S; ++s) {
        float* Hs = a.H[s] + static_cast<long long>(b0 + sc) * a.stride[s];
        if (a.ones[s]) {
          if (e == 0) {
#pragma unroll
            for (int n = 0; n < NP; ++n) if (n < N) Hs[n] = 1.0f;
          }
        } else {
          const int k = a.k[s];
          float* dst = Hs + e * N;
#pragma unroll
          for (int n = 0; n < NP; ++n) if (n < N) dst[n] = rank[n] < k ? 1.0f : 0.0f;
        }
      }
    }
  }
}

template <bool FROM_CORR>
static int launch_topk_small(const float* x, const float* corr, int B, int N, int D,
                             const TopkArgs& a, float* corr_out, cudaStream_t stream) {
  auto smem_bytes = [&](int sg) -> size_t {
    size_t rows = static_cast<size_t>(sg) * N;
    size_t f = (FROM_CORR ? 0 : rows * (D + 4)) + ((rows + 3) & ~size_t(3)) + rows * (N | 1) + 4;
    return f * 4;
  };
  int SG = GN_THREADS / N; if (SG < 1) SG = 1;
  while (SG > 1 && smem_bytes(SG) > 100 * 1024) --SG;    // 2 CTAs / SM (register-limited)
  size_t smem = smem_bytes(SG);
  if (smem > 227 * 1024) return GN_E_SHAPE;
  int ngroups = (B + SG - 1) / SG;
  int ctas_per_sm = smem > 113 * 1024 ? 1 : 2;
  int grid = ngroups < GN_SM_COUNT * ctas_per_sm ? ngroups : GN_SM_COUNT * ctas_per_sm;
  if (grid < 1) grid = 1;
  const int NP = (N + 3) & ~3;
#define GN_TOPK_SMALL(NPV)                                                                       \
  case NPV: {                                                                                    \
    auto kern = corr_topk_small_kernel<NPV, FROM_CORR>;                                          \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                         static_cast<int>(smem));                                \
    if (e != cudaSuccess) return static_cast<int>(e);                                            \
    { ProfScope ps__(FROM_CORR ? "topk_h" : "corr_topk_h", stream);                              \
      kern<<<grid, GN_THREADS, smem, stream>>>(x, corr, B, N, D, SG, a, corr_out); }             \
  } break;
  switch (NP) {
    GN_TOPK_SMALL(4) GN_TOPK_SMALL(8) GN_TOPK_SMALL(12) GN_TOPK_SMALL(16)
    default: return GN_E_SHAPE;
  }
#undef GN_TOPK_SMALL
  GN_LAUNCH_CHECK();
  return GN_OK;
}

template <bool FROM_CORR>
static int launch_topk(const float* x, const float* corr, int B, int N, int D,
                       const TopkArgs& a, float* corr_out, cudaStream_t stream) {
  if (N <= 16) return launch_topk_small<FROM_CORR>(x, corr, B, N, D, a, corr_out, stream);
  // scenes per group: enough rows to keep 8 warps busy, bounded by shared memory
  int SG = 96 / N; if (SG < 1) SG = 1; if (SG > 8) SG = 8;
  auto smem_bytes = [&](int sg) -> size_t {
    size_t rows = static_cast<size_t>(sg) * N;
    size_t f = (FROM_CORR ? 0 : rows * (D + 4)) + ((rows + 3) & ~size_t(3)) +
               ((rows * (N + 1) + 3) & ~size_t(3));
    return f * 4 + rows * ((N + 3) & ~3);
  };
  while (SG > 1 && smem_bytes(SG) > 100 * 1024) --SG;
  size_t smem = smem_bytes(SG);
  if (smem > 227 * 1024) return GN_E_SHAPE;
  int ngroups = (B + SG - 1) / SG;
  int ctas_per_sm = smem > 113 * 1024 ? 1 : (smem > 56 * 1024 ? 2 : 4);
  int grid = ngroups < GN_SM_COUNT * ctas_per_sm ? ngroups : GN_SM_COUNT * ctas_per_sm;
  if (grid < 1) grid = 1;
  int P = 1; while (P < N) P <<= 1;
#define GN_TOPK_CASE(PP)                                                                    \
  case PP: {                                                                                \
    auto kern = corr_topk_kernel<PP, FROM_CORR>;                                            \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         static_cast<int>(smem));                           \
    if (e != cudaSuccess) return static_cast<int>(e);                                       \
    { ProfScope ps__(FROM_CORR ? "topk_h" : "corr_topk_h", stream);                         \
      kern<<<grid, GN_THREADS, smem, stream>>>(x, corr, B, N, D, SG, a, corr_out); }        \
  } break;
  switch (P) {
    GN_TOPK_CASE(1) GN_TOPK_CASE(2) GN_TOPK_CASE(4) GN_TOPK_CASE(8)
    GN_TOPK_CASE(16) GN_TOPK_CASE(32) GN_TOPK_CASE(64)
    default: return GN_E_SHAPE;
  }
#undef GN_TOPK_CASE
  GN_LAUNCH_CHECK();
  return GN_OK;
}

static int fill_args(TopkArgs& a, int N, const int32_t* scales, int S, float* const* H,
                     const int64_t* strides) {
  if (S < 1 || S > GN_MAX_SCALES) return GN_E_SHAPE;
  a.S = S;
  for (int s = 0; s < S; ++s) {
    if (H[s] == nullptr) return GN_E_NULL;
    int sc = scales[s];
    if (sc > N) return GN_E_SCALE;               // reference: RuntimeError at :382
    a.ones[s] = (sc == N);
    a.k[s] = sc < 1 ? 1 : sc;                    // :378-380
    a.H[s] = H[s];
    long long e = a.ones[s] ? 1 : N;
    a.stride[s] = strides ? strides[s] : e * N;
    if (a.stride[s] < e * N) return GN_E_SHAPE;
  }
  return GN_OK;
}

}  // namespace gn

extern "C" int gn_corr_topk_h(const float* x, int32_t B, int32_t N, int32_t D,
                              const int32_t* scales, int32_t S,
                              float* const* H_out, const int64_t* H_scene_stride,
                              float* corr_out, gn_stream_t stream) {
  if (!x || !scales || !H_out) return GN_E_NULL;
  if (B < 0 || N < 1 || N > GN_MAX_AGENTS || D < 4 || (D & 3)) return GN_E_SHAPE;
  if (reinterpret_cast<uintptr_t>(x) & 15) return GN_E_ALIGN;
  gn::TopkArgs a;
  int rc = gn::fill_args(a, N, scales, S, H_out, H_scene_stride);
  if (rc != GN_OK) return rc;
  if (B == 0) return GN_OK;
  return gn::launch_topk<false>(x, nullptr, B, N, D, a, corr_out, static_cast<cudaStream_t>(stream));
}

extern "C" int gn_topk_h(const float* corr, int32_t B, int32_t N, int32_t scale,
                         float* H_out, int64_t H_scene_stride, gn_stream_t stream) {
  if (!corr || !H_out) return GN_E_NULL;
  if (B < 0 || N < 1 || N > GN_MAX_AGENTS) return GN_E_SHAPE;
  gn::TopkArgs a;
  int32_t sc = scale;
  float* Hp = H_out;
  int64_t st = H_scene_stride;
  int rc = gn::fill_args(a, N, &sc, 1, &Hp, H_scene_stride > 0 ? &st : nullptr);
  if (rc != GN_OK) return rc;
  if (B == 0) return GN_OK;
  return gn::launch_topk<true>(nullptr, corr, B, N, 4, a, nullptr, static_cast<cudaStream_t>(stream));
}


// ---------------------------------------------------------------------------
// PastEncoder front-end (model/GroupNet_nba.py:269-280).  In eval mode
// input_fc -> [x ; pos_enc] -> pos_encoder.fc -> input_fc2 -> add_category -> input_fc3 has no
// nonlinearity, so the host folds it into one affine map (groupnet_b200/encoder.py):
//   ftraj[b,n,:] = Mt^T u[b,n,:] + bias_agent[n,:],   u = the agent's (T x in_dim) window, K = T*in_dim
// One thread per (row, 4 output columns); Mt / bias table in shared memory; HBM-bound
// (K*4 bytes in, 256 bytes out per agent row).
// ---------------------------------------------------------------------------
namespace gn {
__global__ void __launch_bounds__(GN_THREADS)
past_frontend_kernel(const float* __restrict__ u, long long R, int K, int N, int C,
                     const float* __restrict__ Mt, const float* __restrict__ bias_agent,
                     float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* sM = smem;                    // [K][C]
  float* sB = sM + K * C;              // [N][C]
  for (int i = threadIdx.x; i < K * C; i += GN_THREADS) sM[i] = __ldg(Mt + i);
  for (int i = threadIdx.x; i < N * C; i += GN_THREADS) sB[i] = __ldg(bias_agent + i);
  __syncthreads();
  const int c4n = C >> 2;
  const long long total = R * c4n;
  for (long long i = static_cast<long long>(blockIdx.x) * GN_THREADS + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * GN_THREADS) {
    const long long r = i / c4n;
    const int c = static_cast<int>(i - r * c4n) * 4;
    const int n = static_cast<int>(r % N);
    float4 acc = *reinterpret_cast<const float4*>(sB + n * C + c);
    const float* ur = u + r * K;
    for (int k = 0; k < K; ++k) {
      const float x = __ldg(ur + k);
      const float4 m = *reinterpret_cast<const float4*>(sM + k * C + c);
      acc.x = fmaf(x, m.x, acc.x); acc.y = fmaf(x, m.y, acc.y);
      acc.z = fmaf(x, m.z, acc.z); acc.w = fmaf(x, m.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + r * C + c) = acc;
  }
}
}  // namespace gn

extern "C" int gn_past_frontend(const float* inputs, int64_t R, int32_t K, int32_t N, int32_t C,
                                const float* Mt, const float* bias_agent, float* out, gn_stream_t stream) {
  if (!inputs || !Mt || !bias_agent || !out) return GN_E_NULL;
  if (R < 0 || K < 1 || N < 1 || C < 4 || (C & 3)) return GN_E_SHAPE;
  if (R == 0) return GN_OK;
  size_t smem = (static_cast<size_t>(K) + N) * C * 4;
  if (smem > 200 * 1024) return GN_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(gn::past_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  long long work = (R * (C >> 2) + GN_THREADS - 1) / GN_THREADS;
  int grid = work < GN_SM_COUNT * 8 ? static_cast<int>(work) : GN_SM_COUNT * 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  { gn::ProfScope ps__("past_frontend", st);
    gn::past_frontend_kernel<<<grid, GN_THREADS, smem, st>>>(inputs, R, K, N, C, Mt, bias_agent, out); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}
