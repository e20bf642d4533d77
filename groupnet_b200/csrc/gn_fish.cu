// Group-wise operators of the fish model, fp32 SIMT kernels behind the C ABI:
//   compute_alpha_im            model/encoder.py:261-303     gn_fish_alpha_im
//   MLPHGE                      model/encoder.py:200-256     gn_fish_bmm_t (normalised) + gn_fish_mlp
//   HyperEdgeAttention          model/encoder.py:102-197     gn_fish_mlp + gn_fish_hga_core + gn_fish_bmm_t
//   TemporalGATLayer            model/encoder.py:331-467     gn_fish_mlp + gn_fish_gat_edges + gn_fish_bmm_t
//   build_dynamic_graph_and_hypergraph   utilities/utils.py:191-244   gn_fish_dynamic_graph
// The reference materialises (B, E, N, M) masks and (B, N, M, 2 hidden) attention inputs and loops over the batch in
// Python; here a scene's small matrices sit in shared memory and every op is one pass.  Eval-mode semantics: the
// BatchNorm layers are folded into the Linears on the host (running statistics), dropout is the identity.
// Shapes are the fish model's: N agents <= 64, M hyperedges <= 32, E = N (N - 1) edges, features <= 512 wide.
#include "gn_common.cuh"

namespace gn {
namespace fish {

constexpr int ACT_NONE = 0, ACT_LEAKY = 1, ACT_ELU = 2;

__device__ __forceinline__ float act(float v, int kind, float slope) {
  if (kind == ACT_LEAKY) return v > 0.f ? v : slope * v;
  if (kind == ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}

// ---------------------------------------------------------------------------------------------
// compute_alpha_im (model/encoder.py:261-303), one CTA per scene:
//   mask[e][m] = (sum_n rel_rec[e][n] I[n][m] > 0) && (sum_n rel_send[e][n] I[n][m] > 0)
//   out[n][m]  = sum_e alpha[e] mask[e][m] rel_rec[e][n] / (sum_n' I[n'][m] - 1 + 1e-8)
// rel_rec / rel_send are (E, N) per scene (stride rel_stride floats between scenes; 0 = shared by all scenes).
// ---------------------------------------------------------------------------------------------
__global__ void alpha_im_kernel(const float* __restrict__ alpha, const float* __restrict__ I, const float* __restrict__ rec,
                                const float* __restrict__ snd, long long rel_stride, int B, int E, int N, int M,
                                float* __restrict__ out) {
  extern __shared__ float sm[];
  float* sI = sm;                    // [N][M]
  float* sA = sI + N * M;            // [E][M] masked alpha
  float* sCnt = sA + E * M;          // [M]
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float* Ib = I + static_cast<size_t>(b) * N * M;
    const float* rb = rec + static_cast<size_t>(b) * rel_stride;
    const float* sb = snd + static_cast<size_t>(b) * rel_stride;
    for (int i = threadIdx.x; i < N * M; i += blockDim.x) sI[i] = __ldg(Ib + i);
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      float c = 0.f;
      for (int n = 0; n < N; ++n) c += sI[n * M + m];
      sCnt[m] = c;
    }
    for (int i = threadIdx.x; i < E * M; i += blockDim.x) {
      const int e = i / M, m = i - e * M;
      float r = 0.f, s = 0.f;
      for (int n = 0; n < N; ++n) {
        const float im = sI[n * M + m];
        r = fmaf(__ldg(rb + e * N + n), im, r);
        s = fmaf(__ldg(sb + e * N + n), im, s);
      }
      sA[i] = (r > 0.f && s > 0.f) ? __ldg(alpha + static_cast<size_t>(b) * E + e) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * M; i += blockDim.x) {
      const int n = i / M, m = i - n * M;
      float acc = 0.f;
      for (int e = 0; e < E; ++e) acc = fmaf(sA[e * M + m], __ldg(rb + e * N + n), acc);
      out[static_cast<size_t>(b) * N * M + i] = acc / (sCnt[m] - 1.f + 1e-8f);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Per-scene C = op(A)^T B:  C[c][f] = sum_r w[r] A[r][c] B[r][f]        (A: (R, Cn), B: (R, F), C: (Cn, F))
//   norm_cols != 0: A's columns are first normalised, A[r][c] / (sum_r' A[r'][c] + 1e-8)   (MLPHGE, :236-238)
//   roww: optional per-row weight (R) per scene                                             (TemporalGATLayer, :451-454)
// covers einsum('bnm,bnf->bmf'), ('bmn,bmf->bnf'), ('behd,ben->bnhd').  A_stride = 0: A shared by all scenes.
// ---------------------------------------------------------------------------------------------
__global__ void bmm_t_kernel(const float* __restrict__ A, long long A_stride, const float* __restrict__ Bm,
                             const float* __restrict__ roww, int batch, int R, int Cn, int F, int norm_cols,
                             float* __restrict__ C) {
  extern __shared__ float sm[];
  float* sA = sm;                    // [R][Cn]
  float* sCol = sA + R * Cn;         // [Cn]
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    const float* Ab = A + static_cast<size_t>(b) * A_stride;
    const float* Bb = Bm + static_cast<size_t>(b) * R * F;
    for (int i = threadIdx.x; i < R * Cn; i += blockDim.x) {
      float v = __ldg(Ab + i);
      if (roww != nullptr) v *= __ldg(roww + static_cast<size_t>(b) * R + i / Cn);
      sA[i] = v;
    }
    __syncthreads();
    if (norm_cols) {
      for (int c = threadIdx.x; c < Cn; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += sA[r * Cn + c];
        sCol[c] = s + 1e-8f;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < R * Cn; i += blockDim.x) sA[i] = sA[i] / sCol[i % Cn];
      __syncthreads();
    }
    for (int i = threadIdx.x; i < Cn * F; i += blockDim.x) {
      const int c = i / F, f = i - c * F;
      float acc = 0.f;
      for (int r = 0; r < R; ++r) acc = fmaf(sA[r * Cn + c], __ldg(Bb + r * F + f), acc);
      C[static_cast<size_t>(b) * Cn * F + i] = acc;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Row MLP: up to 3 Linears (BatchNorm folded in by the host) with an activation after each, over R rows.
// A tile = 32 rows; activations live TRANSPOSED in shared memory ([k][32 rows]) so a thread that owns output column n
// reads one weight Wt[k][n] (coalesced over the threads) and eight 128-bit broadcast loads of the 32 row values per k.
// ---------------------------------------------------------------------------------------------
struct MlpArgs {
  const float* x; long long ldx; long long R; int K0;
  int nlayers;
  const float* Wt[3];        // [K][N]: nn.Linear.weight transposed (and BN-scaled)
  const float* bias[3];      // [N] or null
  int N[3]; int act[3]; float slope[3];
  float* out; long long ldo;
  int kmax;                  // widest activation (smem row count of one buffer)
};
constexpr int MLP_ROWS = 32, MLP_THREADS = 128;

__global__ void __launch_bounds__(MLP_THREADS) mlp_kernel(const MlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* buf0 = sm;
  float* buf1 = sm + static_cast<size_t>(a.kmax) * MLP_ROWS;
  const long long ntiles = (a.R + MLP_ROWS - 1) / MLP_ROWS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r0 = tile * MLP_ROWS;
    for (int i = threadIdx.x; i < MLP_ROWS * a.K0; i += MLP_THREADS) {
      const int k = i / MLP_ROWS, r = i - k * MLP_ROWS;       // row fastest: conflict-free transposed store; the 32 rows'
      buf0[i] = (r0 + r < a.R) ? __ldg(a.x + (r0 + r) * a.ldx + k) : 0.f;   // cache lines are reused by the next 31 k
    }
    __syncthreads();
    float* in = buf0;
    float* nxt = buf1;
    int K = a.K0;
    for (int l = 0; l < a.nlayers; ++l) {
      const int N = a.N[l];
      const bool last = l == a.nlayers - 1;
      for (int n = threadIdx.x; n < N; n += MLP_THREADS) {
        float acc[MLP_ROWS];
        const float bv = a.bias[l] != nullptr ? __ldg(a.bias[l] + n) : 0.f;
#pragma unroll
        for (int r = 0; r < MLP_ROWS; ++r) acc[r] = bv;
        const float* w = a.Wt[l] + n;
        int k = 0;
        for (; k + 8 <= K; k += 8) {                 // eight weight loads in flight (they come from L2 / L1)
          float wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + static_cast<size_t>(k + u) * N);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4* row = reinterpret_cast<const float4*>(in + (k + u) * MLP_ROWS);
#pragma unroll
            for (int q = 0; q < MLP_ROWS / 4; ++q) {
              const float4 v = row[q];
              acc[4 * q] = fmaf(v.x, wv[u], acc[4 * q]); acc[4 * q + 1] = fmaf(v.y, wv[u], acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(v.z, wv[u], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v.w, wv[u], acc[4 * q + 3]);
            }
          }
        }
        for (; k < K; ++k) {
          const float wv = __ldg(w + static_cast<size_t>(k) * N);
          const float4* row = reinterpret_cast<const float4*>(in + k * MLP_ROWS);
#pragma unroll
          for (int q = 0; q < MLP_ROWS / 4; ++q) {
            const float4 v = row[q];
            acc[4 * q] = fmaf(v.x, wv, acc[4 * q]); acc[4 * q + 1] = fmaf(v.y, wv, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v.z, wv, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v.w, wv, acc[4 * q + 3]);
          }
        }
#pragma unroll
        for (int r = 0; r < MLP_ROWS; ++r) acc[r] = act(acc[r], a.act[l], a.slope[l]);
        if (last) {
#pragma unroll
          for (int r = 0; r < MLP_ROWS; ++r)
            if (r0 + r < a.R) a.out[(r0 + r) * a.ldo + n] = acc[r];
        } else {
#pragma unroll
          for (int q = 0; q < MLP_ROWS / 4; ++q)
            *reinterpret_cast<float4*>(nxt + n * MLP_ROWS + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
      }
      __syncthreads();
      float* t = in; in = nxt; nxt = t;
      K = N;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// HyperEdgeAttention core (model/encoder.py:160-181), one CTA per scene.  The attention input cat(e_proj[m], v_proj[n]) .
// a splits into two dots: logit[n][m] = leaky(se[m] + sv[n]), se = e_proj a[:Hd], sv = v_proj a[Hd:];
// masked_fill(I == 0, -inf); softmax over the NODES of logit / 100 (:173); NaN (no member) -> 0; then
// v1[n][f] = sum_m alpha[n][m] e_HG[m][f]  (:177).
// ---------------------------------------------------------------------------------------------
__global__ void hga_core_kernel(const float* __restrict__ e_proj, const float* __restrict__ v_proj,
                                const float* __restrict__ avec, const float* __restrict__ I,
                                const float* __restrict__ e_hg, int B, int N, int M, int Hd, int F, float slope,
                                float* __restrict__ v1) {
  extern __shared__ float sm[];
  float* se = sm;              // [M]
  float* sv = se + M;          // [N]
  float* sAl = sv + N;         // [N][M]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int i = warp; i < M + N; i += nwarp) {                  // one warp per dot product
      const float* row = i < M ? e_proj + (static_cast<size_t>(b) * M + i) * Hd : v_proj + (static_cast<size_t>(b) * N + (i - M)) * Hd;
      const float* av = i < M ? avec : avec + Hd;
      float acc = 0.f;
      for (int k = lane; k < Hd; k += 32) acc = fmaf(__ldg(row + k), __ldg(av + k), acc);
      acc = warp_sum(acc);
      if (lane == 0) { if (i < M) se[i] = acc; else sv[i - M] = acc; }
    }
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x) {          // softmax over the nodes of hyperedge m
      const float* Ib = I + static_cast<size_t>(b) * N * M;
      float mx = -INFINITY;
      for (int n = 0; n < N; ++n)
        if (__ldg(Ib + n * M + m) != 0.f) mx = fmaxf(mx, act(se[m] + sv[n], ACT_LEAKY, slope) / 100.f);
      float den = 0.f;
      for (int n = 0; n < N; ++n)
        if (__ldg(Ib + n * M + m) != 0.f) den += expf(act(se[m] + sv[n], ACT_LEAKY, slope) / 100.f - mx);
      for (int n = 0; n < N; ++n) {
        const bool in_m = __ldg(Ib + n * M + m) != 0.f;
        sAl[n * M + m] = in_m ? expf(act(se[m] + sv[n], ACT_LEAKY, slope) / 100.f - mx) / den : 0.f;   // empty hyperedge: nan_to_num -> 0
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * F; i += blockDim.x) {
      const int n = i / F, f = i - n * F;
      float acc = 0.f;
      for (int m = 0; m < M; ++m) acc = fmaf(sAl[n * M + m], __ldg(e_hg + (static_cast<size_t>(b) * M + m) * F + f), acc);
      v1[static_cast<size_t>(b) * N * F + i] = acc;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// TemporalGATLayer edge stage (model/encoder.py:404-447), one CTA per scene: per edge e and head h
//   src = rel_send[e] @ v_proj, tgt = rel_rec[e] @ v_proj              (dense (E, N) x (N, H D) products)
//   a_ij = leaky(src . a_fwd[h]) / 500,  a_ji = leaky(tgt . a_bwd[h]) / 500
//   s_ij = exp(a_ij - max), s_ji = exp(a_ji - max);  alpha_ij = exp(s_ij) / (exp(s_ij) + exp(s_ji))   (as written, :425-431)
//   edge_input[e][h] = [alpha_ij src | alpha_ji tgt]  (2 D);  alpha_out[e][h] = alpha_ij
// ---------------------------------------------------------------------------------------------
__global__ void gat_edges_kernel(const float* __restrict__ v_proj, const float* __restrict__ rec, const float* __restrict__ snd,
                                 long long rel_stride, const float* __restrict__ a_fwd, const float* __restrict__ a_bwd,
                                 int B, int E, int N, int H, int D, float slope,
                                 float* __restrict__ edge_input, float* __restrict__ alpha_out) {
  extern __shared__ float sm[];
  float* sV = sm;                                    // [N][H*D]
  float* sSrc = sV + N * H * D;                      // per warp: [H*D] src | [H*D] tgt
  const int HD = H * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float* wsrc = sSrc + warp * 2 * HD;
  float* wtgt = wsrc + HD;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int i = threadIdx.x; i < N * HD; i += blockDim.x) sV[i] = __ldg(v_proj + static_cast<size_t>(b) * N * HD + i);
    __syncthreads();
    const float* rb = rec + static_cast<size_t>(b) * rel_stride;
    const float* sb = snd + static_cast<size_t>(b) * rel_stride;
    for (int e = warp; e < E; e += nwarp) {                       // one warp per edge
      for (int c = lane; c < HD; c += 32) {
        float s = 0.f, t = 0.f;
        for (int n = 0; n < N; ++n) {
          const float v = sV[n * HD + c];
          s = fmaf(__ldg(sb + e * N + n), v, s);
          t = fmaf(__ldg(rb + e * N + n), v, t);
        }
        wsrc[c] = s; wtgt[c] = t;
      }
      __syncwarp();
      for (int h = 0; h < H; ++h) {
        float df = 0.f, db = 0.f;
        for (int k = lane; k < D; k += 32) {
          df = fmaf(wsrc[h * D + k], __ldg(a_fwd + h * D + k), df);
          db = fmaf(wtgt[h * D + k], __ldg(a_bwd + h * D + k), db);
        }
        df = warp_sum(df); db = warp_sum(db);
        const float aij = act(df, ACT_LEAKY, slope) / 500.f, aji = act(db, ACT_LEAKY, slope) / 500.f;
        const float mx = fmaxf(aij, aji);
        const float sij = expf(aij - mx), sji = expf(aji - mx);
        const float eij = expf(sij), eji = expf(sji);
        const float al_ij = eij / (eij + eji), al_ji = eji / (eij + eji);
        float* dst = edge_input + ((static_cast<size_t>(b) * E + e) * H + h) * 2 * D;
        for (int k = lane; k < D; k += 32) {
          dst[k] = al_ij * wsrc[h * D + k];
          dst[D + k] = al_ji * wtgt[h * D + k];
        }
        if (lane == 0) alpha_out[(static_cast<size_t>(b) * E + e) * H + h] = al_ij;
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

// build_dynamic_graph_and_hypergraph (utilities/utils.py:191-244): argmax types, rows / columns of type 0 zeroed
__global__ void dynamic_graph_kernel(const float* __restrict__ z_cg, const float* __restrict__ z_hg,
                                     const float* __restrict__ rec, const float* __restrict__ snd, long long rel_stride,
                                     const float* __restrict__ I, int B, int E, int N, int M, int Lc, int Lh,
                                     float* __restrict__ new_rec, float* __restrict__ new_snd, float* __restrict__ new_I,
                                     int64_t* __restrict__ edge_types, int64_t* __restrict__ hyper_types) {
  extern __shared__ int smi[];
  int* sEt = smi;          // [E]
  int* sHt = smi + E;      // [M]
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int i = threadIdx.x; i < E + M; i += blockDim.x) {
      const bool edge = i < E;
      const float* z = edge ? z_cg + (static_cast<size_t>(b) * E + i) * Lc : z_hg + (static_cast<size_t>(b) * M + (i - E)) * Lh;
      const int L = edge ? Lc : Lh;
      int best = 0; float bv = __ldg(z);
      for (int l = 1; l < L; ++l) { const float v = __ldg(z + l); if (v > bv) { bv = v; best = l; } }   // first maximum, like argmax
      if (edge) { sEt[i] = best; edge_types[static_cast<size_t>(b) * E + i] = best; }
      else { sHt[i - E] = best; hyper_types[static_cast<size_t>(b) * M + (i - E)] = best; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < E * N; i += blockDim.x) {
      const bool keep = sEt[i / N] != 0;
      new_rec[static_cast<size_t>(b) * E * N + i] = keep ? __ldg(rec + static_cast<size_t>(b) * rel_stride + i) : 0.f;
      new_snd[static_cast<size_t>(b) * E * N + i] = keep ? __ldg(snd + static_cast<size_t>(b) * rel_stride + i) : 0.f;
    }
    for (int i = threadIdx.x; i < N * M; i += blockDim.x)
      new_I[static_cast<size_t>(b) * N * M + i] = sHt[i % M] != 0 ? __ldg(I + static_cast<size_t>(b) * N * M + i) : 0.f;
    __syncthreads();
  }
}

static int grid_scenes(int B, int per_sm) {
  const long long g = static_cast<long long>(GN_SM_COUNT) * per_sm;
  return static_cast<int>(B < g ? (B < 1 ? 1 : B) : g);
}

}  // namespace fish
}  // namespace gn

using namespace gn;
using namespace gn::fish;

#define FISH_CHECK_PTR(p) do { if (!(p)) return GN_E_NULL; } while (0)

extern "C" int gn_fish_alpha_im(const float* alpha_ij, const float* I_HG, const float* rel_rec, const float* rel_send,
                                int64_t rel_stride, int32_t B, int32_t E, int32_t N, int32_t M, float* out, gn_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FISH_CHECK_PTR(alpha_ij); FISH_CHECK_PTR(I_HG); FISH_CHECK_PTR(rel_rec); FISH_CHECK_PTR(rel_send); FISH_CHECK_PTR(out);
  if (B < 0 || E < 1 || N < 1 || M < 1 || N > 64 || M > 32 || E > 64 * 63) return GN_E_SHAPE;
  if (B == 0) return GN_OK;
  const size_t smem = (static_cast<size_t>(N) * M + static_cast<size_t>(E) * M + M) * 4;
  if (smem > 200 * 1024) return GN_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(alpha_im_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  { ProfScope ps__("fish_alpha_im", st);
    alpha_im_kernel<<<grid_scenes(B, 8), 128, smem, st>>>(alpha_ij, I_HG, rel_rec, rel_send, rel_stride, B, E, N, M, out); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

extern "C" int gn_fish_bmm_t(const float* A, int64_t A_stride, const float* Bm, const float* roww, int32_t batch, int32_t R, int32_t Cn,
                             int32_t F, int32_t norm_cols, float* C, gn_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FISH_CHECK_PTR(A); FISH_CHECK_PTR(Bm); FISH_CHECK_PTR(C);
  if (batch < 0 || R < 1 || Cn < 1 || F < 1) return GN_E_SHAPE;
  if (batch == 0) return GN_OK;
  const size_t smem = (static_cast<size_t>(R) * Cn + Cn) * 4;
  if (smem > 200 * 1024) return GN_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(bmm_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  { ProfScope ps__("fish_bmm_t", st);
    bmm_t_kernel<<<grid_scenes(batch, 8), 128, smem, st>>>(A, A_stride, Bm, roww, batch, R, Cn, F, norm_cols, C); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

extern "C" int gn_fish_mlp(const float* x, int64_t ldx, int64_t R, int32_t K0, int32_t nlayers, const float* const* Wt,
                           const float* const* bias, const int32_t* N, const int32_t* act_kind, const float* slope,
                           float* out, int64_t ldo, gn_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FISH_CHECK_PTR(x); FISH_CHECK_PTR(out); FISH_CHECK_PTR(Wt); FISH_CHECK_PTR(N);
  if (R < 0 || K0 < 1 || K0 > 512 || nlayers < 1 || nlayers > 3) return GN_E_SHAPE;
  if (R == 0) return GN_OK;
  MlpArgs a;
  a.x = x; a.ldx = ldx; a.R = R; a.K0 = K0; a.nlayers = nlayers; a.out = out; a.ldo = ldo; a.kmax = K0;
  for (int l = 0; l < 3; ++l) { a.Wt[l] = nullptr; a.bias[l] = nullptr; a.N[l] = 0; a.act[l] = 0; a.slope[l] = 0.f; }
  for (int l = 0; l < nlayers; ++l) {
    if (!Wt[l] || N[l] < 1 || N[l] > 512) return GN_E_SHAPE;
    a.Wt[l] = Wt[l]; a.bias[l] = bias ? bias[l] : nullptr; a.N[l] = N[l];
    a.act[l] = act_kind ? act_kind[l] : 0; a.slope[l] = slope ? slope[l] : 0.f;
    if (l + 1 < nlayers && N[l] > a.kmax) a.kmax = N[l];
  }
  const size_t smem = 2 * static_cast<size_t>(a.kmax) * MLP_ROWS * 4;
  cudaError_t e = cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long long ntiles = (R + MLP_ROWS - 1) / MLP_ROWS;
  const int per_sm = smem > 100 * 1024 ? 2 : 4;
  const long long g = static_cast<long long>(GN_SM_COUNT) * per_sm;
  { ProfScope ps__("fish_mlp", st);
    mlp_kernel<<<static_cast<int>(ntiles < g ? ntiles : g), MLP_THREADS, smem, st>>>(a); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

extern "C" int gn_fish_hga_core(const float* e_proj, const float* v_proj, const float* avec, const float* I_HG,
                                const float* e_hg, int32_t B, int32_t N, int32_t M, int32_t Hd, int32_t F, float slope, float* v1,
                                gn_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FISH_CHECK_PTR(e_proj); FISH_CHECK_PTR(v_proj); FISH_CHECK_PTR(avec); FISH_CHECK_PTR(I_HG); FISH_CHECK_PTR(e_hg); FISH_CHECK_PTR(v1);
  if (B < 0 || N < 1 || M < 1 || N > 64 || M > 32 || Hd < 1 || F < 1) return GN_E_SHAPE;
  if (B == 0) return GN_OK;
  const size_t smem = (static_cast<size_t>(M) + N + static_cast<size_t>(N) * M) * 4;
  { ProfScope ps__("fish_hga_core", st);
    hga_core_kernel<<<grid_scenes(B, 8), 128, smem, st>>>(e_proj, v_proj, avec, I_HG, e_hg, B, N, M, Hd, F, slope, v1); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

extern "C" int gn_fish_gat_edges(const float* v_proj, const float* rel_rec, const float* rel_send, int64_t rel_stride,
                                 const float* a_fwd, const float* a_bwd, int32_t B, int32_t E, int32_t N, int32_t H, int32_t D, float slope,
                                 float* edge_input, float* alpha_out, gn_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FISH_CHECK_PTR(v_proj); FISH_CHECK_PTR(rel_rec); FISH_CHECK_PTR(rel_send); FISH_CHECK_PTR(a_fwd); FISH_CHECK_PTR(a_bwd);
  FISH_CHECK_PTR(edge_input); FISH_CHECK_PTR(alpha_out);
  if (B < 0 || E < 1 || N < 1 || N > 64 || H < 1 || D < 1) return GN_E_SHAPE;
  if (B == 0) return GN_OK;
  const int threads = 256;
  const size_t smem = (static_cast<size_t>(N) * H * D + static_cast<size_t>(threads / 32) * 2 * H * D) * 4;
  if (smem > 200 * 1024) return GN_E_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(gat_edges_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  { ProfScope ps__("fish_gat_edges", st);
    gat_edges_kernel<<<grid_scenes(B, smem > 56 * 1024 ? 2 : 4), threads, smem, st>>>(v_proj, rel_rec, rel_send, rel_stride, a_fwd, a_bwd,
                                                                                     B, E, N, H, D, slope, edge_input, alpha_out); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}

extern "C" int gn_fish_dynamic_graph(const float* z_cg, const float* z_hg, const float* rel_rec, const float* rel_send,
                                     int64_t rel_stride, const float* I_HG, int32_t B, int32_t E, int32_t N, int32_t M, int32_t Lc, int32_t Lh,
                                     float* new_rec, float* new_send, float* new_I, int64_t* edge_types,
                                     int64_t* hyper_types, gn_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FISH_CHECK_PTR(z_cg); FISH_CHECK_PTR(z_hg); FISH_CHECK_PTR(rel_rec); FISH_CHECK_PTR(rel_send); FISH_CHECK_PTR(I_HG);
  FISH_CHECK_PTR(new_rec); FISH_CHECK_PTR(new_send); FISH_CHECK_PTR(new_I); FISH_CHECK_PTR(edge_types); FISH_CHECK_PTR(hyper_types);
  if (B < 0 || E < 1 || N < 1 || M < 1 || Lc < 1 || Lh < 1) return GN_E_SHAPE;
  if (B == 0) return GN_OK;
  const size_t smem = (static_cast<size_t>(E) + M) * 4;
  { ProfScope ps__("fish_dynamic_graph", st);
    dynamic_graph_kernel<<<grid_scenes(B, 8), 128, smem, st>>>(z_cg, z_hg, rel_rec, rel_send, rel_stride, I_HG, B, E, N, M, Lc, Lh,
                                                               new_rec, new_send, new_I, edge_types, hyper_types); }
  GN_LAUNCH_CHECK();
  return GN_OK;
}
