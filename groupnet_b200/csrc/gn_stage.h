// Internal host-side interfaces between the translation units of libgroupnet_b200.
#pragma once
#include <cstring>
#include <cuda_bf16.h>
#include "gn_common.cuh"

namespace gn {

// generic tcgen05 row-tile linear (gn_tc_linear.cu)
struct TcLinArgs {
  const void* A0; int a0_is_f32; long long lda0; int K0;   // columns [0, K0)
  const float* A1; long long lda1; int K1;                 // columns [K0, K0+K1), fp32, optional
  float a_div;                                             // != 0: A is divided by it (true division)
  const __nv_bfloat16* W; int Ntot; int n0; int N;         // canonical [K/8][Ntot][8]; rows [n0, n0+N)
  const float* bias;                                       // [Ntot] or null, indexed n0 + col
  int relu;
  const float* rowscale; int rs_ld; int rs_shift;          // scale = rowscale[row*rs_ld + ((n0+col) >> rs_shift)]
  const float* bias_mat; int bm_T; int bm_ld;              // out += sum_t rowscale[row*rs_ld + t] * bias_mat[t*bm_ld + n0+col]
  void* out; int out_is_f32; long long ldo; int out_col0;  // row-major, written at column out_col0 + col
  float out_div;                                           // != 0: result divided by it
  long long R;
};
int launch_tc_linear(const TcLinArgs& a, const char* name, cudaStream_t st);

// fused per-edge chain on tcgen05 (gn_edge_mlp_tc.cu); pair = node2edge fused in (pairwise layer)
bool edge_chain_pair_fits(int N);
int launch_edge_chain_tc(bool pair, const float* edges, const float* xprime, const float* pq,
                         int N, int E, int T, long long R, const gn_stage_weights* w,
                         const float* U, int noise_mode, unsigned long long seed, long long scene_offset,
                         int stage_index, float* dist_out, float* edge_feat, cudaStream_t st);

// fused pairwise edge2node on tensor cores (gn_pair_agg_tc.cu)
bool pair_agg_fits(int N, int D, int T);
int launch_pair_agg_tc(const float* h, const float* edge_feat, int B, int N, int T,
                       const gn_stage_weights* w, float* agg, cudaStream_t st);

// fused hyper edge_aggregation on tensor cores (gn_hyper_agg_tc.cu)
bool hyper_agg_fits(int D, int T);
int launch_hyper_agg_tc(const float* eo, const float* edge_feat, long long R, int T,
                        const gn_stage_weights* w, float* ef, cudaStream_t st);

// fused wide hyper aggregation, h_dim 256: H@h gather + T MLPs + H^T scatter (gn_hyper_fused_tc.cu)
bool hyper_fused_fits(int N, int E, int D, int T);
bool hyper_fused_post_fits(int Dout, long long ld_out);
// node_out != nullptr: the closing MLP on [agg | h] / N runs in the same kernel and agg is not written
int launch_hyper_fused_tc(const float* h, const float* H, const float* edge_feat, int B, int N, int T,
                          long long hstride, const gn_stage_weights* w, float* agg,
                          float* node_out, long long ld_out, int Dout, cudaStream_t st);

// fused hyper layer tail for h_dim 64: gather + T MLPs + scatter + closing MLP (gn_hyper_fused64_tc.cu)
bool hyper_fused64_fits(int N, int E, int D, int T, int Dout, long long ld_out);
int launch_hyper_fused64_tc(const float* h, const float* H, const float* edge_feat, int B, int N, int T,
                            long long hstride, const gn_stage_weights* w, float* node_out, long long ld_out,
                            int Dout, cudaStream_t st);

// fused wide node prologue h -> x', pq for h_dim 256 (gn_node_pre256_tc.cu)
bool node_pre256_fits(int D);
int launch_node_pre256_tc(const float* h, long long R, const gn_stage_weights* w, float* xprime, float* pq,
                          cudaStream_t st);

// fused node-level GEMM chains on tensor cores (gn_node_chain_tc.cu)
struct NodeChainStep {
  const __nv_bfloat16* W;     // canonical (N x K)
  const float* bias;          // [N] or null
  int K, N, relu;
  float* out; long long ldo; int out_col0;   // optional fp32 store of this step's output
};
struct NodeChainArgs {
  const float* A0; long long lda0; int K0;   // input columns [0, K0)
  const float* A1; long long lda1; int K1;   // input columns [K0, K0+K1), optional
  float a_div;                               // != 0: input divided by it
  int nsteps;
  NodeChainStep step[3];
  long long R;
};
int launch_node_chain_tc(const NodeChainArgs& a, const char* name, cudaStream_t st);

// 3xTF32 chains on tcgen05 (gn_chain_tf32.cu): the fp32-grade tensor-core path, precision GN_TF32X3
bool edge_chain_tf32_fits(bool pair, int N, int T);
int launch_edge_chain_tf32(bool pair, const float* edges, const float* ypre, const float* pq,
                           int N, int E, int T, long long R, const gn_stage_weights* w,
                           const float* U, int noise_mode, unsigned long long seed, long long scene_offset,
                           int stage_index, float* dist_out, float* edge_feat, cudaStream_t st);
bool node_pre_tf32_fits(int D);
int launch_node_pre_tf32(const float* h, long long R, int D, const gn_stage_weights* w, float* xprime, float* pq,
                         float* ypre, cudaStream_t st);
bool agg_in_tf32_fits(int D, int T);
int launch_agg_in_tf32(const float* h, long long R, int D, int T, const gn_stage_weights* w, float* P, cudaStream_t st);
bool agg_out_tf32_fits(int D, int T);
int launch_agg_out_tf32(const float* G, const float* S, long long R, int D, int T, const gn_stage_weights* w,
                        float* agg, cudaStream_t st);
// fused pairwise aggregation, 3xTF32 (gn_pair_agg_tf32.cu): replaces agg_in_tf32 + edge2node_pair + agg_out_tf32
bool pair_agg_tf32_fits(int N, int D, int T);
int launch_pair_agg_tf32(const float* h, const float* edge_feat, int B, int N, int T,
                         const gn_stage_weights* w, float* agg, cudaStream_t st);
bool hyper_agg_tf32_fits(int D, int T);
int launch_hyper_agg_tf32(const float* eo, const float* edge_feat, long long R, int D, int T,
                          const gn_stage_weights* w, float* ef, cudaStream_t st);
bool node_post_tf32_fits(int D, int Dout, long long ld_out, const float* node_out);
int launch_node_post_tf32(const float* agg, const float* h, long long R, int D, int Nagents, int Dout,
                          const gn_stage_weights* w, float* node_out, long long ld_out, cudaStream_t st);

// stage driver (gn_stage_simt.cu)
int stage_fwd(const gn_stage_cfg* c, const gn_stage_weights* w, const float* h, const float* H,
              const float* U, float* node_out, float* dist_out, void* ws, size_t ws_bytes, cudaStream_t st);
size_t stage_workspace_bytes(const gn_stage_cfg* c);
int stage_saved_offsets(const gn_stage_cfg* c, size_t* out5);

// backward of one stage, fp32 (gn_train.cu)
size_t stage_bwd_workspace_bytes(const gn_stage_cfg* c);
int stage_bwd(const gn_stage_cfg* c, const gn_train_params* P, const float* h, const float* H,
              const float* xprime, const float* pq, const float* edges, const float* efeat, const float* agg,
              const float* d_out, long long ld_dout, const float* d_dist, float* d_h,
              void* ws, size_t ws_bytes, cudaStream_t st);
int stage_launch_count(const gn_stage_cfg* c);

}  // namespace gn
