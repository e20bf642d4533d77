/*
 * groupnet_b200.h — C ABI of libgroupnet_b200.so (sm_100a).
 *
 * Drop-in boundary for the hot path of TaliMotzkin/GroupNet:
 * model/MS_HGNN_batch.py (MS_HGNN_oridinary, MS_HGNN_hyper).  The reference
 * has no FFI of its own (it is pure PyTorch); its boundary is the nn.Module
 * API.  groupnet_b200/layers.py keeps that API (constructor + forward
 * signatures, state_dict schema) and calls the entry points below through
 * ctypes.  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *  - plain C types only; all tensor arguments are DEVICE pointers to
 *    contiguous fp32 buffers owned by the caller; `stream` is a cudaStream_t
 *    passed as void* (NULL = legacy default stream);
 *  - the library never allocates device memory and keeps no global state:
 *    the caller passes outputs and a workspace (gn_stage_workspace_bytes);
 *  - return value: 0 = ok, <0 = GN_E_* argument error (nothing launched),
 *    >0 = cudaError_t from a launch; nothing throws across the ABI;
 *  - re-entrant: one process per GPU or several streams may call concurrently.
 */
#ifndef GROUPNET_B200_H
#define GROUPNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GN_ABI_VERSION 6     /* 3: precision GN_TF32X3, gn_stage_weights gained the six tf_* weight streams;
                                 4: + tf_pagg_w (fused pairwise aggregation, csrc/gn_pair_agg_tf32.cu), gn_fish_* entry points;
                                 5: + gn_decoder_fwd_tc (bf16 tensor-core decoder, csrc/gn_decoder_tc.cu);
                                 6: gn_decoder_tc_weights gained mlp_stream / mlp_bias (fused MLP kernel) */

#define GN_MAX_AGENTS 64      /* N <= 64: one 64-bit membership word per hyperedge */
#define GN_MAX_SCALES 8
#define GN_ATT_DIM 64         /* hdim_extend, MS_HGNN_batch.py:72,:292 */
#define GN_ATT_HIDDEN 32      /* attention_mlp hidden, :80,:301 */
#define GN_NODE_HIDDEN 256    /* node2edge_start_mlp hidden, :84,:306 */
#define GN_MLP_HIDDEN 128     /* every other hidden size, :75,:77,:254 */
#define GN_SMALL_OUT 16       /* padded width of [distribution logits | factor logit] */

enum gn_error {
  GN_OK = 0,
  GN_E_NULL = -1,        /* required pointer is NULL */
  GN_E_SHAPE = -2,       /* unsupported or inconsistent shape */
  GN_E_SCALE = -3,       /* scale > N: the reference raises "selected index k out of range" (:382) */
  GN_E_WORKSPACE = -4,   /* workspace too small */
  GN_E_PRECISION = -5,   /* unknown precision / path not built */
  GN_E_ALIGN = -6        /* pointer not 16-byte aligned */
};

enum gn_precision {
  GN_FP32 = 0,           /* fp32 FFMA everywhere: the 1e-5 parity path */
  GN_BF16_TC = 1,        /* bf16 tcgen05/TMEM GEMMs for the per-edge MLP chain, fp32 accumulate (2e-2) */
  GN_TF32X3 = 2          /* fp32-grade tensor-core path: every Linear as three kind::tf32 tcgen05 MMAs on hi/lo
                            split operands, fp32 accumulate and fp32 epilogues (1e-5, like GN_FP32); shapes the
                            tensor-core chains do not cover run the GN_FP32 kernels */
};

enum gn_noise_mode {
  GN_NOISE_GIVEN = 0,    /* U (B,E,T) uniform[0,1) supplied by the caller (reference RNG order) */
  GN_NOISE_PHILOX = 1,   /* Philox4x32-10 on device, keyed by (seed, stage, global element index) */
  GN_NOISE_PHILOX_DEVICE_SEED = 2  /* same generator; the seed is not cfg->seed but ONE uint64 in device memory that
                                      U points at, read when the kernel runs: a captured CUDA graph can then be
                                      replayed with a seed the device itself advances between replays */
};

typedef void* gn_stream_t;

/* One message-passing stage = node2edge -> MLP_dict_softmax -> edge2node -> MLP
 * (MS_HGNN_batch.py:173-195 / :425-441).  nmp_layers = L is L stages chained by
 * the host.  All weights are fp32 device pointers, pre-arranged by
 * groupnet_b200/packing.py ("t" = transposed to [K][N_out], K-major; K padded
 * to a multiple of 16 and N_out to a multiple of 64 with zeros). */
typedef struct gn_stage_weights {
  /* node2edge_start_mlp[i]  (:84,:125)   D -> 256 -> 64 */
  const float* node_w0t;  /* [Dp][256] */
  const float* node_b0;   /* [256] */
  const float* node_w1t;  /* [256][64] */
  const float* node_b1;   /* [64] */
  /* attention_mlp[i]  (:80,:134)   [x_n ; edge_init_e] 128 -> 32 -> 1, first layer split */
  const float* att_wpqt;  /* [64][64]: cols 0..31 = W0[:, :64]^T (node half), 32..63 = W0[:, 64:]^T (edge half) */
  const float* att_b0;    /* [32] */
  const float* att_w1;    /* [32] */
  const float* att_b1;    /* [1] */
  /* MLP_dict_softmax  (:31-53) */
  const float* init_w0t;  /* [64][128] */
  const float* init_b0;   /* [128] */
  const float* init_w1t;  /* [128][64] */
  const float* init_b1;   /* [64] */
  const float* df_w0t;    /* [64][256]: cols 0..127 MLP_distribution.layers.0, 128..255 MLP_factor.layers.0 */
  const float* df_b0;     /* [256] */
  const float* df_w1;     /* [256][16]: rows <128 -> cols 0..T-1 (distribution), rows >=128 -> col T (factor) */
  const float* df_b1;     /* [16] */
  /* edge_aggregation.agg_mlp[t]  (:247-268)   D -> 128 -> D, T of them */
  const float* agg_w0t;   /* [Dp][T*128] */
  const float* agg_b0;    /* [T*128] */
  const float* agg_w1t;   /* [T*128][Dc], Dc = D rounded up to 64 */
  const float* agg_b1;    /* [T][D] */
  /* nmp_mlps[2l] or nmp_mlp_end  (:77,:195)   2D -> 128 -> Dout */
  const float* post_w0t;  /* [(2D)p][128]: rows 0..D-1 act on the aggregated half, D..2D-1 on the skip half */
  const float* post_b0;   /* [128] */
  const float* post_w1t;  /* [128][Doutc] */
  const float* post_b1;   /* [Dout] */
  /* bf16 copies for the tcgen05 path (required when precision == GN_BF16_TC, else may be NULL):
   * nn.Linear.weight [N][K] in the canonical K-major no-swizzle UMMA operand layout
   * byte(n,k) = (k/8)*(N*16) + n*16 + (k%8)*2   (csrc/gn_tc.cuh) */
  const void* tc_init_w0; /* N=128, K=64   init_MLP.layers.0 */
  const void* tc_init_w1; /* N=64,  K=128  init_MLP.layers.1 */
  const void* tc_df_w0;   /* N=256, K=64   [MLP_distribution | MLP_factor].layers.0 */
  const void* tc_df_w1;   /* N=16,  K=256  rows < T: distribution head on k < 128; row T: factor head on k >= 128 */
  const void* tc_node_w0; /* N=256,   K=D      node2edge_start_mlp.layers.0 */
  const void* tc_node_w1; /* N=64,    K=256    node2edge_start_mlp.layers.1 */
  const void* tc_att_wpq; /* N=64,    K=64     rows 0..31 = attention W0[:, :64], rows 32..63 = W0[:, 64:] */
  const void* tc_agg_w0;  /* N=T*128, K=D      agg_mlp[t].layers.0 stacked over t */
  const void* tc_agg_w1;  /* N=D,     K=T*128  agg_mlp[t].layers.1 concatenated along K */
  const void* tc_post_w0; /* N=128,   K=2D     closing MLP layers.0 */
  const void* tc_post_w1; /* N=Dout,  K=128    closing MLP layers.1 */
  /* weight stream of the fused wide hyper aggregation (csrc/gn_hyper_fused_tc.cu; D == 256 only, else
   * NULL): the agg_mlp chunks in the order the kernel consumes them, each a canonical operand:
   * for s = 0..T: [s < T: W0_s[:, 0:128] (128 x 128), W0_s[:, 128:256] | b0 (hi,lo) (128 x 144)]
   *               [s >= 1, t = s-1: W1_t[:, 0:64] | b1_t (hi,lo,hi) (256 x 80), W1_t[:, 64:128] (256 x 64)]
   * h_dim == 64 hyper layers with Dout in {32, 64} (csrc/gn_hyper_fused64_tc.cu) use the same field:
   * for s = 0..T: [s < T: W0_s | b0 (128 x 80)] [s >= 1: W1_{s-1} | b1 (64 x 144)], then the closing MLP
   * [W0 | b0] (128 x 144), W1 (Dout x 128), b1 block (Dout x 16). */
  const void* tc_hfuse_w;
  /* weight stream of the fused wide node prologue (csrc/gn_node_pre256_tc.cu; D == 256 only, else NULL):
   * node W0[:, 64c:64c+64] (256 x 64) for c = 0..3, b0 (hi,lo) block (256 x 16), [W1 | b1 (hi,lo)] (64 x 272),
   * [Wp ; Wq] (64 x 64) */
  const void* tc_npre_w;
  /* weight streams of the 3xTF32 chains (csrc/gn_chain_tf32.cu; required when precision == GN_TF32X3 for the
   * chains whose shape fits, else may be NULL).  A stream is the Linears of one chain in consumption order; each
   * nn.Linear.weight [N][K] is cut along K into chunks of kc = the largest multiple of 8 that divides K with
   * N*kc*8 <= 16384, and a chunk is its tf32 "hi" copy followed by its "lo" copy (w = hi + lo, hi = w rounded
   * to 11 significand bits), each in the canonical K-major no-swizzle layout with 32-bit elements
   * byte(n,k) = (k/4)*(N*16) + n*16 + (k%4)*4   (csrc/gn_tf32.cuh) */
  const void* tf_chain_w;  /* init_MLP.0 (128x64), init_MLP.1 (64x128), MLP_factor.0 (128x64), MLP_distribution.0
                              (128x64); then plain fp32: MLP_factor.1 (128 floats), MLP_distribution.1 k-major as
                              [128][8] (T <= 8), [128][12] (T <= 12) or [128][16], rows of T logits zero padded (both heads are fp32 dot
                              products inside the drains) */
  const void* tf_pre_w;    /* node W0[0:128] (128xD), W1[:,0:128] (64x128), W0[128:256], W1[:,128:256], [Wp;Wq] (64x64) */
  const void* tf_aggin_w;  /* agg_mlp[t].layers.0 (128xD) for t < T */
  const void* tf_aggout_w; /* cat_t agg_mlp[t].layers.1 along K, (D x T*128), in K blocks of 64 */
  const void* tf_hagg_w;   /* per t: agg_mlp[t].layers.0 (128xD), agg_mlp[t].layers.1 (Dx128) */
  const void* tf_post_w;   /* closing MLP layers.0 (128x2D), layers.1 (Doutx128) */
  /* fused pairwise aggregation (csrc/gn_pair_agg_tf32.cu; pairwise layers with D == 64, else NULL): 32 KB chunks
   * (64 x 64, hi then lo) in MMA issue order.  With unit step u = 2t + half, A(u) = agg_mlp[t].layers.0.weight
   * rows [64 half, 64 half + 64) (64 x D) and B(u) = agg_mlp[t].layers.1.weight columns [64 half, 64 half + 64)
   * (D x 64):  A(0), A(1), A(2), then for st = 0..2T-1: B(st-1) (st >= 1), A(st+3) (st + 3 < 2T); finally B(2T-1). */
  const void* tf_pagg_w;
} gn_stage_weights;

typedef struct gn_stage_cfg {
  int32_t B;            /* scenes in this call */
  int32_t N;            /* agents per scene, 1..GN_MAX_AGENTS */
  int32_t D;            /* h_dim, multiple of 4 */
  int32_t Dout;         /* output width of the stage's closing MLP */
  int32_t E;            /* edges per scene: N*N (pairwise), N or 1 (hyper) */
  int32_t T;            /* edge types: 6 pairwise (:74), 10 hyper (:294); <= 15 */
  int32_t pairwise;     /* 1: ordered pairs incl. self loops, incidence implicit (:143-160); 0: H given */
  int32_t precision;    /* enum gn_precision */
  int32_t noise_mode;   /* enum gn_noise_mode */
  int32_t stage_index;  /* 0..L-1, separates Philox streams of chained stages */
  int32_t out_ld;       /* row stride (floats) of node_out; 0 = Dout.  Lets the stage write its slice of the
                           concatenated feature tensor the encoders build (model/GroupNet_nba.py:301-309) */
  int32_t h_stride;     /* floats between consecutive scenes of H; 0 = E*N.  Lets a hyper stage read its
                           rows of the concatenated (B, sum E, N) incidence (model/GroupNet_nba.py:296,299) */
  uint64_t seed;        /* Philox key */
  int64_t scene_offset; /* global index of scene 0 of this call: results do not depend on sharding */
} gn_stage_cfg;

/* ABI version of the loaded library (== GN_ABI_VERSION). */
int gn_abi_version(void);

/* Static string for a return code (GN_E_* or cudaError_t). */
const char* gn_error_string(int code);

/* Fused feature correlation + per-agent top-k hyperedge selection for every
 * scale + incidence emission.  Replaces F.normalize + matmul
 * (model/GroupNet_nba.py:284-286) and init_adj_attention
 * (model/MS_HGNN_batch.py:372-388) for S scales in one pass over x.
 *   x            (B,N,D)
 *   scales[s]    reference `scale` argument; <1 is clamped to 1, ==N gives the
 *                single all-ones hyperedge, >N returns GN_E_SCALE
 *   H_out[s]     element (b,e,n) is written at H_out[s][b*H_scene_stride[s] + e*N + n],
 *                e < (scales[s]==N ? 1 : N); a stride larger than E*N lets the
 *                caller write straight into the concatenated (B,sum E,N) tensor
 *                the encoders build (model/GroupNet_nba.py:296,299)
 *   corr_out     optional (B,N,N), may be NULL
 * Ties: among equal correlations the lower agent index is selected. */
int gn_corr_topk_h(const float* x, int32_t B, int32_t N, int32_t D,
                   const int32_t* scales, int32_t S,
                   float* const* H_out, const int64_t* H_scene_stride,
                   float* corr_out, gn_stream_t stream);

/* Top-k incidence from a caller-supplied correlation tensor: exactly
 * init_adj_attention (model/MS_HGNN_batch.py:372-388).  corr (B,N,N). */
int gn_topk_h(const float* corr, int32_t B, int32_t N, int32_t scale,
              float* H_out, int64_t H_scene_stride, gn_stream_t stream);

/* Bytes of device workspace gn_stage_fwd needs for this configuration. */
size_t gn_stage_workspace_bytes(const gn_stage_cfg* cfg);

/* One message-passing stage forward.
 *   h_in      (B,N,D)    node features entering the stage
 *   H         (B,E,N)    incidence (0/1 fp32) for hyper stages; ignored (may be NULL) when cfg->pairwise
 *   U         (B,E,T)    uniform draws when noise_mode == GN_NOISE_GIVEN; one device uint64 (the Philox seed,
 *                        8-byte aligned) when GN_NOISE_PHILOX_DEVICE_SEED; may be NULL for GN_NOISE_PHILOX
 *   node_out  (B,N,Dout) output of the closing MLP (node_feat); rows cfg->out_ld floats apart when out_ld != 0
 *   dist_out  (B,E,T)    optional: the categorical `distribution` the reference returns as
 *                        `factors` (:53,:178,:427); may be NULL for stages > 0
 * Replaces node2edge (:122-141/:357-370), MLP_dict_softmax.forward (:41-53),
 * gumbel_softmax (:446-520), edge2node/edge_aggregation (:116-120,:259-268) and
 * the closing MLP (:195,:441). */
int gn_stage_fwd(const gn_stage_cfg* cfg, const gn_stage_weights* w,
                 const float* h_in, const float* H, const float* U,
                 float* node_out, float* dist_out,
                 void* workspace, size_t workspace_bytes, gn_stream_t stream);

/* Number of kernel launches the last-described configuration issues per
 * gn_stage_fwd call (for bench.py's gpu_launches accounting). */
int gn_stage_launch_count(const gn_stage_cfg* cfg);

/* PastEncoder front-end (model/GroupNet_nba.py:269-280), eval mode: the chain input_fc ->
 * PositionalAgentEncoding (concat + fc) -> input_fc2 -> add_category -> input_fc3 is affine, folded by the
 * host into Mt (K x C, K = past_length*in_dim) and a per-agent bias table (N x C):
 *   out[r,:] = Mt^T inputs[r,:] + bias_agent[r % N,:]        inputs (R,K), out (R,C), R = B*N */
int gn_past_frontend(const float* inputs, int64_t R, int32_t K, int32_t N, int32_t C,
                     const float* Mt, const float* bias_agent, float* out, gn_stream_t stream);

/* ---- trajectory decoder, fp32 ---------------------------------------
 * Replaces Decoder.forward (model/GroupNet_nba.py:461-505) and DecomposeBlock.forward (:48-79):
 * per row (scene-agent x sample) and block  res = x_true - x_hat -> conv1d(2->32,k3,pad1)+ReLU ->
 * GRU(32->96) last state -> [past_feature ; z ; state] -> decoder_x / decoder_y (MLPs -> 512 -> 256 -> out);
 * reconstruction = sum x_hat, out_seq = sum y_hat + cur_location.
 * One gn_decoder_weights per DecomposeBlock, packed by groupnet_b200/packing.py::pack_decoder_block
 * (K-major, column-permuted Linear weights as in gn_stage_weights; the three GRU gates r|z|n each padded
 * from 96 to 128 columns). */
typedef struct gn_decoder_weights {
  const float* conv_w;   /* conv_past.weight (32,2,3) as stored */
  const float* conv_b;   /* (32) */
  const float* gru_wx;   /* packed (32, 384): weight_ih_l0^T, gates r|z|n */
  const float* gru_wh;   /* packed (96, 384): weight_hh_l0^T */
  const float* gru_b;    /* (4,128): b_ir+b_hr | b_iz+b_hz | b_in | b_hn, zero padded */
  const float* x_w0; const float* x_b0; const float* x_w1; const float* x_b1; const float* x_w2; const float* x_b2;
  const float* y_w0; const float* y_b0; const float* y_w1; const float* y_b1; const float* y_w2; const float* y_b2;
                         /* decoder_x / decoder_y: (Kp,512) (512) (512,256) (256) (256,64) (64), Kp = round16(F+Z+96) */
} gn_decoder_weights;

/* Bytes of device scratch gn_decoder_fwd needs (the x_hat carried between blocks). */
size_t gn_decoder_workspace_bytes(int64_t A, int32_t S, int32_t Tp);

/*   blocks        host array of num_blocks structs of device pointers (args.num_decompose)
 *   past_feature  (A*S, F)   per-row past features (repeat_interleave'd by the caller as the reference does)
 *   z             (A*S, Z)   latent samples
 *   past_traj     (A, Tp, 2) x_true;  cur_location (A, 1, 2);  A = batch * agents, S = sample_num
 *   out_seq       (A*S, Tf, 2) == (A, S, Tf, 2) for mode='inference';  recover (A*S, Tp, 2)
 * Limits: F % 4 == 0, Z % 4 == 0, Tp <= 32, Tf <= 32, (F + Z) such that the tile fits 227 KB of shared memory
 * (F + Z <= 444). */
int gn_decoder_fwd(const gn_decoder_weights* blocks, int32_t num_blocks, const float* past_feature, const float* z,
                   const float* past_traj, const float* cur_location, int64_t A, int32_t S, int32_t F, int32_t Z,
                   int32_t Tp, int32_t Tf, float* out_seq, float* recover, void* workspace, size_t workspace_bytes,
                   gn_stream_t stream);

/* ---- trajectory decoder, bf16 tensor-core path (tcgen05, 2e-2 parity) --------------------------
 * The same contract as gn_decoder_fwd (Decoder.forward, model/GroupNet_nba.py:461-505; DecomposeBlock.forward
 * :48-79) with bf16 operands and fp32 accumulation: the GRU step as one K = 128 contraction per time step with
 * the gate math in the accumulator drain (fp32 state), decoder_x / decoder_y fused in one kernel with the hidden
 * activations on chip (row-tile GEMMs for feature widths above 384 or not a multiple of 64, or more than 16 time steps).
 * One gn_decoder_tc_weights per DecomposeBlock, packed by groupnet_b200/packing.py::pack_decoder_block_tc;
 * "canonical" = the K-major no-swizzle UMMA operand layout [K/8][N][8] of bf16 (csrc/gn_tc.cuh). */
typedef struct gn_decoder_tc_weights {
  const float* conv_w;   /* conv_past.weight (32,2,3) as stored */
  const float* conv_b;   /* (32) */
  const void* gru_w;     /* bf16, two canonical [192 x 128] operands: rows (r | z), then (n_x | n_h); K = e (32) | h (96),
                            zero blocks where a gate half does not see e or h */
  const float* gru_b;    /* (4,128): b_ir+b_hr | b_iz+b_hz | b_in | b_hn, zero padded */
  const void* w0;        /* bf16 canonical [(F+Z+96)/8][1024][8]: rows 0..511 decoder_x layer 0, 512..1023 decoder_y */
  const float* b0;       /* (1024) */
  const void* x_w1; const float* x_b1;   /* bf16 canonical [64][256][8], (256) */
  const void* x_w2; const float* x_b2;   /* bf16 canonical [32][P][8], (P): P = 2*Tp rounded up to 16, zero rows */
  const void* y_w1; const float* y_b1;
  const void* y_w2; const float* y_b2;   /* P = 2*Tf rounded up to 16 */
  const void* mlp_stream;  /* bf16: both MLPs' weights as 16 KB stages in the fused kernel's consumption order
                              (packing.py::decoder_mlp_stream); required when F + Z + 96 <= 384 is a multiple of 64 and
                              Tp, Tf <= 16 (the fused kernel's shapes), NULL otherwise (the row-tile GEMMs run) */
  const float* mlp_bias;   /* b0 (1024) | x_b1 (256) | y_b1 (256) | x_b2 (32) | y_b2 (32), zero padded */
} gn_decoder_tc_weights;

/* Bytes of device scratch gn_decoder_fwd_tc needs: x_hat, the bf16 feature rows, the last Linears' outputs and — only at
 * shapes the fused MLP kernel does not cover — the hidden activations (1.1 KB against 4.2 KB per row at the NBA shape). */
size_t gn_decoder_tc_workspace_bytes(int64_t A, int32_t S, int32_t F, int32_t Z, int32_t Tp, int32_t Tf);

/* Arguments as gn_decoder_fwd.  Limits: F % 8 == 0, Z % 8 == 0, (F + Z) % 16 == 0, Tp <= 32, Tf <= 32. */
int gn_decoder_fwd_tc(const gn_decoder_tc_weights* blocks, int32_t num_blocks, const float* past_feature,
                      const float* z, const float* past_traj, const float* cur_location, int64_t A, int32_t S,
                      int32_t F, int32_t Z, int32_t Tp, int32_t Tf, float* out_seq, float* recover, void* workspace,
                      size_t workspace_bytes, gn_stream_t stream);

/* ---- training: backward of one stage (fp32) -------------------------------------------------
 * Gradients flow to h_in and to every parameter the forward uses; H, corr and the noise get none
 * (model/MS_HGNN_batch.py:382 uses top-k indices only).  Parameters are read in their native
 * nn.Linear layout (weight (N_out, K) row-major, bias (N_out)); dW / db are ACCUMULATED into the
 * caller's buffers (zero them for a fresh gradient).  A NULL dW skips that parameter. */
typedef struct gn_lin {
  const float* W;   /* (N, K) */
  const float* b;   /* (N) or NULL */
  float* dW;        /* (N, K), accumulated; may be NULL */
  float* db;        /* (N), accumulated; may be NULL */
  int32_t N, K;
} gn_lin;

typedef struct gn_train_params {
  gn_lin node0, node1;            /* node2edge_start_mlp[i].layers.{0,1} */
  gn_lin attpq;                   /* (64, 64): rows 0..31 = attention W0[:, :64], rows 32..63 = W0[:, 64:]; no bias */
  const float* att_b0;            /* attention_mlp[i].layers.0.bias (32) */
  const float* att_w1;            /* attention_mlp[i].layers.1.weight (32) */
  const float* att_b1;            /* attention_mlp[i].layers.1.bias (1) */
  float* d_att_b0; float* d_att_w1; float* d_att_b1;   /* accumulated */
  gn_lin init0, init1, dist0, dist1, fac0, fac1;       /* MLP_dict_softmax */
  gn_lin agg0[15], agg1[15];      /* edge_aggregation.agg_mlp[t].layers.{0,1}, t < T */
  gn_lin post0, post1;            /* closing MLP */
} gn_train_params;

/* Byte offsets, inside the workspace gn_stage_fwd (precision GN_FP32) was given, of the tensors the
 * backward needs: out5 = { x' (B*N,64), pq (B*N,64), edges (B*E,64), edge_feat (B*E,T), agg (B*N,D) }.
 * Keep that workspace alive until gn_stage_bwd has run. */
int gn_stage_saved_offsets(const gn_stage_cfg* cfg, size_t* out5);

/* Scratch bytes gn_stage_bwd needs. */
size_t gn_stage_bwd_workspace_bytes(const gn_stage_cfg* cfg);

/* Backward of gn_stage_fwd (cfg->precision must be GN_FP32).
 *   fwd_workspace  the forward's workspace (read only)
 *   d_node_out     (B,N,Dout) gradient of node_out, rows ld_dout floats apart
 *   d_dist         optional (B,E,T) gradient of dist_out, may be NULL
 *   d_h            (B,N,D) gradient of h_in (overwritten) */
int gn_stage_bwd(const gn_stage_cfg* cfg, const gn_train_params* params,
                 const float* h_in, const float* H, const void* fwd_workspace,
                 const float* d_node_out, int64_t ld_dout, const float* d_dist, float* d_h,
                 void* workspace, size_t workspace_bytes, gn_stream_t stream);

/* Profiling hook (the only process-wide state in the library; off by default).
 * While enabled every kernel launch is bracketed by CUDA events on its launch
 * stream.  gn_profile_collect synchronises those events, sums the durations by
 * kernel name into total_ms[i] / counts[i], writes the names ';'-separated into
 * `names`, clears the records and returns the number of distinct kernels. */
void gn_profile_enable(int on);
int gn_profile_collect(char* names, int names_len, float* total_ms, int* counts, int max_entries);
/* Optional device buffer (>= 2*8*16 uint64) that the edge-chain kernel fills with clock64() phase stamps of
 * block 0's first tiles; NULL (default) disables tracing. */
void gn_profile_set_trace(unsigned long long* device_buffer);

/* ---- group-wise operators of the fish model (csrc/gn_fish.cu, fp32, eval-mode semantics) ----
 * All tensors fp32, contiguous, batch-major; rel_rec / rel_send are (E, N) per scene, `rel_stride` floats apart
 * (0: one matrix shared by every scene, as the reference's expand() does).  BatchNorm layers are folded into the
 * Linears by the host (running statistics). */

/* compute_alpha_im (model/encoder.py:261-303): alpha_ij (B,E), I_HG (B,N,M) -> out (B,N,M). */
int gn_fish_alpha_im(const float* alpha_ij, const float* I_HG, const float* rel_rec, const float* rel_send,
                     int64_t rel_stride, int32_t B, int32_t E, int32_t N, int32_t M, float* out, gn_stream_t stream);

/* Per-scene C = A^T B with optional column normalisation of A (A / (column sum + 1e-8), MLPHGE :236-241) and optional
 * per-row weights: A (R,Cn) [A_stride floats between scenes, 0 = shared], Bm (R,F), roww (batch,R) or NULL -> C (Cn,F).
 * einsum('bnm,bnf->bmf') (model/encoder.py:241,183), ('bmn,bmf->bnf') (:177), ('behd,ben->bnhd') (:454). */
int gn_fish_bmm_t(const float* A, int64_t A_stride, const float* Bm, const float* roww, int32_t batch, int32_t R,
                  int32_t Cn, int32_t F, int32_t norm_cols, float* C, gn_stream_t stream);

/* Row MLP: nlayers <= 3 Linears, Wt[l] = weight^T as [K][N] (BatchNorm folded in), bias[l] [N] or NULL, activation
 * act[l] in {0 none, 1 LeakyReLU(slope[l]), 2 ELU} after every layer; x (R,K0) rows ldx apart -> out (R,N_last) rows
 * ldo apart.  The Linear / BatchNorm1d / activation stacks of model/encoder.py:125-139, :244-248, :364-381. */
int gn_fish_mlp(const float* x, int64_t ldx, int64_t R, int32_t K0, int32_t nlayers, const float* const* Wt,
                const float* const* bias, const int32_t* N, const int32_t* act, const float* slope,
                float* out, int64_t ldo, gn_stream_t stream);

/* HyperEdgeAttention core (model/encoder.py:160-177): e_proj (B,M,Hd), v_proj (B,N,Hd), attention vector (2 Hd),
 * I_HG (B,N,M), e_HG (B,M,F) -> v1 (B,N,F) = softmax over the member nodes of leaky(logit) / 100, times e_HG. */
int gn_fish_hga_core(const float* e_proj, const float* v_proj, const float* attention_vector, const float* I_HG,
                     const float* e_HG, int32_t B, int32_t N, int32_t M, int32_t Hd, int32_t F, float slope, float* v1,
                     gn_stream_t stream);

/* TemporalGATLayer edge stage (model/encoder.py:404-447): v_proj (B,N,H*D) -> edge_input (B,E,H,2D), alpha_ij (B,E,H). */
int gn_fish_gat_edges(const float* v_proj, const float* rel_rec, const float* rel_send, int64_t rel_stride,
                      const float* a_forward, const float* a_backward, int32_t B, int32_t E, int32_t N, int32_t H,
                      int32_t D, float slope, float* edge_input, float* alpha_ij, gn_stream_t stream);

/* build_dynamic_graph_and_hypergraph (utilities/utils.py:191-244): z_CG (B,E,Lc), z_HG (B,M,Lh) -> edges / hyperedges
 * whose argmax type is 0 are zeroed in new_rel_rec / new_rel_send (B,E,N) and new_I_HG (B,N,M); types as int64. */
int gn_fish_dynamic_graph(const float* z_CG, const float* z_HG, const float* rel_rec, const float* rel_send,
                          int64_t rel_stride, const float* I_HG, int32_t B, int32_t E, int32_t N, int32_t M,
                          int32_t Lc, int32_t Lh, float* new_rel_rec, float* new_rel_send, float* new_I_HG,
                          int64_t* edge_types, int64_t* hyperedge_types, gn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GROUPNET_B200_H */
