// Program format of the 3xTF32 chain engine (gn_chain_tf32.cu): a tile of 128 rows runs a fixed list of
// GEMM "ops" whose activations never leave the SM.  See gn_chain_tf32.cu for the execution model.
#pragma once
#include "gn_common.cuh"

namespace gn { namespace tfe {

constexpr int MAX_OPS = 32;
constexpr int MAX_EV = 72;
constexpr uint32_t STAGE_BYTES = 16384;        // one weight-ring stage: a [N x kc] chunk, hi then lo
constexpr int NBAR = 4;                        // depth of the a_ready / acc_ready mbarrier rings
constexpr int MAXN = 36, NLD = 68;             // ST_PAIR node block: rows a tile may span, padded row (floats)

enum { EV_STAGE = 0, EV_DRAIN = 1 };
enum { DR_NONE = 0, DR_TMEM = 1, DR_STORE = 2, DR_TMEM_STORE = 3, DR_DOT = 4, DR_GUMBEL = 5 };
enum { ST_ROWS = 0, ST_PAIR = 1 };
enum { A_SMEM = 0, A_TMEM = 1 };

struct Op {
  // ---- MMA: acc[128 x N] (+)= A[128 x K] * W[N x K]^T, weights streamed as K / kc chunks
  short a_src;        // A_SMEM: staged buffer a_buf; A_TMEM: columns a_col (hi) and a_col + K (lo)
  short a_buf, a_col;
  short K, N, kc;
  short acc_col;
  short accumulate;   // first MMA adds onto the accumulator's content
  short wait_n;       // a_ready phases the issuer consumes before issuing
  short signal;       // commit acc_ready afterwards (a DRAIN event consumes it)
  // ---- staging (EV_STAGE): columns [st_k0, st_k0 + K) of the logical input row -> smem buffer a_buf
  short st_k0;
  // ---- drain (EV_DRAIN): accumulator columns [acc_col, acc_col + dn)
  short drain, dn, relu;
  short dst_col;      // DR_TMEM*: hi at dst_col, lo at dst_col + dn
  short arrive;       // row threads arrive on a_ready after this drain
  short rs_idx;       // >= 0: v *= rs[row * rs_ld + rs_idx]
  short use_bm;       // DR_STORE: v += sum_t rs[row * rs_ld + t] * bm[t * bm_ld + col]
  short out_col0;
  short nsum, sum_stride;  // drain value = sum of nsum accumulators, sum_stride columns apart (0/1 = just acc_col)
  const float* bias;  // [dn] or null
  float* out;         // DR_STORE / DR_TMEM_STORE: row-major fp32, row stride ldo
  long long ldo;
};

struct Args {
  Op ops[MAX_OPS];
  int nops;
  unsigned char ev_type[MAX_EV], ev_op[MAX_EV];
  int nev;
  const unsigned char* wstream;    // the tile's weight chunks in op / chunk order (same for every tile)
  long long R;                     // rows
  long long ntiles;
  int stage_mode;
  // ST_ROWS: logical row = [src0 (k_src0 columns) | src1], divided by a_div when != 0
  const float* src0; long long ld0; int k_src0;
  const float* src1; long long ld1;
  float a_div;
  // ST_PAIR: fused pairwise node2edge (model/MS_HGNN_batch.py:122-141)
  const float* xprime; const float* pq;
  const float* att_b0; const float* att_w1; const float* att_b1;
  int N, E, tps;
  // per-row scales (edge_feat or S) and the rank-T bias of the aggregation output
  const float* rs; int rs_ld;
  const float* bm; int bm_T; int bm_ld;
  // DR_DOT: carry = sum_k relu(acc_k + bias_k) * dot_w[k * dot_stride]
  const float* dot_w; int dot_stride;
  // DR_GUMBEL (MLP_dict_softmax tail, :45-53, :446-520)
  const float* g_bias; int T;
  const float* U; int noise_mode; unsigned long long seed; long long scene_offset; int stage_index;
  float* dist_out; float* edge_feat;
  // shared-memory layout
  uint32_t off_a0, a0_buf_bytes;   // staged A buffers: buffer b at off_a0 + b * a0_buf_bytes, hi then lo
  uint32_t a0_half_bytes;          // bytes of one (hi or lo) copy = 128 * K_buf * 4
  uint32_t off_ring; int nstage;
  uint32_t off_node, off_bar;
};

}}  // namespace gn::tfe
