// Program format of the 3xTF32 chain engine (gn_chain_tf32.cu): a tile of 128 rows runs a fixed list of
// GEMM "ops" whose activations never leave the SM.  See gn_chain_tf32.cu for the execution model.
#pragma once
#include "gn_common.cuh"

namespace gn { namespace tfe {

constexpr int MAX_OPS = 64;
constexpr int MAX_EV = 72;
// weight-ring stage (a [N x kc] chunk, hi then lo): the largest of 64 / 32 / 16 KB that leaves room for two stages
// beside the staged A buffers; must match packing.tf_stage_bytes
constexpr uint32_t SMEM_BUDGET = 227 * 1024;
constexpr int NBAR = 4;                        // depth of the a_ready / acc_ready mbarrier rings
// ST_PAIR node block: the rows a tile may span, pq rows padded to NLD floats, Y rows to YLD (rows 4 banks apart)
constexpr int MAXN = 36, NLD = 68, YLD = 132;
__host__ __device__ constexpr uint32_t node_block_bytes() { return static_cast<uint32_t>(MAXN) * (NLD + YLD) * 4; }
constexpr int NSLICE = 4;                      // threads per tile row (column slices); 16 row warps
constexpr uint32_t SCR_BYTES = (NSLICE * 128 * 2 + NSLICE * 128 + 128 * 17) * 4;   // logit partials | dot partials | y
constexpr int AUX_FLOATS = 3072;               // small constants (biases, attention tail, factor head) kept in smem:
                                               // with ~227 KB of shared memory carved out there is no L1 left
constexpr int MAX_AUX = 12;
constexpr uint32_t FIXED_BYTES = SCR_BYTES + AUX_FLOATS * 4 + 1024;                // + barriers + alignment slack
inline uint32_t ring_stage_bytes(int a0_K, int nbuf, bool node_block) {
  const uint32_t used = static_cast<uint32_t>(nbuf) * a0_K * 1024u + (node_block ? node_block_bytes() : 0u) + FIXED_BYTES;
  const uint32_t avail = used < SMEM_BUDGET ? SMEM_BUDGET - used : 0u;
  return avail >= 2 * 65536u ? 65536u : (avail >= 2 * 32768u ? 32768u : 16384u);
}

enum { EV_STAGE = 0, EV_DRAIN = 1, EV_STAGE_NEXT = 2, EV_PAIR_A_NEXT = 3, EV_PAIR_B_NEXT = 4 };   // STAGE_NEXT: ST_ROWS staging of tile i + 1 inside tile i
// trace layout: tile t (< TR_TILES) owns TR_SLOTS stamps: [0, 3*MAX_OPS) issuer (per op: operands ready, first weight
// chunk landed, all MMAs issued); [TR_ROWS, TR_ROWS + 3*MAX_EV) row thread 0 (per event: start, wait done, end)
constexpr int TR_TILES = 6, TR_ROWS = 3 * MAX_OPS, TR_CHUNK = TR_ROWS + 3 * MAX_EV, TR_MAXCH = 40;
// [TR_CHUNK, +4*TR_MAXCH): per weight chunk: producer issued the copy | issuer: wait start, wait end, MMAs issued
constexpr int TR_STAGE = TR_CHUNK + 4 * TR_MAXCH;      // 8 sub-stamps inside the (pairwise) staging pass
constexpr int TR_SLOTS = TR_STAGE + 8;
enum { DR_NONE = 0, DR_TMEM = 1, DR_STORE = 2, DR_TMEM_STORE = 3, DR_DOT = 4, DR_DOTG = 5 };
enum { ST_ROWS = 0, ST_PAIR = 1 };
enum { A_SMEM = 0, A_TMEM = 1 };

struct Op {
  // ---- MMA: acc[128 x N] (+)= A[128 x K] * W[N x K]^T, weights streamed as K / kc chunks
  short a_src;        // A_SMEM: staged buffer a_buf; A_TMEM: columns a_col (hi) and a_col + K (lo)
  short a_buf, a_col;
  short K, N, kc;
  short acc_col;
  short drain_col;    // accumulator columns the drain reads (acc_col unless the op signals for an EARLIER op's accumulator)
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
  short bias_off;     // >= 0: bias[dn] at this float offset of the smem constants
  short variant;      // compiled drain specialisation (set by validate_program)
  int w_off;          // byte offset of the op's first weight chunk in the stream
  const float* bias;  // [dn] or null
  float* out;         // DR_STORE / DR_TMEM_STORE: row-major fp32, row stride ldo
  long long ldo;
};

struct Args {
  Op ops[MAX_OPS];
  int nops;
  unsigned char ev_type[MAX_EV], ev_op[MAX_EV];
  int nev;
  int stage_first;                 // >= 0: the program stages with EV_STAGE_NEXT; this op's buffer is staged for a CTA's first tile
  int pro_op, skip_last_op;        // software-pipelined pairwise edge chain: op issued once before a CTA's first tile
                                   // (its operand staged by pair_stage_a/b there) and skipped on its last tile; -1: none
  const unsigned char* wstream;    // the tile's weight chunks in op / chunk order (same for every tile)
  long long R;                     // rows
  long long ntiles;
  int stage_mode;
  // ST_ROWS: logical row = [src0 (k_src0 columns) | src1], divided by a_div when != 0
  const float* src0; long long ld0; int k_src0;
  const float* src1; long long ld1;
  float a_div;
  // ST_PAIR: fused pairwise node2edge (model/MS_HGNN_batch.py:122-141) + the first Linear of init_MLP, which commutes
  // with the weighted gather: relu(w_i Y_i + w_j Y_j + b), Y = x' W^T per NODE (ypre, 128 columns), written straight
  // into tensor memory as the A operand of the chain's first op
  const float* ypre; const float* pq;
  int att_off, yb_off;             // smem constants: attention tail b0[32] | w1[32] | b1[4]; bias of the commuted Linear
  // smem constants, copied once per CTA: aux_n[i] floats from aux_src[i] to float offset aux_off[i]
  const float* aux_src[MAX_AUX]; short aux_n[MAX_AUX], aux_off[MAX_AUX]; int naux;
  int N, E, tps;
  // per-row scales (edge_feat or S) and the rank-T bias of the aggregation output
  const float* rs; int rs_ld; int rs_n;   // rs_n leading values of every row are staged in shared memory per tile
  const float* bm; int bm_T; int bm_ld; int bm_off;   // bm_off >= 0: bm lives in the smem constants
  // DR_DOT: carry = sum_k relu(acc_k + bias_k) * dot_w[k * dot_stride]
  int dot_off;                     // smem constants
  // DR_DOTG (distribution head + MLP_dict_softmax tail, :45-53, :446-520)
  int gb_off; int T;               // smem constants: df_b1[16]
  int w4_off, w4_tp;               // smem constants: MLP_distribution.layers.1.weight as [128][w4_tp], w4_tp = 6, 8, 10, 12 or 16
  const float* U; int noise_mode; unsigned long long seed; long long scene_offset; int stage_index;
  float* dist_out; float* edge_feat;
  // shared-memory layout
  uint32_t off_a0, a0_buf_bytes;   // staged A buffers: buffer b at off_a0 + b * a0_buf_bytes, hi then lo
  uint32_t a0_half_bytes;          // bytes of one (hi or lo) copy = 128 * K_buf * 4
  uint32_t off_ring; int nstage; uint32_t stage_bytes;
  uint32_t off_node, off_scr, off_aux, off_bar;
  unsigned long long* trace;       // optional (gn_profile_set_trace): clock64 stamps of block 0's first tiles, see TR_* below
};

}}  // namespace gn::tfe
