"""Weight packing for the C ABI (`struct gn_stage_weights`).

The kernels read every Linear as a K-major matrix `Wt[k][n]` (the transpose of
`nn.Linear.weight`), zero-padded in K to a multiple of 16 and in N to a
multiple of the GEMM chunk width, with the columns of each chunk permuted so a
thread's columns are contiguous in shared memory (csrc/gn_gemm_simt.cuh).
Packing is pure host-side plumbing over the module's own parameters; packed
copies are cached and rebuilt when any parameter's `_version` changes.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import _lib


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def chunk_permutation(tn: int) -> torch.Tensor:
    """src[pos] = original column (inside a TN chunk) stored at packed position
    `pos`.  Thread tx owns original columns tx + 16*j; they are stored at
    (j//4)*64 + tx*4 + j%4 (TN = 128) or tx*4 + j (TN = 64)."""
    src = torch.empty(tn, dtype=torch.long)
    rn = tn // 16
    for tx in range(16):
        for j in range(rn):
            pos = (j // 4) * 64 + tx * 4 + (j % 4)
            src[pos] = tx + 16 * j
    return src


def permute_cols(wt: torch.Tensor, tn: int) -> torch.Tensor:
    k, n = wt.shape
    assert n % tn == 0
    src = chunk_permutation(tn).to(wt.device)
    idx = (torch.arange(0, n, tn, device=wt.device)[:, None] + src[None, :]).reshape(-1)
    return wt[:, idx].contiguous()


def _kmajor(weight: torch.Tensor, kp: int, nc: int, tn: int) -> torch.Tensor:
    """nn.Linear.weight (N_out, K) -> packed K-major (kp, nc)."""
    n_out, k = weight.shape
    wt = torch.zeros(kp, nc, dtype=torch.float32, device=weight.device)
    wt[:k, :n_out] = weight.detach().t()
    return permute_cols(wt, tn)


def canonical_bf16(weight: torch.Tensor) -> torch.Tensor:
    """nn.Linear.weight (N, K) -> bf16 in the canonical K-major no-swizzle UMMA
    operand layout [K/8][N][8] (csrc/gn_tc.cuh): byte(n,k) = (k/8)*(N*16) + n*16 + (k%8)*2."""
    n, k = weight.shape
    assert k % 8 == 0 and n % 8 == 0
    return weight.detach().reshape(n, k // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16)


def _canon(m: torch.Tensor) -> torch.Tensor:
    """(R, K) fp32 -> flat bf16 in the canonical operand layout [K/8][R][8]."""
    r, k = m.shape
    return m.reshape(r, k // 8, 8).permute(1, 0, 2).contiguous().to(torch.bfloat16).reshape(-1)


def _hi_lo(v: torch.Tensor):
    hi = v.to(torch.bfloat16).to(torch.float32)
    return hi, (v - hi).to(torch.bfloat16).to(torch.float32)


def node_pre256_stream(w0, b0, w1, b1, wpq) -> torch.Tensor:
    """Weight stream of csrc/gn_node_pre256_tc.cu: w0 (256, 256) in four K chunks of 64 with its bias block
    after the last, [w1 (64, 256) | b1], wpq (64, 64) = [Wp ; Wq]."""
    dev = w0.device
    parts = [_canon(w0[:, 64 * c:64 * (c + 1)].contiguous()) for c in range(4)]
    hi, lo = _hi_lo(b0)
    blk = torch.zeros(256, 16, dtype=torch.float32, device=dev)
    blk[:, 0], blk[:, 1] = hi, lo
    parts.append(_canon(blk))
    hi, lo = _hi_lo(b1)
    blk = torch.zeros(64, 272, dtype=torch.float32, device=dev)
    blk[:, :256] = w1
    blk[:, 256], blk[:, 257] = hi, lo
    parts.append(_canon(blk))
    parts.append(_canon(wpq))
    return torch.cat(parts).contiguous()


def hyper_fused64_stream(agg_params, post_w0, post_b0, post_w1, post_b1) -> torch.Tensor:
    """Weight stream of csrc/gn_hyper_fused64_tc.cu (h_dim 64): per step s = 0..T the chunks [W0_s | b0_s]
    (128 x 80) for s < T and [W1_{s-1} | b1_{s-1}] (64 x 144) for s >= 1, then the closing MLP:
    [W0 | b0] (128 x 144), W1 (Dout x 128), b1 block (Dout x 16)."""
    t = len(agg_params)
    dev = post_w0.device
    parts = []
    for s in range(t + 1):
        if s < t:
            w0, b0 = agg_params[s][0], agg_params[s][1]               # (128, 64), (128,)
            hi, lo = _hi_lo(b0)
            blk = torch.zeros(128, 80, dtype=torch.float32, device=dev)
            blk[:, :64] = w0
            blk[:, 64], blk[:, 65] = hi, lo
            parts.append(_canon(blk))
        if s >= 1:
            w1, b1 = agg_params[s - 1][2], agg_params[s - 1][3]       # (64, 128), (64,)
            hi, lo = _hi_lo(b1)
            blk = torch.zeros(64, 144, dtype=torch.float32, device=dev)
            blk[:, :128] = w1
            blk[:, 128], blk[:, 129], blk[:, 130] = hi, lo, hi
            parts.append(_canon(blk))
    hi, lo = _hi_lo(post_b0)
    blk = torch.zeros(128, 144, dtype=torch.float32, device=dev)
    blk[:, :128] = post_w0
    blk[:, 128], blk[:, 129] = hi, lo
    parts.append(_canon(blk))
    parts.append(_canon(post_w1.contiguous()))
    hi, lo = _hi_lo(post_b1)
    blk = torch.zeros(post_w1.shape[0], 16, dtype=torch.float32, device=dev)
    blk[:, 0], blk[:, 1] = hi, lo
    parts.append(_canon(blk))
    return torch.cat(parts).contiguous()


def hyper_fused_post_stream(w0, b0, w1, b1) -> torch.Tensor:
    """Closing-MLP chunks appended to the fused kernel's stream (Dout % 32 == 0, <= 256):
    post_w0 in four K chunks of 128 (the last with its bias block), post_w1 in two K chunks of 64."""
    dev = w0.device                                                    # w0 (128, 2D), w1 (Dout, 128), fp32
    parts = [_canon(w0[:, 128 * c:128 * (c + 1)].contiguous()) for c in range(4)]
    hi, lo = _hi_lo(b0)
    blk = torch.zeros(128, 16, dtype=torch.float32, device=dev)
    blk[:, 0], blk[:, 1] = hi, lo
    parts.append(_canon(blk))
    dout = w1.shape[0]
    parts.append(_canon(w1[:, :64].contiguous()))
    hi, lo = _hi_lo(b1)
    blk = torch.zeros(dout, 16, dtype=torch.float32, device=dev)
    blk[:, 0], blk[:, 1] = hi, lo
    parts.append(_canon(blk))
    parts.append(_canon(w1[:, 64:].contiguous()))
    return torch.cat(parts).contiguous()


def hyper_fused_stream(agg_params, d: int) -> torch.Tensor:
    """Weight stream of csrc/gn_hyper_fused_tc.cu (h_dim 256): the aggregation MLPs' chunks in MMA
    consumption order — G1 of step s before G2 of step s-1 — each chunk one canonical operand."""
    t = len(agg_params)                                                # [(w0 (128,D), b0, w1 (D,128), b1)] fp32
    dev = agg_params[0][0].device
    parts = []
    for s in range(t + 1):
        if s < t:
            w0 = agg_params[s][0]
            hi, lo = _hi_lo(agg_params[s][1])
            parts.append(_canon(w0[:, :128].contiguous()))             # K chunk 0
            blk = torch.zeros(128, d - 128 + 16, dtype=torch.float32, device=dev)
            blk[:, :d - 128] = w0[:, 128:]
            blk[:, d - 128], blk[:, d - 127] = hi, lo
            parts.append(_canon(blk))                                  # K chunk 1 | bias k-columns
        if s >= 1:
            w1 = agg_params[s - 1][2]
            hi, lo = _hi_lo(agg_params[s - 1][3])
            blk = torch.zeros(d, 80, dtype=torch.float32, device=dev)
            blk[:, :64] = w1[:, :64]
            blk[:, 64], blk[:, 65], blk[:, 66] = hi, lo, hi
            parts.append(_canon(blk))
            parts.append(_canon(w1[:, 64:].contiguous()))
    return torch.cat(parts).contiguous()


# ---------------------------------------------------------------------------
# 3xTF32 weight streams (csrc/gn_chain_tf32.cu, csrc/gn_tf32.cuh)
# ---------------------------------------------------------------------------
TF_SMEM_BUDGET = 227 * 1024
TF_NODE_BLOCK = 36 * (68 + 132) * 4
TF_FIXED = (4 * 128 * 2 + 4 * 128 + 128 * 17) * 4 + 3072 * 4 + 1024


def tf_stage_bytes(a0_k: int, nbuf: int, node_block: bool) -> int:
    """Weight-ring stage of a chain whose staged A buffers hold `nbuf` x (128 x a0_k) hi+lo operands: the largest of
    64 / 32 / 16 KB that leaves room for two stages; must match tfe::ring_stage_bytes (csrc/gn_chain_tf32.cuh)."""
    used = nbuf * a0_k * 1024 + (TF_NODE_BLOCK if node_block else 0) + TF_FIXED
    avail = max(TF_SMEM_BUDGET - used, 0)
    return 65536 if avail >= 2 * 65536 else (32768 if avail >= 2 * 32768 else 16384)



def tf32_split(w: torch.Tensor):
    """w (fp32) -> (hi, lo): hi = w rounded to tf32 (11 significand bits, nearest, ties away in magnitude),
    lo = w - hi rounded the same way.  Both are fp32 tensors whose low 13 bits are zero."""
    def rnd(x):
        bits = x.contiguous().view(torch.int32)
        return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    hi = rnd(w)
    return hi, rnd(w - hi)


def tf_chunk_k(n: int, k: int, stage_bytes: int) -> int:
    """Largest multiple of 8 that divides K and keeps an (N x kc) hi+lo chunk inside one ring stage;
    must match tfe::Builder::chunk_k."""
    kc = (stage_bytes // 8) // n // 8 * 8
    kc = min(kc, k)
    while kc > 8 and k % kc:
        kc -= 8
    return kc


def _canon32(m: torch.Tensor) -> torch.Tensor:
    """(R, K) fp32 -> flat, canonical K-major operand with 32-bit elements [K/4][R][4]."""
    r, k = m.shape
    return m.reshape(r, k // 4, 4).permute(1, 0, 2).contiguous().reshape(-1)


def tf_stream(mats, stage_bytes: int, tail=None) -> torch.Tensor:
    """Weight stream of a chain: for every nn.Linear.weight (N, K) in consumption order, K chunks of kc columns,
    each chunk = canonical hi copy followed by canonical lo copy; `tail`: plain fp32 values appended after the chunks."""
    parts = []
    for w in mats:
        n, k = w.shape
        assert n % 16 == 0 and 16 <= n <= 256 and k % 8 == 0, (n, k)
        kc = tf_chunk_k(n, k, stage_bytes)
        hi, lo = tf32_split(w.contiguous())
        for c in range(k // kc):
            parts.append(_canon32(hi[:, c * kc:(c + 1) * kc].contiguous()))
            parts.append(_canon32(lo[:, c * kc:(c + 1) * kc].contiguous()))
    if tail is not None:
        parts.append(tail.reshape(-1).to(torch.float32))
    return torch.cat(parts).contiguous()


def pair_agg_tf32_stream(w0s, w1s) -> torch.Tensor:
    """Weight stream of the fused pairwise aggregation (csrc/gn_pair_agg_tf32.cu): one (64 x 64) hi | lo chunk per
    GEMM in issue order.  Unit step u = 2t + half uses A(u) = W0_t[64 half : 64 half + 64, :] (first Linear, 64 of its
    128 hidden units) and B(u) = W1_t[:, 64 half : 64 half + 64] (second Linear, the matching K slice).  GEMM 1 runs
    ahead of the drains: A(0), A(1), A(2), then per step st = 0..U-1: B(st - 1) (st >= 1), A(st + 3) (st + 3 < U); B(U - 1)."""
    t = len(w0s)
    u_n = 2 * t

    def a_of(u):
        return w0s[u // 2][64 * (u % 2):64 * (u % 2) + 64, :]

    def b_of(u):
        return w1s[u // 2][:, 64 * (u % 2):64 * (u % 2) + 64]

    return tf_stream([m.contiguous() for _, m in pair_agg_tf32_order(u_n, a_of, b_of)], 32768)


def pair_agg_tf32_order(u_n: int, a_of=None, b_of=None):
    """[(("a" | "b", u), matrix or None)] in the issue order of pair_agg_tf32_kernel's MMA warp."""
    a_of = a_of or (lambda u: None)
    b_of = b_of or (lambda u: None)
    out = [(("a", 0), a_of(0)), (("a", 1), a_of(1))]
    if u_n > 2:
        out.append((("a", 2), a_of(2)))
    for st in range(u_n):
        if st >= 1:
            out.append((("b", st - 1), b_of(st - 1)))
        if st + 3 < u_n:
            out.append((("a", st + 3), a_of(st + 3)))
    out.append((("b", u_n - 1), b_of(u_n - 1)))
    return out


def agg_out_cols(d: int) -> Tuple[int, int]:
    """(Dc, TN) of the aggregation output GEMM; must match make_plan() in
    csrc/gn_stage_simt.cu."""
    if d <= 64:
        return 64, 64
    return _round_up(d, 128), 128


def stage_modules(layer, s: int):
    """Sub-modules used by message-passing stage `s` of a layer with L = nmp_layers
    (MS_HGNN_batch.py:173-195): dict-softmax, closing MLP."""
    n_stage = max(layer.nmp_layers, 1)
    dict_mod = layer.nmp_mlp_start if s == 0 else layer.nmp_mlps[2 * s - 1]
    post_mod = layer.nmp_mlp_end if s == n_stage - 1 else layer.nmp_mlps[2 * s]
    return dict_mod, post_mod


def pack_stage(layer, s: int, device: torch.device) -> Dict[str, torch.Tensor]:
    d = layer.h_dim
    t = layer.edge_types
    dp = _round_up(d, 16)
    k2p = _round_up(2 * d, 16)
    dict_mod, post_mod = stage_modules(layer, s)
    node = layer.node2edge_start_mlp[s].layers
    att = layer.attention_mlp[s].layers
    agg = layer.edge_aggregation_list[s].agg_mlp
    dout = post_mod.layers[1].weight.shape[0]
    doutc = _round_up(dout, 64)
    dc, agg_tn = agg_out_cols(d)

    def dev(x):
        return x.detach().to(device=device, dtype=torch.float32).contiguous()

    out: Dict[str, torch.Tensor] = {}
    out["node_w0t"] = _kmajor(dev(node[0].weight), dp, 256, 128)
    out["node_b0"] = dev(node[0].bias)
    out["node_w1t"] = _kmajor(dev(node[1].weight), 256, 64, 64)
    out["node_b1"] = dev(node[1].bias)

    w0 = dev(att[0].weight)                                   # (32, 128): [x_n ; edge_init_e]
    wpq = torch.cat((w0[:, :64].t(), w0[:, 64:].t()), dim=1)  # (64, 64): pn | q
    out["att_wpqt"] = permute_cols(wpq.contiguous(), 64)
    out["att_b0"] = dev(att[0].bias)
    out["att_w1"] = dev(att[1].weight).reshape(-1)
    out["att_b1"] = dev(att[1].bias).reshape(-1)

    init, dist, fac = dict_mod.init_MLP.layers, dict_mod.MLP_distribution.layers, dict_mod.MLP_factor.layers
    out["init_w0t"] = _kmajor(dev(init[0].weight), 64, 128, 128)
    out["init_b0"] = dev(init[0].bias)
    out["init_w1t"] = _kmajor(dev(init[1].weight), 128, 64, 64)
    out["init_b1"] = dev(init[1].bias)
    df0 = torch.cat((dev(dist[0].weight), dev(fac[0].weight)), dim=0)     # (256, 64)
    out["df_w0t"] = _kmajor(df0, 64, 256, 128)
    out["df_b0"] = torch.cat((dev(dist[0].bias), dev(fac[0].bias)))
    w1 = torch.zeros(256, 16, dtype=torch.float32, device=device)
    w1[:128, :t] = dev(dist[1].weight).t()
    w1[128:, t] = dev(fac[1].weight).reshape(-1)
    out["df_w1"] = w1
    b1 = torch.zeros(16, dtype=torch.float32, device=device)
    b1[:t] = dev(dist[1].bias)
    b1[t] = dev(fac[1].bias).reshape(-1)[0]
    out["df_b1"] = b1

    a0 = torch.cat([dev(m.layers[0].weight) for m in agg], dim=0)         # (T*128, D)
    out["agg_w0t"] = _kmajor(a0, dp, t * 128, 128)
    out["agg_b0"] = torch.cat([dev(m.layers[0].bias) for m in agg])
    a1 = torch.zeros(t * 128, dc, dtype=torch.float32, device=device)
    for i, m in enumerate(agg):
        a1[i * 128:(i + 1) * 128, :d] = dev(m.layers[1].weight).t()
    out["agg_w1t"] = permute_cols(a1, agg_tn)
    out["agg_b1"] = torch.stack([dev(m.layers[1].bias) for m in agg]).contiguous()

    # tensor-core (bf16) copies of the per-edge MLP chain
    out["tc_init_w0"] = canonical_bf16(dev(init[0].weight))                      # (128, 64)
    out["tc_init_w1"] = canonical_bf16(dev(init[1].weight))                      # (64, 128)
    out["tc_df_w0"] = canonical_bf16(df0)                                        # (256, 64)
    out["tc_df_w1"] = canonical_bf16(w1.t().contiguous())                        # (16, 256)

    tc_ok = d % 16 == 0 and dout % 16 == 0 and dout <= 256       # must match make_plan() (tc_nodes)
    if tc_ok:
        out["tc_node_w0"] = canonical_bf16(dev(node[0].weight))                                   # (256, D)
        out["tc_node_w1"] = canonical_bf16(dev(node[1].weight))                                   # (64, 256)
        out["tc_att_wpq"] = canonical_bf16(torch.cat((w0[:, :64], w0[:, 64:]), dim=0))            # (64, 64)
        out["tc_agg_w0"] = canonical_bf16(a0)                                                     # (T*128, D)
        out["tc_agg_w1"] = canonical_bf16(torch.cat([dev(m.layers[1].weight) for m in agg], dim=1))  # (D, T*128)
        out["tc_post_w0"] = canonical_bf16(dev(post_mod.layers[0].weight))                        # (128, 2D)
        out["tc_post_w1"] = canonical_bf16(dev(post_mod.layers[1].weight))                        # (Dout, 128)
        if d == 256:
            out["tc_npre_w"] = node_pre256_stream(dev(node[0].weight), dev(node[0].bias), dev(node[1].weight),
                                                  dev(node[1].bias), torch.cat((w0[:, :64], w0[:, 64:]), dim=0).contiguous())
        if d == 64 and not layer._pairwise and t <= 15 and dout in (32, 64):
            out["tc_hfuse_w"] = hyper_fused64_stream(
                [(dev(m.layers[0].weight), dev(m.layers[0].bias), dev(m.layers[1].weight), dev(m.layers[1].bias))
                 for m in agg],
                dev(post_mod.layers[0].weight), dev(post_mod.layers[0].bias),
                dev(post_mod.layers[1].weight), dev(post_mod.layers[1].bias))
        if d == 256 and not layer._pairwise and t <= 15:
            stream = hyper_fused_stream([(dev(m.layers[0].weight), dev(m.layers[0].bias),
                                          dev(m.layers[1].weight), dev(m.layers[1].bias)) for m in agg], d)
            if dout % 32 == 0:
                stream = torch.cat((stream, hyper_fused_post_stream(
                    dev(post_mod.layers[0].weight), dev(post_mod.layers[0].bias),
                    dev(post_mod.layers[1].weight), dev(post_mod.layers[1].bias))))
            out["tc_hfuse_w"] = stream.contiguous()

    # fp32-grade tensor-core path (GN_TF32X3): 3xTF32 weight streams of the chains whose shape fits
    # (the *_tf32_fits() predicates of csrc/gn_chain_tf32.cu; absent streams fall back to the FFMA kernels)
    # the factor head (128 -> 1) and the distribution head (128 -> T) are fp32 dots in the drains: plain fp32 tail,
    # the latter k-major as [128][6, 8, 10, 12 or 16] (rows of T logits, zero padded)
    w4 = torch.zeros(128, next(tp for tp in (6, 8, 10, 12, 16) if t <= tp), dtype=torch.float32, device=device)
    w4[:, :t] = dev(dist[1].weight).t()
    out["tf_chain_w"] = tf_stream([dev(init[0].weight), dev(init[1].weight), dev(fac[0].weight), dev(dist[0].weight)],
                                  min(tf_stage_bytes(64, 1, False), tf_stage_bytes(0, 1, True)),
                                  tail=torch.cat((dev(fac[1].weight).reshape(-1), w4.reshape(-1))))
    if d % 8 == 0 and d <= 128:
        nw0, nw1 = dev(node[0].weight), dev(node[1].weight)
        pre = [nw0[:128], nw1[:, :128], nw0[128:], nw1[:, 128:], torch.cat((w0[:, :64], w0[:, 64:]), dim=0)]
        if layer._pairwise:
            pre.append(dev(init[0].weight))          # Y = x' W_init0^T per node (csrc/gn_chain_tf32.cu, pair form)
        out["tf_pre_w"] = tf_stream(pre, tf_stage_bytes(d, 1, False))
        if layer._pairwise:
            out["tf_aggin_w"] = tf_stream([dev(m.layers[0].weight) for m in agg], tf_stage_bytes(d, 1, False))
    if d in (64, 128):
        if layer._pairwise:
            w1cat = torch.cat([dev(m.layers[1].weight) for m in agg], dim=1)      # (D, T*128)
            out["tf_aggout_w"] = tf_stream([w1cat[:, 64 * c:64 * (c + 1)] for c in range(2 * t)],
                                           tf_stage_bytes(64, 2, False))
            if d == 64:
                out["tf_pagg_w"] = pair_agg_tf32_stream([dev(m.layers[0].weight) for m in agg],
                                                        [dev(m.layers[1].weight) for m in agg])
        else:
            mats = []
            for m in agg:
                mats += [dev(m.layers[0].weight), dev(m.layers[1].weight)]
            out["tf_hagg_w"] = tf_stream(mats, tf_stage_bytes(d, 1, False))
    if d % 4 == 0 and d <= 64 and dout in (64, 128):
        out["tf_post_w"] = tf_stream([dev(post_mod.layers[0].weight), dev(post_mod.layers[1].weight)],
                                     tf_stage_bytes(2 * d, 1, False))

    out["post_w0t"] = _kmajor(dev(post_mod.layers[0].weight), k2p, 128, 128)
    out["post_b0"] = dev(post_mod.layers[0].bias)
    out["post_w1t"] = _kmajor(dev(post_mod.layers[1].weight), 128, doutc, 64)
    out["post_b1"] = dev(post_mod.layers[1].bias)
    return out


class PackedStage:
    """Device copies of one stage's packed weights + the ctypes struct over them."""

    def __init__(self, layer, s: int, device: torch.device):
        self.tensors = pack_stage(layer, s, device)
        self.dout = int(self.tensors["post_b1"].numel())
        self.struct = _lib.StageWeights()
        for name in _lib.StageWeights.FIELDS:
            tens = self.tensors.get(name)
            if tens is None:                      # optional tensor-core copies (shape not eligible)
                setattr(self.struct, name, C.c_void_p(0))
                continue
            assert tens.is_contiguous() and tens.dtype == (torch.bfloat16 if name.startswith("tc_") else torch.float32)
            assert tens.data_ptr() % 16 == 0
            setattr(self.struct, name, C.c_void_p(tens.data_ptr()))


def _invalidate_after_load(module, incompatible_keys) -> None:
    """load_state_dict post-hook (module-level so that modules carrying it still pickle)."""
    module.invalidate_packs()


class RuntimeStateMixin:
    """Runtime caches (packed weights, workspaces, streams) live as plain attributes on the modules; they hold
    ctypes structures and device scratch that must not travel with a copy of the module.  Subclasses list the
    attribute names in `_RUNTIME_ATTRS` and rebuild them in `_reset_runtime()`:

      * `copy.deepcopy(m)`, `pickle` / `torch.save(m)` drop them and the copy starts with fresh caches;
      * `nn.DataParallel` replicas (shallow `__dict__` copies) get their own caches, not a shared workspace;
      * `load_state_dict` invalidates the packed weights;
      * `invalidate_packs()` is the explicit hook for writes that autograd's version counter does not see
        (`p.data.mul_(2)`, `nn.init.*_(p.data)`): the pack cache is keyed on (data_ptr, _version)."""
    _RUNTIME_ATTRS: Tuple[str, ...] = ()

    def _reset_runtime(self) -> None:          # pragma: no cover - overridden
        pass

    def invalidate_packs(self):
        """Forget every packed / folded copy of the weights held by this module and its sub-modules."""
        for m in self.modules():
            if isinstance(m, RuntimeStateMixin):
                m._reset_runtime()
        return self

    def _install_runtime_hooks(self) -> None:
        self.register_load_state_dict_post_hook(_invalidate_after_load)

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._RUNTIME_ATTRS:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._reset_runtime()

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica._reset_runtime()
        return replica


class PackCache:
    """Per-layer cache of PackedStage objects, invalidated when a parameter is
    modified in place through autograd-visible ops (optimizer step, load_state_dict, copy_) or moved.
    Writes through `.data` do not bump `_version`: call `invalidate_packs()` on the module after them."""

    def __init__(self) -> None:
        self._key = None
        self._stages: List[PackedStage] = []

    @staticmethod
    def _fingerprint(layer, device) -> tuple:
        """(device, data_ptr, version, data_ptr, version, ...) over the parameters in registration order.  Walks
        `_parameters` / `_modules` directly: `layer.parameters()` builds qualified names and a de-duplication
        set on the way and costs 2.5x as much, on every forward."""
        out = [str(device)]

        def walk(m):
            for prm in m._parameters.values():
                if prm is not None:
                    out.append(prm.data_ptr())
                    out.append(prm._version)
            for child in m._modules.values():
                if child is not None:
                    walk(child)
        walk(layer)
        return tuple(out)

    def get(self, layer, device: torch.device) -> List[PackedStage]:
        key = self._fingerprint(layer, device)
        if key != self._key:
            n_stage = max(layer.nmp_layers, 1)
            self._stages = [PackedStage(layer, s, device) for s in range(n_stage)]
            self._key = key
        return self._stages


# ---------------------------------------------------------------------------
# trajectory decoder (csrc/gn_decoder_simt.cu)
# ---------------------------------------------------------------------------
DEC_STATE, DEC_GATE = 96, 128


def _pad_gates(w: torch.Tensor) -> torch.Tensor:
    """(3*96, K) GRU weight with PyTorch's gate order r|z|n -> (3*128, K), each gate zero-padded to 128 rows."""
    k = w.shape[1]
    out = torch.zeros(3 * DEC_GATE, k, dtype=torch.float32, device=w.device)
    for g in range(3):
        out[g * DEC_GATE:g * DEC_GATE + DEC_STATE] = w[g * DEC_STATE:(g + 1) * DEC_STATE]
    return out


def pack_decoder_block(block, device: torch.device) -> Dict[str, torch.Tensor]:
    """Tensors of one `struct gn_decoder_weights` from a DecomposeBlock (model/GroupNet_nba.py:13-46)."""

    def dev(x):
        return x.detach().to(device=device, dtype=torch.float32).contiguous()

    gru = block.encoder_past
    if gru.hidden_size != DEC_STATE or gru.input_size != 32 or gru.num_layers != 1 or gru.bidirectional:
        raise ValueError("decoder kernel is built for GRU(32 -> 96), one layer")
    out: Dict[str, torch.Tensor] = {}
    out["conv_w"] = dev(block.conv_past.weight)                                   # (32, 2, 3)
    out["conv_b"] = dev(block.conv_past.bias)
    out["gru_wx"] = _kmajor(_pad_gates(dev(gru.weight_ih_l0)), 32, 3 * DEC_GATE, 128)
    out["gru_wh"] = _kmajor(_pad_gates(dev(gru.weight_hh_l0)), DEC_STATE, 3 * DEC_GATE, 128)
    b_ih, b_hh = dev(gru.bias_ih_l0), dev(gru.bias_hh_l0)
    gb = torch.zeros(4, DEC_GATE, dtype=torch.float32, device=device)
    gb[0, :DEC_STATE] = b_ih[:DEC_STATE] + b_hh[:DEC_STATE]                       # r
    gb[1, :DEC_STATE] = b_ih[DEC_STATE:2 * DEC_STATE] + b_hh[DEC_STATE:2 * DEC_STATE]   # z
    gb[2, :DEC_STATE] = b_ih[2 * DEC_STATE:]                                      # b_in
    gb[3, :DEC_STATE] = b_hh[2 * DEC_STATE:]                                      # b_hn (inside r * (.))
    out["gru_b"] = gb
    for tag, mlp in (("x", block.decoder_x), ("y", block.decoder_y)):
        l0, l1, l2 = mlp.layers
        if l0.weight.shape[0] != 512 or l1.weight.shape != (256, 512) or l2.weight.shape[1] != 256 \
                or l2.weight.shape[0] > 64:
            raise ValueError("decoder kernel is built for MLPs in -> 512 -> 256 -> (<= 64)")
        kp = _round_up(l0.weight.shape[1], 16)
        out[f"{tag}_w0"] = _kmajor(dev(l0.weight), kp, 512, 128)
        out[f"{tag}_b0"] = dev(l0.bias)
        out[f"{tag}_w1"] = _kmajor(dev(l1.weight), 512, 256, 128)
        out[f"{tag}_b1"] = dev(l1.bias)
        out[f"{tag}_w2"] = _kmajor(dev(l2.weight), 256, 64, 64)
        b2 = torch.zeros(64, dtype=torch.float32, device=device)
        b2[:l2.bias.numel()] = dev(l2.bias)
        out[f"{tag}_b2"] = b2
    return out


def pack_decoder_block_tc(block, device: torch.device) -> Dict[str, torch.Tensor]:
    """Tensors of one `struct gn_decoder_tc_weights` (bf16 tensor-core decoder, csrc/gn_decoder_tc.cu) from a
    DecomposeBlock (model/GroupNet_nba.py:13-46).  bf16 tensors are flat canonical operands ([K/8][N][8])."""

    def dev(x):
        return x.detach().to(device=device, dtype=torch.float32).contiguous()

    gru = block.encoder_past
    if gru.hidden_size != DEC_STATE or gru.input_size != 32 or gru.num_layers != 1 or gru.bidirectional:
        raise ValueError("decoder kernel is built for GRU(32 -> 96), one layer")
    h = DEC_STATE
    w_ih, w_hh = dev(gru.weight_ih_l0), dev(gru.weight_hh_l0)                     # (288, 32), (288, 96); gates r|z|n
    # gate matrix over K = [e (32) | h (96)]: r and z see both, n_x only e, n_h only h (it is scaled by r in the drain)
    rz = torch.cat([w_ih[:2 * h], w_hh[:2 * h]], dim=1)                           # (192, 128)
    nx = torch.cat([w_ih[2 * h:], torch.zeros(h, h, device=device)], dim=1)       # (96, 128)
    nh = torch.cat([torch.zeros(h, 32, device=device), w_hh[2 * h:]], dim=1)      # (96, 128)
    out: Dict[str, torch.Tensor] = {}
    out["conv_w"] = dev(block.conv_past.weight)
    out["conv_b"] = dev(block.conv_past.bias)
    out["gru_w"] = torch.cat([_canon(rz), _canon(torch.cat([nx, nh], dim=0))]).contiguous()
    b_ih, b_hh = dev(gru.bias_ih_l0), dev(gru.bias_hh_l0)
    gb = torch.zeros(4, DEC_GATE, dtype=torch.float32, device=device)
    gb[0, :h] = b_ih[:h] + b_hh[:h]
    gb[1, :h] = b_ih[h:2 * h] + b_hh[h:2 * h]
    gb[2, :h] = b_ih[2 * h:]
    gb[3, :h] = b_hh[2 * h:]
    out["gru_b"] = gb
    x0, x1, x2 = block.decoder_x.layers
    y0, y1, y2 = block.decoder_y.layers
    for l0, l1, l2 in ((x0, x1, x2), (y0, y1, y2)):
        if l0.weight.shape[0] != 512 or l1.weight.shape != (256, 512) or l2.weight.shape[1] != 256 \
                or l2.weight.shape[0] > 64 or l0.weight.shape[1] % 16:
            raise ValueError("tensor-core decoder is built for MLPs in (multiple of 16) -> 512 -> 256 -> (<= 64)")
    out["w0"] = _canon(torch.cat([dev(x0.weight), dev(y0.weight)], dim=0))
    out["b0"] = torch.cat([dev(x0.bias), dev(y0.bias)]).contiguous()
    for tag, l1, l2 in (("x", x1, x2), ("y", y1, y2)):
        out[f"{tag}_w1"] = _canon(dev(l1.weight))
        out[f"{tag}_b1"] = dev(l1.bias)
        n_out = l2.weight.shape[0]
        pad = _round_up(n_out, 16)
        w2 = torch.zeros(pad, 256, dtype=torch.float32, device=device)
        w2[:n_out] = dev(l2.weight)
        b2 = torch.zeros(pad, dtype=torch.float32, device=device)
        b2[:n_out] = dev(l2.bias)
        out[f"{tag}_w2"] = _canon(w2)
        out[f"{tag}_b2"] = b2
    out["mlp_stream"], out["mlp_bias"] = decoder_mlp_stream(block, device)
    return out


def decoder_mlp_stream(block, device: torch.device):
    """Weights of decoder_x then decoder_y as the stage stream of `decoder_mlp_fused_kernel` (csrc/gn_decoder_tc.cu): 16 KB
    stages of canonical [8 k-groups][rows][8] bf16 tiles in the issuer's order per MLP,
        G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) G2(3) G3
    G1(c): K / 64 stages W0[c*128:(c+1)*128, ks*64:(ks+1)*64];  G2(c): 4 stages W1[nh*128:(nh+1)*128, c*128 + kh*64 : + 64]
    (kh outer, nh inner);  G3: 4 stages W2 (rows zero-padded to 32)[:, ks*64:(ks+1)*64], 4 KB used of each 16 KB slot.
    Biases: b0 x (512) | b0 y (512) | b1 x (256) | b1 y (256) | b2 x (32) | b2 y (32).  Feature widths that are not a
    multiple of 64 get no stream (a zero-length tensor: the row-tile GEMMs run)."""
    def dev(x):
        return x.detach().to(device=device, dtype=torch.float32).contiguous()

    kf = block.decoder_x.layers[0].weight.shape[1]
    bias = torch.zeros(1600, dtype=torch.float32, device=device)
    if kf % 64 or kf > 384 or block.decoder_x.layers[2].weight.shape[0] > 32 or block.decoder_y.layers[2].weight.shape[0] > 32:
        return torch.zeros(0, dtype=torch.bfloat16, device=device), bias
    stage = 8192                                                     # bf16 elements per 16 KB slot
    chunks = []

    def put(tile):
        flat = _canon(tile)
        slot = torch.zeros(stage, dtype=torch.bfloat16, device=device)
        slot[:flat.numel()] = flat
        chunks.append(slot)

    for i, mlp in enumerate((block.decoder_x, block.decoder_y)):
        l0, l1, l2 = mlp.layers
        w0, w1 = dev(l0.weight), dev(l1.weight)
        w2 = torch.zeros(32, 256, dtype=torch.float32, device=device)
        w2[:l2.weight.shape[0]] = dev(l2.weight)

        def g1(c):
            for ks in range(kf // 64):
                put(w0[c * 128:(c + 1) * 128, ks * 64:(ks + 1) * 64])

        def g2(c):
            for kh in range(2):
                for nh in range(2):
                    put(w1[nh * 128:(nh + 1) * 128, c * 128 + kh * 64:c * 128 + (kh + 1) * 64])

        g1(0); g1(1); g2(0); g1(2); g2(1); g1(3); g2(2); g2(3)
        for ks in range(4):
            put(w2[:, ks * 64:(ks + 1) * 64])
        bias[i * 512:(i + 1) * 512] = dev(l0.bias)
        bias[1024 + i * 256:1024 + (i + 1) * 256] = dev(l1.bias)
        bias[1536 + i * 32:1536 + i * 32 + l2.bias.numel()] = dev(l2.bias)
    return torch.cat(chunks).contiguous(), bias
