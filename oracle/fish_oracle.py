"""CPU oracle for the fish model's group-wise operators.

TEST INFRASTRUCTURE ONLY (see oracle/ms_hgnn_oracle.py): nothing under ``groupnet_b200/`` imports this module.

Functional restatements (pure functions over a ``state_dict``, eval-mode BatchNorm, dropout = identity) of
``model/encoder.py`` and ``utilities/utils.py`` of TaliMotzkin/GroupNet, op for op as written.  Parity status: PINNED
against the reference itself — live (``tests/test_oracle_vs_reference.py::test_fish_*``, wherever /root/reference
exists) and through fixtures generated from the unmodified reference classes (``tests/golden/make_golden_fish.py``).
Citations are ``model/encoder.py:<line>`` unless another file is named.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]
BN_EPS = 1e-5


def _lin(sd: StateDict, p: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _bn(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """nn.BatchNorm1d in eval mode on (rows, C)."""
    return (x - sd[p + ".running_mean"]) / torch.sqrt(sd[p + ".running_var"] + BN_EPS) * sd[p + ".weight"] + sd[p + ".bias"]


def _seq(sd: StateDict, p: str, x: Tensor) -> Tensor:
    """Sequential(Linear, BatchNorm1d, LeakyReLU(0.01), Linear, BatchNorm1d) (:125-139, :364-381)."""
    x = _bn(sd, p + ".1", _lin(sd, p + ".0", x))
    x = F.leaky_relu(x, 0.01)
    return _bn(sd, p + ".4", _lin(sd, p + ".3", x))


def compute_alpha_im(alpha_ij: Tensor, I_HG: Tensor, rel_rec: Tensor, rel_send: Tensor) -> Tensor:
    """:261-303, as written (the (B, E, N, M) masks are materialised)."""
    rec_m = (rel_rec.unsqueeze(-1) * I_HG.unsqueeze(1)).sum(2) > 0          # :280
    send_m = (rel_send.unsqueeze(-1) * I_HG.unsqueeze(1)).sum(2) > 0        # :281
    masked = alpha_ij * (rec_m & send_m)                                    # :287
    alpha_im = torch.einsum("bem,ben->bnm", masked, rel_rec)                # :291
    n_hm = I_HG.sum(dim=1, keepdim=True)                                    # :297
    return alpha_im / (n_hm - 1 + 1e-8)                                     # :299


def mlp_hge(sd: StateDict, alpha_im: Tensor, v_cg: Tensor) -> Tensor:
    """MLPHGE.forward (:224-251), eval mode."""
    alpha_norm = alpha_im / (alpha_im.sum(dim=1).unsqueeze(1) + 1e-8)       # :236-238
    w = torch.einsum("bnm,bnf->bmf", alpha_norm, v_cg)                      # :241
    b, m, _ = w.shape

    def bn(p, x):
        return _bn(sd, p, x.reshape(b * m, -1)).reshape(b, m, -1)           # :215-222

    x = F.elu(bn("bn", _lin(sd, "fc1", w)))                                 # :244
    x = F.elu(bn("bn", _lin(sd, "fc2", x)))                                 # :246
    return F.elu(bn("bn2", _lin(sd, "fc3", x)))                             # :248


def hyperedge_attention(sd: StateDict, e_hg: Tensor, v_cg: Tensor, I_HG: Tensor, alpha: float = 0.2) -> Tensor:
    """HyperEdgeAttention.forward (:141-197), eval mode."""
    b, n, m = I_HG.shape
    e_proj = F.leaky_relu(_lin(sd, "W1", e_hg), alpha)                      # :160
    v_proj = F.leaky_relu(_lin(sd, "W2", v_cg), alpha)                      # :161
    comb = torch.cat([e_proj.unsqueeze(1).expand(-1, n, -1, -1), v_proj.unsqueeze(2).expand(-1, -1, m, -1)], dim=-1)   # :165-169
    logits = F.leaky_relu(torch.einsum("bnmf,f->bnm", comb, sd["attention_vector"]), alpha)                            # :172
    logits = logits.masked_fill(I_HG == 0, float("-inf"))                   # :177
    a = torch.nan_to_num(F.softmax(logits / 100, dim=1), nan=0.0).transpose(1, 2)                                      # :181-182
    v1 = torch.einsum("bmn,bmf->bnf", a, e_hg)                              # :186
    v1 = F.leaky_relu(_seq(sd, "f_HG_v", v1.reshape(b * n, -1)).view(b, n, -1), alpha)                                 # :190-191
    e2 = torch.einsum("bnm,bnf->bmf", I_HG, v1)                             # :194
    return F.leaky_relu(_seq(sd, "f_HG_2", e2.reshape(b * m, -1)).view(b, m, -1), alpha)                               # :197-198


def temporal_gat(sd: StateDict, v_self: Tensor, rel_rec: Tensor, rel_send: Tensor, num_heads: int, out_dim: int,
                 concat_heads: bool = True, alpha: float = 0.2) -> Tuple[Tensor, Tensor]:
    """TemporalGATLayer.forward (:385-467), eval mode."""
    b, n, _ = v_self.shape
    h, d = num_heads, out_dim
    v_proj = F.leaky_relu(_lin(sd, "projection", v_self), alpha).view(b, n, h, d)            # :404-405
    h_src = torch.einsum("ben,bnhd->behd", rel_send, v_proj)                                  # :415
    h_tgt = torch.einsum("ben,bnhd->behd", rel_rec, v_proj)                                   # :416
    a_ij = F.leaky_relu(torch.einsum("behd,hd->beh", h_src, sd["a_forward"]), alpha) / 500    # :421
    a_ji = F.leaky_relu(torch.einsum("behd,hd->beh", h_tgt, sd["a_backward"]), alpha) / 500   # :424
    a_max = torch.maximum(a_ij, a_ji)                                                          # :427
    s_ij, s_ji = torch.exp(a_ij - a_max), torch.exp(a_ji - a_max)                              # :428-429
    a_sum = torch.exp(s_ij) + torch.exp(s_ji)                                                  # :432
    al_ij, al_ji = torch.exp(s_ij) / a_sum, torch.exp(s_ji) / a_sum                            # :433-434
    e_in = torch.cat([al_ij.unsqueeze(-1) * h_src, al_ji.unsqueeze(-1) * h_tgt], dim=-1)       # :443-449
    e = e_in.shape[1]
    e_cg = F.leaky_relu(_seq(sd, "f_CG_e", e_in.view(b * e * h, 2 * d)).view(b, e, h, d), alpha)                       # :450
    v_soc = torch.einsum("behd,ben->bnhd", e_cg * al_ij.unsqueeze(-1), rel_rec)               # :454-457
    v_soc = F.leaky_relu(_seq(sd, "f_CG_v", v_soc.reshape(b * n * h, d)), alpha).reshape(b, n, h, -1)                  # :459
    v_soc = v_soc.reshape(b, n, -1) if concat_heads else v_soc.mean(dim=2)                     # :462-465
    return v_soc, al_ij


def build_dynamic_graph_and_hypergraph(z_cg: Tensor, z_hg: Tensor, rel_rec: Tensor, rel_send: Tensor, I_HG: Tensor):
    """utilities/utils.py:191-244 (the per-batch Python loop is a mask product)."""
    b = z_cg.shape[0]
    rel_rec, rel_send = rel_rec.expand(b, -1, -1), rel_send.expand(b, -1, -1)
    et, ht = z_cg.argmax(dim=-1), z_hg.argmax(dim=-1)                        # utils.py:217-218
    ke, kh = (et != 0).to(rel_rec.dtype), (ht != 0).to(I_HG.dtype)           # :222-223
    return rel_rec * ke.unsqueeze(-1), rel_send * ke.unsqueeze(-1), I_HG * kh.unsqueeze(1), et, ht
