"""Drop-in `PastEncoder` (model/GroupNet_nba.py:196-315) — SURVEY.md §8(f) rank 1.

Same constructor (`PastEncoder(args, in_dim=4)`), attribute names and state_dict schema as the
reference (`input_fc`, `input_fc2`, `input_fc3`, `interaction`, `interaction_hyper[2,3]`,
`pos_encoder.{fc, pe}`), same forward signature and outputs:
    forward(inputs (B*N, T, in_dim), batch_size, agent_num) -> (output_feature (B*N, D*(2+S)), new_H (B, sum E, N))

Front-end: in eval mode `input_fc -> [x ; pos_enc] -> pos_encoder.fc -> input_fc2 -> add_category ->
input_fc3` (:269-280) contains no nonlinearity (the only stochastic op is `pos_encoder.dropout`, identity
in eval), so it is folded on the host into ONE affine map R^{T*in_dim} -> R^D with a per-agent bias (the
3-way player/ball category, :256-265) and evaluated by `gn_past_frontend`.  The interaction block is
`MultiScaleInteraction` (fused corr/top-k, in-place concatenation).  Training through the front-end
(dropout active) is not implemented in round 1: the fused path requires `torch.no_grad()` / eval.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from .interaction import MultiScaleInteraction, _HYPER_NAMES
from .layers import MS_HGNN_hyper, MS_HGNN_oridinary
from .packing import RuntimeStateMixin


class PositionalAgentEncoding(nn.Module):
    """Schema of the reference class (:156-195): `fc` (2D -> D) when concat, buffer `pe` (max_t_len, D)."""

    def __init__(self, d_model, dropout=0.1, max_t_len=200, concat=True):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.concat = concat
        self.d_model = d_model
        if concat:
            self.fc = nn.Linear(2 * d_model, d_model)
        pe = torch.zeros(max_t_len, d_model)
        position = torch.arange(0, max_t_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe)

    def forward(self, *a, **k):
        raise RuntimeError("groupnet_b200.PositionalAgentEncoding is a parameter container (gn_past_frontend)")


class PastEncoder(RuntimeStateMixin, nn.Module):
    def __init__(self, args, in_dim=4):
        super().__init__()
        self.args = args
        self.model_dim = args.hidden_dim
        self.scale_number = len(args.hyper_scales)
        d = self.model_dim
        # registration / construction order of the reference (:203-248)
        self.input_fc = nn.Linear(in_dim, d)
        self.input_fc2 = nn.Linear(d * args.past_length, d)
        self.input_fc3 = nn.Linear(d + 3, d)
        self.interaction = MS_HGNN_oridinary(embedding_dim=16, h_dim=d, mlp_dim=64, bottleneck_dim=d,
                                             batch_norm=0, nmp_layers=1)
        for name, scale in zip(_HYPER_NAMES, args.hyper_scales):
            setattr(self, name, MS_HGNN_hyper(embedding_dim=d, h_dim=d, mlp_dim=64, bottleneck_dim=d,
                                              batch_norm=0, nmp_layers=1, scale=scale))
        self.pos_encoder = PositionalAgentEncoding(d, 0.1, concat=True)
        self._reset_runtime()
        self._install_runtime_hooks()

    _RUNTIME_ATTRS = ("_fold_key", "_fold", "_block")

    def _reset_runtime(self) -> None:
        self.__dict__["_fold_key"] = None
        self.__dict__["_fold"] = None
        self.__dict__["_block"] = None

    # ------------------------------------------------------------------
    def _interaction_block(self) -> MultiScaleInteraction:
        """A MultiScaleInteraction view over THIS module's layers (shared parameters, no copies)."""
        if self._block is None:
            blk = MultiScaleInteraction.__new__(MultiScaleInteraction)
            nn.Module.__init__(blk)
            blk.model_dim = self.model_dim
            blk.hyper_scales = [int(s) for s in self.args.hyper_scales]
            object.__setattr__(blk, "_shared_owner", self)
            blk.__dict__["interaction"] = self.interaction          # plain attributes: not re-registered
            for name in _HYPER_NAMES[:len(blk.hyper_scales)]:
                blk.__dict__[name] = getattr(self, name)
            self.__dict__["_block"] = blk
        return self._block

    def layers(self):
        return self._interaction_block().layers()

    def folded_frontend(self, agent_num: int, length: int, device):
        """(Mt (K,D), bias_agent (N,D)) of the affine front-end, folded in float64; cached per weight version."""
        ps = [self.input_fc.weight, self.input_fc.bias, self.input_fc2.weight, self.input_fc2.bias,
              self.input_fc3.weight, self.input_fc3.bias, self.pos_encoder.fc.weight, self.pos_encoder.fc.bias]
        key = (str(device), agent_num, length) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._fold_key:
            d = self.model_dim
            f64 = lambda t: t.detach().double().cpu()
            w1, b1 = f64(ps[0]), f64(ps[1])                       # (D,in), (D)
            w2, b2 = f64(ps[2]), f64(ps[3])                       # (D, D*T), (D)
            w3, b3 = f64(ps[4]), f64(ps[5])                       # (D, D+3), (D)
            wp, bp = f64(ps[6]), f64(ps[7])                       # (D, 2D), (D)
            pe = self.pos_encoder.pe.detach().double().cpu()[:length]             # (T, D)
            wpx, wpe = wp[:, :d], wp[:, d:]
            w3f, w3c = w3[:, :d], w3[:, d:]
            in_dim = w1.shape[1]
            m = torch.zeros(d, length * in_dim, dtype=torch.float64)
            const = b2.clone()
            for t in range(length):
                w2t = w2[:, t * d:(t + 1) * d]
                m[:, t * in_dim:(t + 1) * in_dim] = w3f @ w2t @ wpx @ w1
                const = const + w2t @ (wpx @ b1 + wpe @ pe[t] + bp)
            base = w3f @ const + b3
            # add_category (:256-265): agents 0-4 -> class 0, 5-9 -> class 1, 10 -> class 2, others none
            bias = base[None, :].repeat(agent_num, 1)
            for n in range(agent_num):
                if n < 5:
                    bias[n] += w3c[:, 0]
                elif n < 10:
                    bias[n] += w3c[:, 1]
                elif n == 10:
                    bias[n] += w3c[:, 2]
            self._fold = (m.t().contiguous().float().to(device), bias.contiguous().float().to(device))
            self._fold_key = key
        return self._fold

    def frontend(self, inputs: torch.Tensor, batch_size: int, agent_num: int) -> torch.Tensor:
        """ftraj_input (B, N, D) == model/GroupNet_nba.py:269-280 in eval mode."""
        ops._require_cuda_f32(inputs, "inputs")
        if agent_num < 11:
            # the reference indexes agent 10 in add_category (:264) and fails for smaller scenes
            raise IndexError("index 10 is out of bounds for dimension 0 with size %d" % agent_num)
        rows, length, in_dim = inputs.shape
        if rows != batch_size * agent_num:
            raise RuntimeError("shape mismatch: inputs must be (batch_size*agent_num, T, in_dim)")
        if self.training:
            raise NotImplementedError("groupnet_b200.PastEncoder: the fused front-end is eval-only "
                                      "(pos_encoder.dropout is active in training): call .eval() first")
        if torch.is_grad_enabled() and (inputs.requires_grad or any(
                p.requires_grad for p in (self.input_fc.weight, self.input_fc2.weight, self.input_fc3.weight,
                                          self.pos_encoder.fc.weight))):
            raise NotImplementedError("groupnet_b200.PastEncoder: the folded front-end has no backward, so gradients "
                                      "to input_fc* / pos_encoder would be dropped silently: wrap the call in "
                                      "torch.no_grad() (or freeze the front-end parameters)")
        mt, bias = self.folded_frontend(agent_num, length, inputs.device)
        x = inputs.contiguous()
        out = torch.empty(batch_size, agent_num, self.model_dim, dtype=torch.float32, device=inputs.device)
        if rows == 0:
            return out
        lib = _lib.load()
        with torch.cuda.device(inputs.device):
            rc = lib.gn_past_frontend(C.c_void_p(x.data_ptr()), rows, length * in_dim, agent_num, self.model_dim,
                                      C.c_void_p(mt.data_ptr()), C.c_void_p(bias.data_ptr()),
                                      C.c_void_p(out.data_ptr()), ops._stream_ptr(inputs.device))
        _lib.check(rc, "gn_past_frontend")
        return out

    def forward(self, inputs, batch_size, agent_num, *, noise=None):
        ftraj = self.frontend(inputs, batch_size, agent_num)
        final_feature, new_h = self._interaction_block()(ftraj, noise=noise)
        return final_feature.view(batch_size * agent_num, -1), new_h
