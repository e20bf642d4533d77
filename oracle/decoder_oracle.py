"""CPU oracle for GroupNet's trajectory decoder (`DecomposeBlock` / `Decoder`) — SURVEY.md §8(f) rank 2.

TEST INFRASTRUCTURE ONLY (same rules as ms_hgnn_oracle.py: nothing under ``groupnet_b200/`` may import it).
Round 1 ships the oracle, its pin and the fixtures; the CUDA path for this row is round-2 work.

Functional restatement (pure functions over a ``state_dict``) of ``model/GroupNet_nba.py:13-79`` and
``:441-505`` of TaliMotzkin/GroupNet.  The GRU recurrence is written out step by step (the reference calls
``nn.GRU``) so that it doubles as the specification of the fused kernel: gates in PyTorch's (r, z, n) order,
``n = tanh(W_in x + b_in + r * (W_hn h + b_hn))``, ``h' = (1 - z) * n + z * h``.

Parity status: PINNED against the reference itself — live differential test where /root/reference exists
(``tests/test_decoder_oracle.py``) and fixtures generated from the unmodified reference
(``tests/golden/make_golden_decoder.py`` -> ``tests/golden/decoder/*.npz``).

Citations are ``model/GroupNet_nba.py:<line>``.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .ms_hgnn_oracle import mlp

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


def conv_past(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    """`relu(conv_past(x^T))^T` (:60-67): Conv1d(2 -> 32, kernel 3, stride 1, padding 1) along time.
    x (R, T, 2) -> (R, T, 32)."""
    y = F.conv1d(x.transpose(1, 2), sd[f"{prefix}.conv_past.weight"], sd[f"{prefix}.conv_past.bias"], padding=1)
    return torch.relu(y).transpose(1, 2)


def gru_last_state(sd: StateDict, prefix: str, x: Tensor) -> Tensor:
    """`_, state = encoder_past(x); state.squeeze(0)` (:69-70): one-layer GRU(32 -> 96), batch_first, h0 = 0.
    x (R, T, I) -> (R, H)."""
    w_ih, w_hh = sd[f"{prefix}.encoder_past.weight_ih_l0"], sd[f"{prefix}.encoder_past.weight_hh_l0"]
    b_ih, b_hh = sd[f"{prefix}.encoder_past.bias_ih_l0"], sd[f"{prefix}.encoder_past.bias_hh_l0"]
    hid = w_hh.shape[1]
    h = torch.zeros(x.shape[0], hid, dtype=x.dtype)
    for t in range(x.shape[1]):
        gi = F.linear(x[:, t], w_ih, b_ih)             # (R, 3H): [r | z | n] input parts
        gh = F.linear(h, w_hh, b_hh)                   # (R, 3H): [r | z | n] hidden parts
        r = torch.sigmoid(gi[:, :hid] + gh[:, :hid])
        z = torch.sigmoid(gi[:, hid:2 * hid] + gh[:, hid:2 * hid])
        n = torch.tanh(gi[:, 2 * hid:] + r * gh[:, 2 * hid:])
        h = (1.0 - z) * n + z * h
    return h


def decompose_block(sd: StateDict, prefix: str, x_true: Tensor, x_hat: Tensor, f: Tensor,
                    past_len: int, future_len: int) -> Tuple[Tensor, Tensor]:
    """`DecomposeBlock.forward` (:48-79): residual past (R,T_p,2) -> conv+ReLU -> GRU state (R,96);
    [f ; state] -> decoder_x (-> R,T_p,2) and decoder_y (-> R,T_f,2), both MLPs 384 -> 512 -> 256 -> out."""
    state = gru_last_state(sd, prefix, conv_past(sd, prefix, x_true - x_hat))     # :58-70
    feat = torch.cat((f, state), dim=1)                                           # :72
    x_hat_after = mlp(sd, f"{prefix}.decoder_x", feat).view(-1, past_len, 2)      # :74
    y_hat = mlp(sd, f"{prefix}.decoder_y", feat).view(-1, future_len, 2)          # :75
    return x_hat_after, y_hat


def decoder_forward(sd: StateDict, past_feature: Tensor, z: Tensor, batch_size: int, agents_per_scene: int,
                    past_traj: Tensor, cur_location: Tensor, sample_num: int, *, past_len: int, future_len: int,
                    num_decompose: int, mode: str = "train") -> Tuple[Tensor, Tensor]:
    """`Decoder.forward` (:461-505).  past_feature (A*S, F), z (A*S, zdim), past_traj (A, T_p, 2),
    cur_location (A, 1, 2) with A = batch_size * agents_per_scene, S = sample_num.
    Returns (out_seq, recover_pre_seq): (A*S, T_f, 2) — (A, S, T_f, 2) in 'inference' mode — and (A*S, T_p, 2)."""
    agent_num = batch_size * agents_per_scene
    x_true = past_traj.repeat_interleave(sample_num, dim=0)                       # :464
    hidden = torch.cat((past_feature.view(-1, sample_num, past_feature.shape[-1]),
                        z.view(-1, sample_num, z.shape[-1])), dim=-1)             # :466-475
    hidden = hidden.view(agent_num * sample_num, -1)                              # :477
    x_hat = torch.zeros_like(x_true)                                              # :481
    prediction = torch.zeros(x_true.shape[0], future_len, 2, dtype=x_true.dtype)  # :485
    reconstruction = torch.zeros(x_true.shape[0], past_len, 2, dtype=x_true.dtype)
    for i in range(num_decompose):                                                # :490-493
        x_hat, y_hat = decompose_block(sd, f"decompose.{i}", x_true, x_hat, hidden, past_len, future_len)
        prediction = prediction + y_hat
        reconstruction = reconstruction + x_hat
    out_seq = prediction + cur_location.repeat_interleave(sample_num, dim=0)      # :501-502
    if mode == "inference":
        out_seq = out_seq.view(-1, sample_num, *out_seq.shape[1:])                # :503-504
    return out_seq, reconstruction
