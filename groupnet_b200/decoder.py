"""Drop-in `DecomposeBlock` / `Decoder` (model/GroupNet_nba.py:13-79, :441-505) — SURVEY.md §8(f) rank 2.

Same constructors, attribute names, parameter registration / initialisation order (hence the same seeded
default weights and state_dict schema) and the same forward signature and outputs as the reference:

    Decoder(args).forward(past_feature (A*S, F), z (A*S, zdim), batch_size_curr, agent_num_perscene,
                          past_traj (A, T_p, 2), cur_location (A, 1, 2), sample_num, mode='train')
        -> (out_seq (A*S, T_f, 2)  [(A, S, T_f, 2) when mode == 'inference'],  recover_pre_seq (A*S, T_p, 2))

The forward is `gn_decoder_fwd` (csrc/gn_decoder_simt.cu, fp32 FFMA path, 1e-5 parity: one kernel launch per
DecomposeBlock) or, after `set_precision("bf16")`, `gn_decoder_fwd_tc` (csrc/gn_decoder_tc.cu, tcgen05 tensor cores,
bf16 operands / fp32 accumulation, 2e-2 parity); there is no PyTorch implementation of the math here.  Inference
only in this round (no autograd through the decoder).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, ops
from .layers import MLP
from .packing import PackCache, RuntimeStateMixin, pack_decoder_block, pack_decoder_block_tc


class DecomposeBlock(nn.Module):
    """Parameter container with the reference's schema (:18-46); evaluated by `Decoder.forward`."""

    def __init__(self, past_len, future_len, input_dim):
        super().__init__()
        channel_in, channel_out, dim_kernel, dim_embedding_key = 2, 32, 3, 96
        self.past_len = past_len
        self.future_len = future_len
        self.conv_past = nn.Conv1d(channel_in, channel_out, dim_kernel, stride=1, padding=1)
        self.encoder_past = nn.GRU(channel_out, dim_embedding_key, 1, batch_first=True)
        self.decoder_y = MLP(dim_embedding_key + input_dim, future_len * 2, hidden_size=(512, 256))
        self.decoder_x = MLP(dim_embedding_key + input_dim, past_len * 2, hidden_size=(512, 256))
        self.relu = nn.ReLU()
        self.init_parameters()

    def init_parameters(self):
        nn.init.kaiming_normal_(self.conv_past.weight)
        nn.init.kaiming_normal_(self.encoder_past.weight_ih_l0)
        nn.init.kaiming_normal_(self.encoder_past.weight_hh_l0)
        nn.init.zeros_(self.conv_past.bias)
        nn.init.zeros_(self.encoder_past.bias_ih_l0)
        nn.init.zeros_(self.encoder_past.bias_hh_l0)

    def forward(self, *a, **k):
        raise RuntimeError("groupnet_b200.DecomposeBlock is evaluated by groupnet_b200.Decoder (gn_decoder_fwd)")


class Decoder(RuntimeStateMixin, nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.model_dim = args.hidden_dim
        self.decode_way = 'RES'
        scale_num = 2 + len(self.args.hyper_scales)
        self.num_decompose = args.num_decompose
        input_dim = scale_num * self.model_dim + self.args.zdim
        self.past_length = self.args.past_length
        self.future_length = self.args.future_length
        self.decompose = nn.ModuleList(
            [DecomposeBlock(self.args.past_length, self.args.future_length, input_dim)
             for _ in range(self.num_decompose)])
        self.precision = "fp32"
        self._reset_runtime()
        self._install_runtime_hooks()

    _RUNTIME_ATTRS = ("_pack_key", "_packed", "_ws")

    def set_precision(self, precision: str):
        """"fp32": the FFMA kernel (1e-5 of the reference); "bf16": the tcgen05 path (2e-2)."""
        if precision not in ("fp32", "bf16"):
            raise ValueError("Decoder precision must be 'fp32' or 'bf16'")
        if precision != self.precision:
            self.precision = precision
            self.invalidate_packs()
        return self

    def _reset_runtime(self) -> None:
        self.__dict__["_pack_key"] = None
        self.__dict__["_packed"] = None
        self.__dict__["_ws"] = ops.Workspace()

    def _packs(self, device):
        tc = self.precision == "bf16"
        key = (tc,) + tuple(PackCache._fingerprint(self, device))
        if key != self._pack_key:
            pack, struct = (pack_decoder_block_tc, _lib.DecoderTcWeights) if tc else (pack_decoder_block, _lib.DecoderWeights)
            tensors = [pack(blk, device) for blk in self.decompose]
            structs = (struct * len(tensors))()
            for st, t in zip(structs, tensors):
                for name in struct.FIELDS:
                    assert t[name].is_contiguous() and t[name].dtype in (torch.float32, torch.bfloat16)
                    setattr(st, name, C.c_void_p(t[name].data_ptr() if t[name].numel() else 0))
            self._packed, self._pack_key = (tensors, structs), key
        return self._packed[1]

    def forward(self, past_feature, z, batch_size_curr, agent_num_perscene, past_traj, cur_location, sample_num,
                mode='train'):
        for name, t in (("past_feature", past_feature), ("z", z), ("past_traj", past_traj),
                        ("cur_location", cur_location)):
            ops._require_cuda_f32(t, name)
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or past_feature.requires_grad
                                        or z.requires_grad):
            raise NotImplementedError("groupnet_b200.Decoder is inference-only in this round: "
                                      "call it under torch.no_grad()")
        agents = int(batch_size_curr) * int(agent_num_perscene)
        s = int(sample_num)
        rows = agents * s
        f_dim, z_dim = past_feature.shape[-1], z.shape[-1]
        if past_feature.numel() != rows * f_dim or z.numel() != rows * z_dim:
            raise RuntimeError(f"shape mismatch: past_feature and z must hold {rows} rows (agents * sample_num)")
        if tuple(past_traj.shape) != (agents, self.past_length, 2) or cur_location.numel() != agents * 2:
            raise RuntimeError("shape mismatch: past_traj must be (agents, past_length, 2), cur_location (agents, 1, 2)")
        if f_dim + z_dim + 96 != self.decompose[0].decoder_x.layers[0].weight.shape[1]:
            raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (feature width does not match the decoder)")
        dev = past_feature.device
        pf = past_feature.detach().reshape(rows, f_dim).contiguous()
        zz = z.detach().reshape(rows, z_dim).contiguous()
        pt = past_traj.detach().contiguous()
        cl = cur_location.detach().reshape(agents, 2).contiguous()
        out_seq = torch.empty(rows, self.future_length, 2, dtype=torch.float32, device=dev)
        recover = torch.empty(rows, self.past_length, 2, dtype=torch.float32, device=dev)
        if rows:
            lib = _lib.load()
            structs = self._packs(dev)
            if self.precision == "bf16":
                fn, name = lib.gn_decoder_fwd_tc, "gn_decoder_fwd_tc"
                need = lib.gn_decoder_tc_workspace_bytes(agents, s, f_dim, z_dim, self.past_length, self.future_length)
            else:
                fn, name = lib.gn_decoder_fwd, "gn_decoder_fwd"
                need = lib.gn_decoder_workspace_bytes(agents, s, self.past_length)
            ws = self._ws.get(int(need), dev)
            with torch.cuda.device(dev):
                rc = fn(structs, len(self.decompose), C.c_void_p(pf.data_ptr()),
                        C.c_void_p(zz.data_ptr()), C.c_void_p(pt.data_ptr()), C.c_void_p(cl.data_ptr()),
                        agents, s, f_dim, z_dim, self.past_length, self.future_length,
                        C.c_void_p(out_seq.data_ptr()), C.c_void_p(recover.data_ptr()),
                        C.c_void_p(ws.data_ptr()), ws.numel(), ops._stream_ptr(dev))
            _lib.check(rc, name)
        if mode == 'inference':
            out_seq = out_seq.view(-1, s, *out_seq.shape[1:])
        return out_seq, recover
