"""Autograd for one message-passing stage: forward = gn_stage_fwd (fp32 kernels), backward =
gn_stage_bwd (hand-written CUDA, csrc/gn_train.cu).  Gradients flow to `h_states` and to every
parameter the reference forward uses; `H`, `corr` and the Gumbel noise get none, and the
never-used parameters (`edge_aggregation.mlp`, `spatial_embedding`, `spatial_transform`) keep
`grad = None`, exactly like the reference under torch autograd.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib, ops
from .packing import stage_modules


def stage_param_list(layer, s: int) -> List[torch.nn.Parameter]:
    """Parameters used by stage `s`, in the fixed order the autograd Function returns their grads."""
    dict_mod, post_mod = stage_modules(layer, s)
    node = layer.node2edge_start_mlp[s].layers
    att = layer.attention_mlp[s].layers
    agg = layer.edge_aggregation_list[s].agg_mlp
    out = [node[0].weight, node[0].bias, node[1].weight, node[1].bias,
           att[0].weight, att[0].bias, att[1].weight, att[1].bias]
    for m in (dict_mod.init_MLP, dict_mod.MLP_distribution, dict_mod.MLP_factor):
        out += [m.layers[0].weight, m.layers[0].bias, m.layers[1].weight, m.layers[1].bias]
    for m in agg:
        out += [m.layers[0].weight, m.layers[0].bias, m.layers[1].weight, m.layers[1].bias]
    out += [post_mod.layers[0].weight, post_mod.layers[0].bias, post_mod.layers[1].weight, post_mod.layers[1].bias]
    return out


def _lin(w, b, dw, db) -> _lib.Lin:
    l = _lib.Lin()
    l.W = C.c_void_p(w.data_ptr())
    l.b = C.c_void_p(b.data_ptr() if b is not None else 0)
    l.dW = C.c_void_p(dw.data_ptr() if dw is not None else 0)
    l.db = C.c_void_p(db.data_ptr() if db is not None else 0)
    l.N, l.K = int(w.shape[0]), int(w.shape[1])
    return l


class StageFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, inc, u, meta, weights, *params):
        cfg = _lib.StageCfg()
        (cfg.B, cfg.N, cfg.D, cfg.Dout, cfg.E, cfg.T, cfg.pairwise, cfg.noise_mode, cfg.stage_index,
         cfg.seed, cfg.scene_offset) = meta
        cfg.precision = _lib.GN_FP32
        cfg.h_stride = 0 if inc is None else inc.stride(0)
        dev = h.device
        h = h.contiguous()
        ws = torch.empty(max(ops.stage_workspace_bytes(cfg), 256), dtype=torch.uint8, device=dev)
        node_out = torch.empty(cfg.B, cfg.N, cfg.Dout, dtype=torch.float32, device=dev)
        dist = torch.empty(cfg.B, cfg.E, cfg.T, dtype=torch.float32, device=dev)
        ops.stage_forward(cfg, weights, h, inc, u, node_out, dist, ws)
        ctx.meta = meta
        ctx.h_stride = cfg.h_stride
        ctx.has_inc = inc is not None
        ctx.save_for_backward(h, inc if inc is not None else h.new_empty(0), ws, *params)
        return node_out, dist

    @staticmethod
    def backward(ctx, d_out, d_dist):
        lib = _lib.load()
        h, inc, ws, *params = ctx.saved_tensors
        cfg = _lib.StageCfg()
        (cfg.B, cfg.N, cfg.D, cfg.Dout, cfg.E, cfg.T, cfg.pairwise, cfg.noise_mode, cfg.stage_index,
         cfg.seed, cfg.scene_offset) = ctx.meta
        cfg.precision = _lib.GN_FP32
        cfg.h_stride = ctx.h_stride
        dev = h.device
        t = cfg.T
        d_out = d_out.contiguous() if d_out is not None else torch.zeros(cfg.B, cfg.N, cfg.Dout, device=dev)
        d_dist = d_dist.contiguous() if d_dist is not None else None
        # wgrad kernels ACCUMULATE into their dW / db buffers.  A parameter whose .grad already exists as a dense fp32
        # tensor (e.g. a view into ddp.FlatGradBucket's arena) takes its gradient in place and autograd gets None for
        # it — no per-parameter allocation, memset, accumulation or bucket copy; all others get fresh zeroed buffers
        # carved out of ONE allocation
        inplace = [p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous()
                   and p.grad.device == p.device and p.grad.shape == p.shape for p in params]
        fresh = torch.zeros(sum(p.numel() for p, ip in zip(params, inplace) if not ip), dtype=torch.float32, device=dev)
        grads, off = [], 0
        for p, ip in zip(params, inplace):
            if ip:
                grads.append(p.grad)
            else:
                grads.append(fresh[off:off + p.numel()].view(p.shape))
                off += p.numel()
        it = iter(range(len(params)))

        def nxt():
            i = next(it)
            return params[i], grads[i]

        tp = _lib.TrainParams()
        (w, dw), (b, db) = nxt(), nxt(); tp.node0 = _lin(w, b, dw, db)
        (w, dw), (b, db) = nxt(), nxt(); tp.node1 = _lin(w, b, dw, db)
        (aw0, daw0), (ab0, dab0), (aw1, daw1), (ab1, dab1) = nxt(), nxt(), nxt(), nxt()
        wpq = torch.cat((aw0[:, :64], aw0[:, 64:]), dim=0).contiguous()          # (64, 64): pn rows | q rows
        dwpq = torch.zeros_like(wpq)
        tp.attpq = _lin(wpq, None, dwpq, None)
        aw1c = aw1.contiguous()
        tp.att_b0, tp.att_w1, tp.att_b1 = (C.c_void_p(ab0.data_ptr()), C.c_void_p(aw1c.data_ptr()),
                                           C.c_void_p(ab1.data_ptr()))
        tp.d_att_b0, tp.d_att_w1, tp.d_att_b1 = (C.c_void_p(dab0.data_ptr()), C.c_void_p(daw1.data_ptr()),
                                                 C.c_void_p(dab1.data_ptr()))
        for name in ("init", "dist", "fac"):
            (w, dw), (b, db) = nxt(), nxt(); setattr(tp, name + "0", _lin(w, b, dw, db))
            (w, dw), (b, db) = nxt(), nxt(); setattr(tp, name + "1", _lin(w, b, dw, db))
        for k in range(t):
            (w, dw), (b, db) = nxt(), nxt(); tp.agg0[k] = _lin(w, b, dw, db)
            (w, dw), (b, db) = nxt(), nxt(); tp.agg1[k] = _lin(w, b, dw, db)
        (w, dw), (b, db) = nxt(), nxt(); tp.post0 = _lin(w, b, dw, db)
        (w, dw), (b, db) = nxt(), nxt(); tp.post1 = _lin(w, b, dw, db)

        d_h = torch.empty_like(h)
        bws = torch.empty(max(int(lib.gn_stage_bwd_workspace_bytes(C.byref(cfg))), 256), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.gn_stage_bwd(C.byref(cfg), C.byref(tp), C.c_void_p(h.data_ptr()),
                                  C.c_void_p(inc.data_ptr() if ctx.has_inc else 0), C.c_void_p(ws.data_ptr()),
                                  C.c_void_p(d_out.data_ptr()), cfg.Dout,
                                  C.c_void_p(d_dist.data_ptr() if d_dist is not None else 0),
                                  C.c_void_p(d_h.data_ptr()), C.c_void_p(bws.data_ptr()), bws.numel(),
                                  ops._stream_ptr(dev))
        _lib.check(rc, "gn_stage_bwd")
        grads[4].add_(torch.cat((dwpq[:32], dwpq[32:]), dim=1))                   # attention layers.0.weight
        return (d_h, None, None, None, None, *[None if ip else g for g, ip in zip(grads, inplace)])
