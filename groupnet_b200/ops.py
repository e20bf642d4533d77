"""Python wrappers over the C ABI: device pointers in, torch tensors out.

PyTorch is used for device memory and streams only; every computation below
happens inside libgroupnet_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import StageCfg


def _require_cuda_f32(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"groupnet_b200: `{name}` must be a CUDA tensor — the B200 kernels have no CPU fallback")
    if t.dtype != torch.float32:
        # the reference fails the same way (init_adj builds FloatTensors, MS_HGNN_batch.py:148-149)
        raise RuntimeError(f"expected scalar type Float but found {str(t.dtype).replace('torch.', '').capitalize()}")


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def incidence_rows(n: int, scale: int) -> int:
    """E of the scale's incidence: 1 when scale == N (MS_HGNN_batch.py:375-377), else N."""
    return 1 if scale == n else n


def corr_topk_h(x: torch.Tensor, scales: Sequence[int], *, concat: bool = False,
                return_corr: bool = False):
    """Fused F.normalize + q q^T + per-scale top-k + incidence emission
    (model/GroupNet_nba.py:284-286, model/MS_HGNN_batch.py:372-388).

    Returns a list of H_s (B,E_s,N) tensors — views into one (B,sum E_s,N)
    tensor when `concat` (the layout the encoders build with torch.cat,
    model/GroupNet_nba.py:296,299) — and optionally corr (B,N,N)."""
    _require_cuda_f32(x, "x")
    lib = _lib.load()
    x = x.contiguous()
    b, n, d = x.shape
    scales = [int(s) for s in scales]
    for s in scales:
        if s > n:
            raise RuntimeError("selected index k out of range")
    rows = [incidence_rows(n, s) for s in scales]
    with torch.cuda.device(x.device):
        if concat:
            full = torch.empty(b, sum(rows), n, dtype=torch.float32, device=x.device)
            outs, off = [], 0
            for r in rows:
                outs.append(full[:, off:off + r, :])
                off += r
            strides = [sum(rows) * n] * len(scales)
        else:
            full = None
            outs = [torch.empty(b, r, n, dtype=torch.float32, device=x.device) for r in rows]
            strides = [r * n for r in rows]
        corr = torch.empty(b, n, n, dtype=torch.float32, device=x.device) if return_corr else None
        if b == 0:          # empty batch: nothing to launch (empty tensors have NULL data pointers)
            result = (outs, full) if concat else outs
            return (result, corr) if return_corr else result
        sc = (C.c_int32 * len(scales))(*scales)
        hp = (C.c_void_p * len(scales))(*[o.data_ptr() for o in outs])
        st = (C.c_int64 * len(scales))(*strides)
        rc = lib.gn_corr_topk_h(C.c_void_p(x.data_ptr()), b, n, d, sc, len(scales), hp, st,
                                C.c_void_p(corr.data_ptr() if corr is not None else 0),
                                _stream_ptr(x.device))
    _lib.check(rc, "gn_corr_topk_h")
    result = (outs, full) if concat else outs
    return (result, corr) if return_corr else result


def corr_topk_h_into(x: torch.Tensor, scales: Sequence[int], out_h: torch.Tensor) -> List[torch.Tensor]:
    """Same kernel, writing every H_s into a caller-owned concatenated
    (B, sum E_s, N) tensor; returns the per-scale views."""
    _require_cuda_f32(x, "x")
    _require_cuda_f32(out_h, "out_H")
    lib = _lib.load()
    b, n, d = x.shape
    scales = [int(s) for s in scales]
    for s in scales:
        if s > n:
            raise RuntimeError("selected index k out of range")
    rows = [incidence_rows(n, s) for s in scales]
    if tuple(out_h.shape) != (b, sum(rows), n) or not out_h.is_contiguous() or not x.is_contiguous():
        raise ValueError("out_H must be a contiguous (B, sum E_s, N) tensor and x contiguous")
    views, off = [], 0
    for r in rows:
        views.append(out_h[:, off:off + r, :])
        off += r
    if b == 0:
        return views
    with torch.cuda.device(x.device):
        sc = (C.c_int32 * len(scales))(*scales)
        hp = (C.c_void_p * len(scales))(*[v.data_ptr() for v in views])
        st = (C.c_int64 * len(scales))(*([sum(rows) * n] * len(scales)))
        rc = lib.gn_corr_topk_h(C.c_void_p(x.data_ptr()), b, n, d, sc, len(scales), hp, st,
                                C.c_void_p(0), _stream_ptr(x.device))
    _lib.check(rc, "gn_corr_topk_h")
    return views


def topk_h(corr: torch.Tensor, scale: int) -> torch.Tensor:
    """init_adj_attention (model/MS_HGNN_batch.py:372-388) on a supplied corr (B,N,N)."""
    _require_cuda_f32(corr, "corr")
    lib = _lib.load()
    corr = corr.contiguous()
    b, n, n2 = corr.shape
    if n != n2:
        raise ValueError("corr must be (B,N,N)")
    scale = int(scale)
    if scale > n:
        raise RuntimeError("selected index k out of range")
    e = incidence_rows(n, scale)
    with torch.cuda.device(corr.device):
        h = torch.empty(b, e, n, dtype=torch.float32, device=corr.device)
        if b == 0:
            return h
        rc = lib.gn_topk_h(C.c_void_p(corr.data_ptr()), b, n, scale, C.c_void_p(h.data_ptr()),
                           e * n, _stream_ptr(corr.device))
    _lib.check(rc, "gn_topk_h")
    return h


class Workspace:
    """Grow-only device scratch owned by a layer (the library never allocates)."""

    def __init__(self) -> None:
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int, device: torch.device) -> torch.Tensor:
        if self.buf is None or self.buf.device != device or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        return self.buf


def stage_workspace_bytes(cfg: StageCfg) -> int:
    return int(_lib.load().gn_stage_workspace_bytes(C.byref(cfg)))


def stage_forward(cfg: StageCfg, weights, h_in: torch.Tensor, inc: Optional[torch.Tensor],
                  u: Optional[torch.Tensor], node_out: torch.Tensor,
                  dist_out: Optional[torch.Tensor], ws: torch.Tensor) -> None:
    """One gn_stage_fwd call on the current stream of h_in's device."""
    lib = _lib.load()

    def ptr(t):
        return C.c_void_p(t.data_ptr() if t is not None else 0)

    with torch.cuda.device(h_in.device):
        rc = lib.gn_stage_fwd(C.byref(cfg), C.byref(weights.struct), ptr(h_in), ptr(inc), ptr(u),
                              ptr(node_out), ptr(dist_out), ptr(ws), ws.numel(),
                              _stream_ptr(h_in.device))
    _lib.check(rc, "gn_stage_fwd")


def stage_launch_count(cfg: StageCfg) -> int:
    return int(_lib.load().gn_stage_launch_count(C.byref(cfg)))
