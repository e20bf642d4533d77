"""The hot path exactly as the reference encoders drive it.

`PastEncoder.forward` (model/GroupNet_nba.py:284-309) computes the feature
correlation, runs one pairwise layer and 1-3 hyper layers on the SAME
`ftraj_input`, concatenates the incidences along dim 1 and the features along
dim -1.  `MultiScaleInteraction` is that block with the reference's attribute
names (`interaction`, `interaction_hyper`, `interaction_hyper2`,
`interaction_hyper3`), so a PastEncoder's sub-modules can be moved in and out
by name.  One fused kernel produces every H_s; each layer writes its slice of
the concatenated feature tensor directly (no torch.cat on the device path).

`forward_host` is the end-to-end entry bench.py times: host (pinned) features
in, host features + incidence out, H2D / compute / D2H overlapped over chunks
of scenes on three streams.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from .layers import MS_HGNN_hyper, MS_HGNN_oridinary
from .packing import RuntimeStateMixin

_HYPER_NAMES = ("interaction_hyper", "interaction_hyper2", "interaction_hyper3")

_cudart = None


def _loaded_cudart_path() -> str:
    """The CUDA runtime this process already uses (the one torch / libgroupnet_b200.so loaded), whatever its
    major version; opening a second runtime by a hard-coded soname would be wrong under another CUDA build."""
    try:
        with open("/proc/self/maps") as maps:
            for line in maps:
                path = line.rsplit(" ", 1)[-1].strip()
                if "libcudart" in path and ".so" in path:
                    return path
    except OSError:
        pass
    return "libcudart.so"


def _memcpy2d_d2h(dst_host: torch.Tensor, src_dev: torch.Tensor, col0: int, ncols: int, stream) -> None:
    """Copy columns [col0, col0+ncols) of every (b, n) row of a (m, N, W) fp32 device tensor into the
    same columns of a pinned host tensor of the same shape: one strided cudaMemcpy2DAsync."""
    global _cudart
    import ctypes
    if _cudart is None:
        _cudart = ctypes.CDLL(_loaded_cudart_path())
        _cudart.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                              ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        _cudart.cudaMemcpy2DAsync.restype = ctypes.c_int
    m, n, w = src_dev.shape
    assert dst_host.shape == src_dev.shape and dst_host.is_contiguous() and src_dev.is_contiguous()
    rc = _cudart.cudaMemcpy2DAsync(dst_host.data_ptr() + 4 * col0, 4 * w, src_dev.data_ptr() + 4 * col0, 4 * w,
                                   4 * ncols, m * n, 2, stream.cuda_stream)      # 2 = cudaMemcpyDeviceToHost
    if rc != 0:
        raise RuntimeError(f"cudaMemcpy2DAsync failed with cudaError {rc}")


class MultiScaleInteraction(RuntimeStateMixin, nn.Module):
    _RUNTIME_ATTRS = ("_stream_cache", "_buf_cache")

    def _reset_runtime(self) -> None:
        self.__dict__.pop("_stream_cache", None)
        self.__dict__.pop("_buf_cache", None)

    def __init__(self, model_dim: int = 64, hyper_scales: Sequence[int] = (5, 11), nmp_layers: int = 1):
        super().__init__()
        if len(hyper_scales) > 3:
            raise ValueError("the reference encoders take at most 3 hyper scales (GroupNet_nba.py:219-248)")
        self.model_dim = model_dim
        self.hyper_scales = [int(s) for s in hyper_scales]
        # construction order and arguments of PastEncoder.__init__ (model/GroupNet_nba.py:209-248)
        self.interaction = MS_HGNN_oridinary(embedding_dim=16, h_dim=model_dim, mlp_dim=64,
                                             bottleneck_dim=model_dim, batch_norm=0, nmp_layers=nmp_layers)
        for name, scale in zip(_HYPER_NAMES, self.hyper_scales):
            setattr(self, name, MS_HGNN_hyper(embedding_dim=model_dim, h_dim=model_dim, mlp_dim=64,
                                              bottleneck_dim=model_dim, batch_norm=0,
                                              nmp_layers=nmp_layers, scale=scale))

    # ------------------------------------------------------------------
    def layers(self) -> List[nn.Module]:
        return [self.interaction] + [getattr(self, n) for n in _HYPER_NAMES[:len(self.hyper_scales)]]

    def set_rng(self, mode: str, seed: int = 0, scene_offset: int = 0):
        for i, l in enumerate(self.layers()):
            l.set_rng(mode, seed + 1000003 * i, scene_offset)
        return self

    def set_precision(self, precision: str):
        for l in self.layers():
            l.set_precision(precision)
        return self

    def feature_width(self) -> int:
        return self.model_dim * (2 + len(self.hyper_scales))

    def incidence_rows(self, n: int) -> int:
        return sum(ops.incidence_rows(n, s) for s in self.hyper_scales)

    def launches_per_forward(self, batch: int, n: int) -> int:
        total = 1 if self.hyper_scales else 0                      # fused corr + top-k
        total += self.interaction.launches_per_forward(batch, n, n * n)
        for name, s in zip(_HYPER_NAMES, self.hyper_scales):
            total += getattr(self, name).launches_per_forward(batch, n, ops.incidence_rows(n, s))
        return total

    # ------------------------------------------------------------------
    def forward(self, ftraj_input: torch.Tensor, *, out_feature: Optional[torch.Tensor] = None,
                out_H: Optional[torch.Tensor] = None, noise=None,
                write_input_slice: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """ftraj_input (B,N,D) -> (final_feature (B,N,D*(2+S)), new_H (B,sum E_s,N))
        == model/GroupNet_nba.py:284-309 (before the final .view)."""
        ops._require_cuda_f32(ftraj_input, "ftraj_input")
        x = ftraj_input.contiguous()
        b, n, d = x.shape
        dev = x.device
        width = self.feature_width()
        if out_feature is None:
            out_feature = torch.empty(b, n, width, dtype=torch.float32, device=dev)
        new_h = None
        hs: List[torch.Tensor] = []
        if self.hyper_scales:
            rows = [ops.incidence_rows(n, s) for s in self.hyper_scales]
            if out_H is None:
                out_H = torch.empty(b, sum(rows), n, dtype=torch.float32, device=dev)
            hs = ops.corr_topk_h_into(x, self.hyper_scales, out_H)
            new_h = out_H
        if write_input_slice:
            out_feature[:, :, :d].copy_(x)                          # skip connection slice (:301-309)
        noise = list(noise) if noise is not None else [None] * (1 + len(self.hyper_scales))
        self.interaction(x, out=out_feature[:, :, d:2 * d], want_factors=False, noise=noise[0])
        for i, name in enumerate(_HYPER_NAMES[:len(self.hyper_scales)]):
            layer = getattr(self, name)
            # each layer reads its rows of new_H in place (scene stride = sum E * N)
            layer(x, H=hs[i], out=out_feature[:, :, (2 + i) * d:(3 + i) * d],
                  want_factors=False, noise=noise[1 + i])
        return out_feature, new_h

    # ------------------------------------------------------------------
    @torch.no_grad()
    def forward_host(self, x_host: torch.Tensor, out_feature_host: Optional[torch.Tensor] = None,
                     out_H_host: Optional[torch.Tensor] = None, chunk_scenes: int = 8192,
                     device: Optional[torch.device] = None, input_slice: str = "auto", sync: bool = True):
        """Host tensors in, host tensors out; copies and compute overlapped.

        x_host (B,N,D) fp32, ideally pinned.  Returns (final_feature, new_H) on
        the host (pinned).  With `sync=True` (default) the call returns when the
        outputs are complete on the host; `sync=False` returns as soon as the work
        is enqueued and hands back a third value, the CUDA event to wait on before
        reading them.  Scene i of chunk c gets Philox offset = global scene index,
        and with rng="cpu-compat" the uniforms of the WHOLE batch are drawn first,
        one torch.rand(B,E,T) per layer in the reference's order
        (model/MS_HGNN_batch.py:454), so the result does not depend on `chunk_scenes`."""
        if x_host.is_cuda:
            raise ValueError("forward_host takes host tensors; call forward() for device tensors")
        dev = device or next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("groupnet_b200: module must live on a CUDA device — no CPU fallback")
        b, n, d = x_host.shape
        width, rows = self.feature_width(), self.incidence_rows(n)
        # the x slice of final_feature: "host" fills it on the CPU (saves 25 % of the D2H bytes; right when the
        # process has cores to spare), "device" lets the GPU write it and copies whole rows back (one contiguous
        # D2H per chunk, no CPU work: right when many ranks share the host; measured on a 32-core box with 4 ranks:
        # 33.2 ms device vs 37.4 ms host, while a single rank with 16 threads prefers host, 14.7 vs 15.9 ms)
        if input_slice == "auto":
            input_slice = "host" if torch.get_num_threads() >= 16 else "device"
        if input_slice not in ("host", "device"):
            raise ValueError("input_slice must be 'auto', 'host' or 'device'")
        on_host = input_slice == "host"
        if out_feature_host is None:
            out_feature_host = torch.empty(b, n, width, dtype=torch.float32, pin_memory=True)
        if out_H_host is None and rows:
            out_H_host = torch.empty(b, rows, n, dtype=torch.float32, pin_memory=True)
        cs = max(1, min(chunk_scenes, b))
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            s_in, s_out = self._streams(dev)
            bufs = self._host_bufs(dev, cs, n, d, width, rows)
            ev_in = [torch.cuda.Event() for _ in range(2)]
            ev_cmp = [torch.cuda.Event() for _ in range(2)]
            ev_out = [torch.cuda.Event() for _ in range(2)]
            base_offsets = [l.scene_offset for l in self.layers()]
            base_calls = [l._philox_calls for l in self.layers()]   # the whole host batch is ONE forward
            start = torch.cuda.Event()
            start.record(main)
            s_in.wait_event(start)
            s_out.wait_event(start)
            nchunk = (b + cs - 1) // cs
            full_noise = None
            if nchunk > 1 and any(l.rng == "cpu-compat" for l in self.layers()):
                if any(l.nmp_layers > 1 for l in self.layers()):
                    raise ValueError('forward_host: rng="cpu-compat" with nmp_layers > 1 needs the whole batch in one '
                                     "chunk (chunk_scenes >= B) to keep the reference's draw order")
                full_noise = [torch.rand(b, e_l, l.edge_types).float() if l.rng == "cpu-compat" else None
                              for l, e_l in zip(self.layers(), [n * n] + [ops.incidence_rows(n, s_) for s_ in self.hyper_scales])]
            for c in range(nchunk):
                b0, b1 = c * cs, min(b, (c + 1) * cs)
                k, m = c & 1, b1 - b0
                xd, fd, hd = bufs[k]
                with torch.cuda.stream(s_in):
                    if c >= 2:
                        s_in.wait_event(ev_cmp[k])               # compute of chunk c-2 finished reading xd
                    xd[:m].copy_(x_host[b0:b1], non_blocking=True)
                    ev_in[k].record(s_in)
                main.wait_event(ev_in[k])
                if c >= 2:
                    main.wait_event(ev_out[k])                   # D2H of chunk c-2 finished reading fd / hd
                for l, off, calls in zip(self.layers(), base_offsets, base_calls):
                    l.scene_offset, l._philox_calls = off + b0, calls
                # the x slice of final_feature is already on the host: do not move it over PCIe twice
                self.forward(xd[:m], out_feature=fd[:m], out_H=hd[:m] if rows else None,
                             write_input_slice=not on_host,
                             noise=None if full_noise is None else [None if u is None else [u[b0:b1]] for u in full_noise])
                ev_cmp[k].record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[k])
                    if on_host:
                        _memcpy2d_d2h(out_feature_host[b0:b1], fd[:m], d, width - d, s_out)
                    else:
                        out_feature_host[b0:b1].copy_(fd[:m], non_blocking=True)
                    if rows:
                        out_H_host[b0:b1].copy_(hd[:m], non_blocking=True)
                    ev_out[k].record(s_out)
                if on_host:
                    out_feature_host[b0:b1, :, :d].copy_(x_host[b0:b1])   # host-side, overlaps the GPU work
            for l, off, calls in zip(self.layers(), base_offsets, base_calls):
                l.scene_offset, l._philox_calls = off, calls + 1
            main.wait_stream(s_out)
            done = torch.cuda.Event()
            done.record(main)
        if sync:
            done.synchronize()                                   # the D2H copies are non_blocking: wait for them here
            return out_feature_host, out_H_host
        return out_feature_host, out_H_host, done

    def _streams(self, dev):
        cache = self.__dict__.setdefault("_stream_cache", {})
        if dev not in cache:
            cache[dev] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        return cache[dev]

    def _host_bufs(self, dev, cs, n, d, width, rows):
        cache = self.__dict__.setdefault("_buf_cache", {})
        key = (str(dev), cs, n, d, width, rows)
        if key not in cache:
            cache.clear()
            cache[key] = [(torch.empty(cs, n, d, dtype=torch.float32, device=dev),
                           torch.empty(cs, n, width, dtype=torch.float32, device=dev),
                           torch.empty(cs, max(rows, 1), n, dtype=torch.float32, device=dev))
                          for _ in range(2)]
        return cache[key]
