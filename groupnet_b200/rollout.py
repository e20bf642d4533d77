"""Graph-replayed `PastEncoder` forward for rollouts at a fixed small batch.

The reference's simulator calls the model thousands of times at batch 1 (Simulator.py:231-238,332-334); at that
size a forward is launch-bound (16 kernels; ~520 us of host time launched eagerly against ~170 us of device
time on a B200).  `libgroupnet_b200.so` allocates nothing and launches only on the caller's stream, so one
forward is captured once into a CUDA graph and replayed per call.

RNG contract (model/MS_HGNN_batch.py:454): every call still draws one `torch.rand(B, E, T)` per
`MLP_dict_softmax` from the global CPU generator, in module call order (pairwise, then the hyper scales) —
the draws land in static device buffers the captured kernels read, so a replay sees fresh noise and the
results are bit-identical to the eager `PastEncoder.forward` under the same `torch.manual_seed`.
`rng="philox"` instead keeps the whole step on the device: the layers run in their "philox-device" mode, where
the per-call seed lives in device memory (GN_NOISE_PHILOX_DEVICE_SEED) and the captured graph itself advances
it, so replay k is bit-identical to the k-th eager forward under `set_rng("philox", seed)` with no host work
per call beyond the input copy and the graph launch.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops
from .encoder import PastEncoder


class GraphedPastEncoder:
    """encoder: a `groupnet_b200.PastEncoder` on a CUDA device, in eval mode.

        g = GraphedPastEncoder(encoder, batch_size=1, agent_num=11, length=5)
        feature, new_H = g(inputs)          # inputs (B*N, T, in_dim), host or device

    The returned tensors are the graph's static outputs: they are overwritten by the next call
    (pass `clone=True` to get private copies).  Weights are read in place: after an optimizer step or
    `load_state_dict` call `recapture()`."""

    def __init__(self, encoder: PastEncoder, batch_size: int, agent_num: int, length: int, in_dim: int = 4,
                 rng: str = "cpu-compat", seed: int = 0):
        if rng not in ("cpu-compat", "philox"):
            raise ValueError(rng)
        self.rng, self.seed = rng, int(seed)
        dev = next(encoder.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedPastEncoder needs the encoder on a CUDA device")
        if encoder.training:
            raise RuntimeError("GraphedPastEncoder is eval-only (PastEncoder's fused front-end is)")
        self.encoder, self.device = encoder, dev
        self.batch_size, self.agent_num = int(batch_size), int(agent_num)
        b, n = self.batch_size, self.agent_num
        self._x = torch.zeros(b * n, length, in_dim, dtype=torch.float32, device=dev)
        self._shapes: List[List[Tuple[int, int, int]]] = []
        for i, layer in enumerate(encoder.layers()):
            e = n * n if i == 0 else ops.incidence_rows(n, layer.scale)
            self._shapes.append([(b, e, layer.edge_types)] * max(layer.nmp_layers, 1))
        # one flat pinned buffer and one flat device buffer for all draws of a call: one H2D copy per call
        total = sum(s[0] * s[1] * s[2] for per in self._shapes for s in per)
        self._u_host_flat = torch.zeros(total, dtype=torch.float32).pin_memory()
        self._u_dev_flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self._u_host, self._u_dev, off = [], [], 0
        for per in self._shapes:
            hs, ds = [], []
            for s in per:
                cnt = s[0] * s[1] * s[2]
                hs.append(self._u_host_flat[off:off + cnt].view(s))
                ds.append(self._u_dev_flat[off:off + cnt].view(s))
                off += cnt
            self._u_host.append(hs)
            self._u_dev.append(ds)
        self._copied = torch.cuda.Event()
        self._graph = None
        self.recapture()

    def _forward(self):
        return self.encoder(self._x, self.batch_size, self.agent_num,
                            noise=self._u_dev if self.rng == "cpu-compat" else None)

    @torch.no_grad()
    def recapture(self) -> None:
        block = self.encoder._interaction_block()
        if self.rng == "philox":
            block.set_rng("philox-device", self.seed)     # NOTE: changes the rng mode of the encoder's layers
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):               # warm-up outside the capture: packs weights, sizes the workspaces
            self._forward()
        cur.wait_stream(side)
        # the captured kernels hold raw pointers: keep what they point at alive even if the layers later swap in
        # a larger workspace (bigger eager batch) or re-packed weights (optimizer step)
        self._keepalive = [(l._ws.buf, l._packs._stages) for l in self.encoder.layers()]
        self._keepalive.append(self.encoder.folded_frontend(self.agent_num, self._x.shape[1], self.device))
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._feature, self._new_h = self._forward()
        if self.rng == "philox":
            block.set_rng("philox-device", self.seed)     # rewind the device seeds (in place): replay 0 == eager call 0
            self._keepalive.append([l._seed_dev for l in self.encoder.layers()])
        self._copied.record(cur)

    @torch.no_grad()
    def __call__(self, inputs: torch.Tensor, clone: bool = False):
        if tuple(inputs.shape) != tuple(self._x.shape):
            raise RuntimeError(f"inputs must be {tuple(self._x.shape)}, got {tuple(inputs.shape)}")
        if inputs.dtype != torch.float32:
            raise RuntimeError("expected scalar type Float")
        if self.rng == "cpu-compat":
            self._copied.synchronize()              # the previous call's H2D copies have left the pinned buffers
            for per_h in self._u_host:
                for uh in per_h:
                    torch.rand(uh.shape, out=uh)    # global CPU generator, module call order (:454)
            self._u_dev_flat.copy_(self._u_host_flat, non_blocking=True)
        self._x.copy_(inputs, non_blocking=True)
        if self.rng == "cpu-compat":
            self._copied.record(torch.cuda.current_stream(self.device))
        self._graph.replay()
        if clone:
            return self._feature.clone(), self._new_h.clone()
        return self._feature, self._new_h
