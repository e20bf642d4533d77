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



class GraphedInference(GraphedPastEncoder):
    """The whole of `GroupNet.inference_simulator` (model/GroupNet_nba.py:830-869) — velocities, `PastEncoder`, the
    prior sample z ~ N(0, I) for `sample_k` futures per agent, `Decoder`, the final permute — captured ONCE into a CUDA
    graph and replayed per call: the rollout loops of the reference (Simulator.py:231-238,332-334) call it thousands of
    times at batch 1, where the eager path is bound by ~20 kernel launches and the small torch ops between them.

        g = GraphedInference(past_encoder, decoder, batch_size=1, agent_num=11)
        diverse_pred_traj, H = g(data)      # data (B, N, past_length, 2) -> (sample_k, B*N, future_length, 2), (B, sum E, N)

    RNG contract of the reference, kept with rng="cpu-compat": per call one `torch.rand(B, E, T)` per
    `MLP_dict_softmax` (pairwise, then the hyper scales, MS_HGNN_batch.py:454) and then one
    `torch.randn(B*N*sample_k, zdim)` (`Normal.fixed_eps`, GroupNet_nba.py:93) from the global CPU generator, in that
    order, into pinned buffers the captured kernels read.  rng="philox": Gumbel noise from the device-resident
    Philox seed and z from torch's CUDA generator inside the graph — no host work per call beyond the input copy.
    Only the reference's default prior (`learn_prior=False`: mu = 0, logvar = 0) is built."""

    def __init__(self, encoder: PastEncoder, decoder, batch_size: int, agent_num: int, sample_k: int = 20,
                 zdim: int = 32, rng: str = "cpu-compat", seed: int = 0):
        dev = next(encoder.parameters()).device
        self.decoder = decoder
        self.sample_k, self.zdim = int(sample_k), int(zdim)
        length = int(decoder.past_length)
        rows = int(batch_size) * int(agent_num) * self.sample_k
        self._data = torch.zeros(int(batch_size), int(agent_num), length, 2, dtype=torch.float32, device=dev)
        self._eps_host = torch.zeros(rows, self.zdim, dtype=torch.float32).pin_memory()
        self._eps_dev = torch.zeros(rows, self.zdim, dtype=torch.float32, device=dev)
        super().__init__(encoder, batch_size, agent_num, length, in_dim=4, rng=rng, seed=seed)

    @torch.no_grad()
    def recapture(self) -> None:
        super().recapture()
        self._keepalive.append((self.decoder._packed, self.decoder._ws.buf))

    def _forward(self):
        b, n = self.batch_size, self.agent_num
        past_traj = self._data.reshape(b * n, self._data.shape[2], 2)
        past_vel = past_traj[:, 1:] - past_traj[:, :-1, :]
        past_vel = torch.cat([past_vel[:, :1], past_vel], dim=1)             # :835-836 (slices: an index list would be an
        cur_location = past_traj[:, -1:]                                     #  H2D copy, which a graph capture forbids)
        inputs = torch.cat((past_traj, past_vel), dim=-1)
        feature, new_h = self.encoder(inputs, b, n, noise=self._u_dev if self.rng == "cpu-compat" else None)
        feature_rep = feature.repeat_interleave(self.sample_k, dim=0)          # :848
        if self.rng == "cpu-compat":
            z = self._eps_dev                                                  # mu + eps * sigma with mu = 0, sigma = 1 (:849-856)
        else:
            z = torch.randn(feature_rep.shape[0], self.zdim, dtype=torch.float32, device=self.device)
        pred, _ = self.decoder(feature_rep, z, b, n, past_traj.contiguous(), cur_location.contiguous(),
                               sample_num=self.sample_k, mode='inference')
        return pred.permute(1, 0, 2, 3), new_h                                 # :867

    @torch.no_grad()
    def __call__(self, data: torch.Tensor, clone: bool = False):
        if tuple(data.shape) != tuple(self._data.shape):
            raise RuntimeError(f"data must be {tuple(self._data.shape)}, got {tuple(data.shape)}")
        if data.dtype != torch.float32:
            raise RuntimeError("expected scalar type Float")
        if self.rng == "cpu-compat":
            self._copied.synchronize()
            for per_h in self._u_host:
                for uh in per_h:
                    torch.rand(uh.shape, out=uh)
            torch.randn(self._eps_host.shape, out=self._eps_host)              # Normal.fixed_eps (:93), after the encoder's draws
            self._u_dev_flat.copy_(self._u_host_flat, non_blocking=True)
            self._eps_dev.copy_(self._eps_host, non_blocking=True)
        self._data.copy_(data, non_blocking=True)
        if self.rng == "cpu-compat":
            self._copied.record(torch.cuda.current_stream(self.device))
        self._graph.replay()
        if clone:
            return self._feature.clone(), self._new_h.clone()
        return self._feature, self._new_h


def inference_simulator(encoder: PastEncoder, decoder, data: torch.Tensor, sample_k: int = 20, zdim: int = 32):
    """Eager form of the same pipeline (one call of `GroupNet.inference_simulator`, default prior), drawing the noise
    from the global CPU generator in the reference's order.  data (B, N, past_length, 2) on the model's device."""
    b, n = data.shape[:2]
    past_traj = data.reshape(b * n, data.shape[2], 2).contiguous()
    past_vel = past_traj[:, 1:] - past_traj[:, :-1, :]
    past_vel = torch.cat([past_vel[:, [0]], past_vel], dim=1)
    cur_location = past_traj[:, [-1]].contiguous()
    with torch.no_grad():
        feature, new_h = encoder(torch.cat((past_traj, past_vel), dim=-1), b, n)
        feature_rep = feature.repeat_interleave(sample_k, dim=0)
        z = torch.randn(feature_rep.shape[0], zdim).to(data.device)
        pred, _ = decoder(feature_rep, z, b, n, past_traj, cur_location, sample_num=sample_k, mode='inference')
    return pred.permute(1, 0, 2, 3), new_h
