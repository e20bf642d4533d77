"""Drop-in `MS_HGNN_oridinary` / `MS_HGNN_hyper` for TaliMotzkin/GroupNet.

Same constructor and forward signatures, same sub-module tree (hence the same
state_dict keys, shapes and parameter registration order) as
`model/MS_HGNN_batch.py:55-198` and `:270-443`, so the classes can be swapped
into `PastEncoder` / `FutureEncoder` (`model/GroupNet_nba.py:209-248,329-371`)
and `load_state_dict(strict=True)` keeps working.  The forward bodies call the
sm_100a kernels of libgroupnet_b200.so through the C ABI; there is no PyTorch
or CPU implementation of the math in this package.

Differences a caller can observe (all documented in DESIGN.md):
  * inputs must be CUDA fp32 tensors (the reference is CPU-only);
  * top-k ties are broken towards the lower agent index (the reference inherits
    whatever torch.topk does);
  * Gumbel noise: `rng="cpu-compat"` (default) draws `torch.rand(B,E,T)` from
    the global CPU generator in the reference's order, `rng="philox"` draws on
    the device, `rng="philox-device"` is the same stream with the per-call seed
    kept (and advanced) in device memory so a captured CUDA graph can be
    replayed, `forward(..., noise=[U])` injects the uniforms.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from .packing import PackCache, RuntimeStateMixin

_MASK64 = (1 << 64) - 1
# "fp32": FFMA kernels (1e-5); "tf32": the same 1e-5 contract on tensor cores (3xTF32 tcgen05 chains);
# "bf16": bf16 tcgen05 operands, fp32 accumulation (2e-2)
_PRECISION = {"fp32": _lib.GN_FP32, "bf16": _lib.GN_BF16_TC, "tf32": _lib.GN_TF32X3}
_PHILOX_CALL_STRIDE = 0x9E3779B97F4A7C15        # seed of call k = philox_seed + k * stride (mod 2^64)


def _as_int64(v: int) -> int:
    v &= _MASK64
    return v - (1 << 64) if v >= (1 << 63) else v


# --------------------------------------------------------------------------
# module-level helpers of the reference that nothing on the path calls: kept importable
# under their names with the reference's behaviour; plain tensor utilities, no kernels involved
# --------------------------------------------------------------------------
def encode_onehot(labels):
    """One-hot rows in first-seen class order of `set(labels)` (model/MS_HGNN_batch.py:9-15)."""
    import numpy as np
    classes = set(labels)
    classes_dict = {c: np.identity(len(classes))[i, :] for i, c in enumerate(classes)}
    return np.array(list(map(classes_dict.get, labels)), dtype=np.int32)


def make_mlp(dim_list, activation='relu', batch_norm=True, dropout=0):
    """nn.Sequential of Linear (+BatchNorm1d) (+ReLU / LeakyReLU) (+Dropout) blocks (model/MS_HGNN_batch.py:17-29)."""
    blocks = []
    for dim_in, dim_out in zip(dim_list[:-1], dim_list[1:]):
        blocks.append(nn.Linear(dim_in, dim_out))
        if batch_norm:
            blocks.append(nn.BatchNorm1d(dim_out))
        if activation == 'relu':
            blocks.append(nn.ReLU())
        elif activation == 'leakyrelu':
            blocks.append(nn.LeakyReLU())
        if dropout > 0:
            blocks.append(nn.Dropout(p=dropout))
    return nn.Sequential(*blocks)


def sample_gumbel(shape, eps=1e-10):
    """-log(eps - log(U + eps)), U ~ torch.rand(shape) on the CPU generator (model/MS_HGNN_batch.py:446-455)."""
    uniform = torch.rand(shape).float()
    return -torch.log(eps - torch.log(uniform + eps))


def my_softmax(input, axis=1):
    """Softmax over `axis` through a transpose, as model/MS_HGNN_batch.py:517-520 computes it for the 3-D
    tensors the layers pass (for those the reference's implicit-dim softmax acts on dim 0 of the transposed
    tensor, i.e. on `axis`)."""
    trans_input = input.transpose(axis, 0).contiguous()
    soft_max_1d = torch.softmax(trans_input, dim=0 if trans_input.dim() in (0, 1, 3) else 1)
    return soft_max_1d.transpose(axis, 0)


def gumbel_softmax_sample(logits, tau=1, eps=1e-10):
    """softmax((logits + g) / tau) over the last dim (model/MS_HGNN_batch.py:458-473)."""
    gumbel_noise = sample_gumbel(logits.size(), eps=eps).to(logits.device)
    return my_softmax((logits + gumbel_noise) / tau, axis=-1)


def gumbel_softmax(logits, tau=1, hard=False, eps=1e-10):
    """Gumbel-softmax sample, optionally straight-through one-hot (model/MS_HGNN_batch.py:475-515)."""
    y_soft = gumbel_softmax_sample(logits, tau=tau, eps=eps)
    if not hard:
        return y_soft
    _, k = y_soft.data.max(-1)
    y_hard = torch.zeros(*logits.size(), device=logits.device)
    y_hard = y_hard.scatter_(-1, k.view(tuple(logits.size()[:-1]) + (1,)), 1.0)
    return (y_hard - y_soft.data) + y_soft


# --------------------------------------------------------------------------
# parameter containers (schema of MS_HGNN_batch.py:201-268)
# --------------------------------------------------------------------------
class MLP(nn.Module):
    """Linear stack (`MLP`, :201-229).  Container only: the layer kernels read
    `layers[i].weight/bias` directly."""

    def __init__(self, input_dim, output_dim, hidden_size=(1024, 512), activation='relu',
                 discrim=False, dropout=-1):
        super().__init__()
        widths = [input_dim, *hidden_size, output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        if activation == 'relu':
            self.activation = nn.ReLU()
        elif activation == 'sigmoid':
            self.activation = nn.Sigmoid()
        self.sigmoid = nn.Sigmoid() if discrim else None
        self.dropout = dropout

    def forward(self, x):
        raise RuntimeError("groupnet_b200.MLP is a parameter container; its math runs inside "
                           "the fused stage kernels (gn_stage_fwd)")


class MLP_dict_softmax(nn.Module):
    """Interaction-category / strength head (`MLP_dict_softmax`, :31-39)."""

    def __init__(self, input_dim, output_dim, hidden_size=(1024, 512), activation='relu',
                 discrim=False, dropout=-1, edge_types=5):
        super().__init__()
        self.bottleneck_dim = edge_types
        self.MLP_distribution = MLP(input_dim=input_dim, output_dim=edge_types, hidden_size=hidden_size)
        self.MLP_factor = MLP(input_dim=input_dim, output_dim=1, hidden_size=hidden_size)
        self.init_MLP = MLP(input_dim=input_dim, output_dim=input_dim, hidden_size=hidden_size)

    def forward(self, x):
        raise RuntimeError("groupnet_b200.MLP_dict_softmax is a parameter container (gn_stage_fwd)")


class MLP_dict(MLP_dict_softmax):
    """Unused helper of the reference (:232-245); kept importable."""


class edge_aggregation(nn.Module):
    """Per-edge-type aggregation MLPs (`edge_aggregation`, :247-257).  `mlp` is
    never used by the reference forward but is part of its state_dict."""

    def __init__(self, input_dim, output_dim, hidden_size=(1024, 512), activation='relu',
                 discrim=False, dropout=-1, edge_types=5):
        super().__init__()
        self.edge_types = edge_types
        self.dict_dim = input_dim
        self.agg_mlp = nn.ModuleList(
            MLP(input_dim=input_dim, output_dim=input_dim, hidden_size=(128,)) for _ in range(edge_types))
        self.mlp = MLP(input_dim=input_dim, output_dim=input_dim, hidden_size=(128,))

    def forward(self, *a, **k):
        raise RuntimeError("groupnet_b200.edge_aggregation is a parameter container (gn_stage_fwd)")


# --------------------------------------------------------------------------
# shared machinery of both layers
# --------------------------------------------------------------------------
class _MessagePassingLayer(RuntimeStateMixin, nn.Module):
    edge_types: int
    _pairwise: bool
    _RUNTIME_ATTRS = ("_packs", "_ws", "_seed_dev")

    def _reset_runtime(self) -> None:
        self.__dict__["_packs"] = PackCache()
        self.__dict__["_ws"] = ops.Workspace()
        self.__dict__.pop("_seed_dev", None)

    def _build_tree(self, h_dim, bottleneck_dim, nmp_layers):
        """Sub-module tree in the reference's registration order (:75-91 / :296-311)."""
        ext = self.hdim_extend
        t = self.edge_types
        self.nmp_mlp_start = MLP_dict_softmax(input_dim=ext, output_dim=h_dim, hidden_size=(128,), edge_types=t)
        self.nmp_mlps = self.make_nmp_mlp()
        self.nmp_mlp_end = MLP(input_dim=h_dim * 2, output_dim=bottleneck_dim, hidden_size=(128,))
        self.attention_mlp = nn.ModuleList(
            MLP(input_dim=ext * 2, output_dim=1, hidden_size=(32,)) for _ in range(nmp_layers))
        self.node2edge_start_mlp = nn.ModuleList(
            MLP(input_dim=h_dim, output_dim=ext, hidden_size=(256,)) for _ in range(nmp_layers))
        self.edge_aggregation_list = nn.ModuleList(
            edge_aggregation(input_dim=h_dim, output_dim=bottleneck_dim, hidden_size=(128,), edge_types=t)
            for _ in range(nmp_layers))
        # runtime (not part of the state_dict, dropped by deepcopy / pickle: RuntimeStateMixin)
        self._reset_runtime()
        self._install_runtime_hooks()
        self.rng = "cpu-compat"
        self.precision = "fp32"
        self.philox_seed = 0
        self.scene_offset = 0
        self._philox_calls = 0
        self.workspace_limit_bytes = 4 << 30

    def make_nmp_mlp(self):
        mods = []
        for _ in range(self.nmp_layers - 1):
            mods.append(MLP(input_dim=self.h_dim * 2, output_dim=self.h_dim, hidden_size=(128,)))
            mods.append(MLP_dict_softmax(input_dim=self.hdim_extend, output_dim=self.h_dim,
                                         hidden_size=(128,), edge_types=self.edge_types))
        return nn.ModuleList(mods)

    # ---- configuration helpers -------------------------------------------------
    def set_rng(self, mode: str, seed: int = 0, scene_offset: int = 0):
        """mode in {"cpu-compat", "philox"}; `scene_offset` is the global index of
        this shard's first scene (results are then independent of the sharding)."""
        if mode not in ("cpu-compat", "philox", "philox-device"):
            raise ValueError(mode)
        self.rng, self.philox_seed, self.scene_offset, self._philox_calls = mode, int(seed), int(scene_offset), 0
        seed_dev = self.__dict__.get("_seed_dev")
        if seed_dev is not None:
            seed_dev.fill_(_as_int64(self.philox_seed))      # in place: a captured graph keeps reading this buffer
        return self

    def _device_seed(self, device) -> torch.Tensor:
        """One int64 in device memory holding the Philox seed of the NEXT forward (rng="philox-device")."""
        t = self.__dict__.get("_seed_dev")
        if t is None or t.device != device:
            t = torch.full((1,), _as_int64(self.philox_seed + _PHILOX_CALL_STRIDE * self._philox_calls),
                           dtype=torch.int64, device=device)
            self.__dict__["_seed_dev"] = t
        return t

    def set_precision(self, precision: str):
        if precision not in ("fp32", "bf16", "tf32"):
            raise ValueError(precision)
        self.precision = precision
        return self

    # ---- forward plumbing ---------------------------------------------------------
    def _noise_list(self, noise, batch, e, n_stage, device) -> Optional[List[torch.Tensor]]:
        t = self.edge_types
        if noise is not None:
            lst = [noise] if torch.is_tensor(noise) else list(noise)
            if len(lst) != n_stage:
                raise ValueError(f"need {n_stage} uniform tensors (one per MLP_dict_softmax call), got {len(lst)}")
            out = []
            for u in lst:
                if tuple(u.shape) != (batch, e, t):
                    raise ValueError(f"noise must be {(batch, e, t)}, got {tuple(u.shape)}")
                out.append(u.to(device=device, dtype=torch.float32, non_blocking=True).contiguous())
            return out
        if self.rng == "cpu-compat":
            # one torch.rand(B,E,T) per MLP_dict_softmax call, global CPU generator (:454)
            return [torch.rand(batch, e, t).float().to(device, non_blocking=True) for _ in range(n_stage)]
        return None

    def _run(self, h_states: torch.Tensor, inc: Optional[torch.Tensor], e: int, noise,
             node_out: Optional[torch.Tensor] = None, dist_out: Optional[torch.Tensor] = None,
             want_dist: bool = True):
        """Chain the L stages through gn_stage_fwd.  `node_out` may be a view whose
        rows are further apart than bottleneck_dim (a slice of a concatenated
        feature tensor); it must have unit stride in the last dim."""
        ops._require_cuda_f32(h_states, "h_states")
        if h_states.dim() != 3:
            raise ValueError("h_states must be (B, N, h_dim)")
        if torch.is_grad_enabled() and (h_states.requires_grad or any(p.requires_grad for p in self.parameters())):
            if self.precision != "fp32" and not self.__dict__.get("_warned_train_precision"):
                import warnings
                warnings.warn(f'groupnet_b200: precision="{self.precision}" is an inference path; with autograd enabled '
                              "the differentiable fp32 kernels run instead (wrap inference in torch.no_grad())",
                              RuntimeWarning, stacklevel=3)
                self.__dict__["_warned_train_precision"] = True
            return self._run_train(h_states, inc, e, noise, node_out, want_dist)
        h = h_states.detach().contiguous()
        b, n, d = h.shape
        if d != self.h_dim:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied (h_dim {d} != {self.h_dim})")
        dev = h.device
        stages = self._packs.get(self, dev)
        n_stage = len(stages)
        t = self.edge_types
        us = self._noise_list(noise, b, e, n_stage, dev)
        seed_dev = None
        if us is None:
            seed = (self.philox_seed + _PHILOX_CALL_STRIDE * self._philox_calls) & _MASK64
            if self.rng == "philox-device":
                seed_dev = self._device_seed(dev)
            self._philox_calls += 1
        else:
            seed = 0
        dout = stages[-1].dout
        if node_out is None:
            node_out = torch.empty(b, n, dout, dtype=torch.float32, device=dev)
        else:
            ops._require_cuda_f32(node_out, "node_out")
            if tuple(node_out.shape) != (b, n, dout) or node_out.stride(2) != 1 \
                    or node_out.stride(0) != n * node_out.stride(1):
                raise ValueError("node_out must be (B,N,bottleneck_dim) with uniform row stride")
        out_ld = node_out.stride(1) if b * n > 1 else dout
        if dist_out is None and want_dist:
            dist_out = torch.empty(b, e, t, dtype=torch.float32, device=dev)
        mids = [torch.empty(b, n, stages[s].dout, dtype=torch.float32, device=dev) for s in range(n_stage - 1)]
        if b == 0:
            if seed_dev is not None:
                seed_dev.add_(_as_int64(_PHILOX_CALL_STRIDE))
            return node_out, dist_out

        cfg = _lib.StageCfg()
        cfg.N, cfg.D, cfg.E, cfg.T = n, d, e, t
        cfg.pairwise = 1 if self._pairwise else 0
        cfg.precision = _PRECISION[self.precision]
        cfg.noise_mode = (_lib.GN_NOISE_GIVEN if us is not None else
                          _lib.GN_NOISE_PHILOX_DEVICE_SEED if seed_dev is not None else _lib.GN_NOISE_PHILOX)
        cfg.seed = seed
        # bound the scratch: split the batch so one call's workspace stays under the limit
        cfg.B, cfg.Dout, cfg.stage_index, cfg.scene_offset = 1, max(s.dout for s in stages), 0, 0
        per_scene = max(ops.stage_workspace_bytes(cfg), 1)
        chunk = max(1, min(b, self.workspace_limit_bytes // per_scene))
        for b0 in range(0, b, chunk):
            b1 = min(b, b0 + chunk)
            cur = h[b0:b1]
            for s in range(n_stage):
                last = s == n_stage - 1
                cfg.B, cfg.Dout, cfg.stage_index = b1 - b0, stages[s].dout, s
                cfg.out_ld = out_ld if last else 0
                cfg.h_stride = 0 if inc is None else inc.stride(0)
                cfg.scene_offset = self.scene_offset + b0
                ws = self._ws.get(ops.stage_workspace_bytes(cfg), dev)
                dst = node_out[b0:b1] if last else mids[s][b0:b1]
                ops.stage_forward(cfg, stages[s], cur,
                                  None if inc is None else inc[b0:b1],
                                  seed_dev if seed_dev is not None else None if us is None else us[s][b0:b1],
                                  dst, dist_out[b0:b1] if (s == 0 and dist_out is not None) else None, ws)
                cur = dst
        if seed_dev is not None:
            seed_dev.add_(_as_int64(_PHILOX_CALL_STRIDE))      # stream-ordered (and captured): next call's seed
        return node_out, dist_out

    def _run_train(self, h_states, inc, e, noise, node_out, want_dist):
        """Differentiable forward: every stage goes through autograd.StageFunction (fp32 kernels
        forward, gn_stage_bwd backward).  The tensor-core path is inference-only."""
        from .autograd import StageFunction, stage_param_list
        b, n, d = h_states.shape
        if d != self.h_dim:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied (h_dim {d} != {self.h_dim})")
        dev = h_states.device
        stages = self._packs.get(self, dev)
        n_stage = len(stages)
        t = self.edge_types
        if b == 0:                                   # empty batch: empty outputs, like the inference path
            out = h_states.new_zeros(0, n, stages[-1].dout) + 0.0 * h_states.sum()
            return out, (h_states.new_zeros(0, e, t) if want_dist else None)
        us = self._noise_list(noise, b, e, n_stage, dev)
        if us is None:
            if self.rng == "philox-device":
                raise RuntimeError('rng="philox-device" is an inference mode (graph replay); train with "philox"')
            seed = (self.philox_seed + _PHILOX_CALL_STRIDE * self._philox_calls) & _MASK64
            self._philox_calls += 1
        else:
            seed = 0
        cur = h_states
        dist0 = None
        for s in range(n_stage):
            meta = (b, n, d, stages[s].dout, e, t, 1 if self._pairwise else 0,
                    _lib.GN_NOISE_GIVEN if us is not None else _lib.GN_NOISE_PHILOX, s, seed, self.scene_offset)
            cur, dist = StageFunction.apply(cur, inc, None if us is None else us[s], meta, stages[s],
                                            *stage_param_list(self, s))
            if s == 0:
                dist0 = dist
        if node_out is not None:
            node_out.copy_(cur)
        return cur, (dist0 if want_dist else None)

    def launches_per_forward(self, batch: int, n: int, e: int) -> int:
        """Kernel launches one forward issues (for bench.py's gpu_launches)."""
        cfg = _lib.StageCfg()
        cfg.B, cfg.N, cfg.D, cfg.E, cfg.T = batch, n, self.h_dim, e, self.edge_types
        cfg.Dout, cfg.pairwise = self.bottleneck_dim, 1 if self._pairwise else 0
        cfg.precision = _PRECISION[self.precision]
        return ops.stage_launch_count(cfg) * max(self.nmp_layers, 1)


# --------------------------------------------------------------------------
# the two public layers
# --------------------------------------------------------------------------
class MS_HGNN_oridinary(_MessagePassingLayer):
    """Pairwise neural message passing over all N^2 ordered agent pairs
    (`MS_HGNN_oridinary`, model/MS_HGNN_batch.py:55-198).

    forward(h_states (B,N,h_dim)) -> (node_feat (B,N,bottleneck_dim), factors (B,N*N,6))"""
    _pairwise = True

    def __init__(self, embedding_dim=64, h_dim=64, mlp_dim=1024, bottleneck_dim=1024,
                 activation='relu', batch_norm=True, dropout=0.0, nmp_layers=4, vis=False):
        super().__init__()
        self.mlp_dim = mlp_dim
        self.h_dim = h_dim
        self.bottleneck_dim = bottleneck_dim
        self.embedding_dim = embedding_dim
        self.nmp_layers = nmp_layers
        self.batch_norm = batch_norm
        self.activation = activation
        self.vis = vis
        self.hdim_extend = 64
        self.edge_types = 6
        self._build_tree(h_dim, bottleneck_dim, nmp_layers)

    def forward(self, h_states, *, noise=None, out=None, want_factors=True):
        n = h_states.shape[1]
        return self._run(h_states, None, n * n, noise, node_out=out, want_dist=want_factors)


class MS_HGNN_hyper(_MessagePassingLayer):
    """Group-wise message passing over top-`scale` correlation hyperedges
    (`MS_HGNN_hyper`, model/MS_HGNN_batch.py:270-443).

    forward(h_states (B,N,h_dim), corr (B,N,N)) ->
        (node_feat (B,N,bottleneck_dim), factor (B,E,10), H (B,E,N)),  E = 1 if scale == N else N.
    `H=` lets a caller pass the incidence produced by `ops.corr_topk_h` for all
    scales in one fused pass; `corr` is then ignored."""
    _pairwise = False

    def __init__(self, embedding_dim=64, h_dim=64, mlp_dim=1024, bottleneck_dim=1024,
                 activation='relu', batch_norm=True, dropout=0.0, nmp_layers=4, scale=2, vis=False,
                 actor_number=11):
        super().__init__()
        self.mlp_dim = mlp_dim
        self.h_dim = h_dim
        self.bottleneck_dim = bottleneck_dim
        self.embedding_dim = embedding_dim
        self.nmp_layers = nmp_layers
        self.batch_norm = batch_norm
        self.activation = activation
        self.scale = scale
        self.vis = vis
        # created and never used by the reference forward (:290-291); part of the state_dict
        self.spatial_embedding = nn.Linear(2, embedding_dim)
        self.spatial_transform = nn.Linear(h_dim, h_dim)
        self.hdim_extend = 64
        self.edge_types = 10
        self._build_tree(h_dim, bottleneck_dim, nmp_layers)
        self.listall = False

    def init_adj_attention(self, feat, feat_corr, scale_factor=2):
        """Top-k incidence (:372-388) on the GPU."""
        ops._require_cuda_f32(feat, "feat")
        return ops.topk_h(feat_corr, scale_factor)

    def forward(self, h_states, corr=None, *, noise=None, H=None, out=None, want_factors=True):
        if H is None:
            if corr is None:
                raise TypeError("forward() missing 1 required positional argument: 'corr'")
            H = self.init_adj_attention(h_states, corr, scale_factor=self.scale)
        else:
            ops._require_cuda_f32(H, "H")
        # rows of a concatenated (B, sum E, N) incidence are accepted as they are
        ok = H.dim() == 3 and H.stride(2) == 1 and H.stride(1) == H.shape[2]
        inc = H if ok else H.contiguous()
        node_feat, factor = self._run(h_states, inc, inc.shape[1], noise, node_out=out, want_dist=want_factors)
        return node_feat, factor, H
