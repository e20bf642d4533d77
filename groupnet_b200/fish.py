"""Drop-ins for the group-wise operators of the fish model:

    compute_alpha_im(alpha_ij, I_HG, rel_rec, rel_send)                     model/encoder.py:261-303
    MLPHGE(n_in, n_hid, n_out, do_prob)                                     model/encoder.py:200-256
    HyperEdgeAttention(input_dim_e, input_dim_v, hidden_dim, node_dim)      model/encoder.py:102-197
    TemporalGATLayer(out_dim, input_dim, hidden_dim, num_heads, concat)     model/encoder.py:331-467
    build_dynamic_graph_and_hypergraph(z_CG, z_HG, rel_rec, rel_send, I)    utilities/utils.py:191-244

Same constructors, attribute names, registration and initialisation order (hence the same seeded default weights and
state_dict) and the same forward signatures / outputs as the reference.  The forwards are the `gn_fish_*` entry points
of `libgroupnet_b200.so` (csrc/gn_fish.cu): no PyTorch math on the path.  Eval mode only: the BatchNorm1d layers are
folded into their Linears from the running statistics, dropout is the identity; in training mode (batch statistics,
dropout) the modules raise.  CUDA fp32 tensors only, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib, ops

ACT_NONE, ACT_LEAKY, ACT_ELU = 0, 1, 2


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _rel(rel: torch.Tensor, batch: int, e: int, n: int, name: str) -> Tuple[torch.Tensor, int]:
    """(tensor, scene stride in floats) for a (E, N), (1, E, N) or (B, E, N) relation matrix."""
    ops._require_cuda_f32(rel, name)
    if rel.dim() == 2:
        rel = rel.unsqueeze(0)
    if tuple(rel.shape[1:]) != (e, n) or rel.shape[0] not in (1, batch):
        raise RuntimeError(f"{name}: expected (B, {e}, {n}), got {tuple(rel.shape)}")
    rel = rel.contiguous()
    return rel, (0 if rel.shape[0] == 1 else e * n)


def _run_mlp(x2d: torch.Tensor, layers: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor], int, float]]) -> torch.Tensor:
    """gn_fish_mlp: x2d (R, K0) -> (R, N_last); layers = [(Wt (K, N), bias (N) | None, act, slope)]."""
    lib = _lib.load()
    r, k0 = x2d.shape
    n = len(layers)
    wt = (C.c_void_p * n)(*[w.data_ptr() for w, _, _, _ in layers])
    bs = (C.c_void_p * n)(*[0 if b is None else b.data_ptr() for _, b, _, _ in layers])
    ns = (C.c_int32 * n)(*[int(w.shape[1]) for w, _, _, _ in layers])
    acts = (C.c_int32 * n)(*[a for _, _, a, _ in layers])
    slopes = (C.c_float * n)(*[s for _, _, _, s in layers])
    out = torch.empty(r, int(layers[-1][0].shape[1]), dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = lib.gn_fish_mlp(_ptr(x2d), x2d.stride(0), r, k0, n, wt, bs, ns, acts, slopes, _ptr(out), out.stride(0),
                             ops._stream_ptr(x2d.device))
    _lib.check(rc, "gn_fish_mlp")
    return out


def _bmm_t(a: torch.Tensor, a_stride: int, b: torch.Tensor, batch: int, r: int, cn: int, f: int,
           roww: Optional[torch.Tensor] = None, norm_cols: bool = False) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(batch, cn, f, dtype=torch.float32, device=b.device)
    with torch.cuda.device(b.device):
        rc = lib.gn_fish_bmm_t(_ptr(a), a_stride, _ptr(b), _ptr(roww), batch, r, cn, f, 1 if norm_cols else 0, _ptr(out),
                               ops._stream_ptr(b.device))
    _lib.check(rc, "gn_fish_bmm_t")
    return out


def _fold(lin: nn.Linear, bn: Optional[nn.BatchNorm1d]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(Wt (K, N), bias (N)) of Linear followed by an eval-mode BatchNorm1d, folded in float64."""
    w = lin.weight.detach().double()
    b = lin.bias.detach().double() if lin.bias is not None else None
    if bn is not None:
        s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        w = w * s[:, None]
        b = ((b if b is not None else 0.0) - bn.running_mean.detach().double()) * s + bn.bias.detach().double()
    return w.t().contiguous().float(), (None if b is None else b.float().contiguous())


class _FoldedModule(nn.Module):
    """Caches the BN-folded, transposed weights per version of the module's parameters and buffers."""

    def _fingerprint(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _folded(self):
        key = self._fingerprint()
        cache = self.__dict__.get("_fold_cache")
        if cache is None or cache[0] != key:
            cache = (key, self._fold_all())
            self.__dict__["_fold_cache"] = cache
        return cache[1]

    def invalidate_packs(self):
        self.__dict__.pop("_fold_cache", None)

    def _check_eval(self):
        if self.training:
            raise NotImplementedError(f"groupnet_b200.{type(self).__name__} is eval-only (BatchNorm running statistics "
                                      "are folded into the Linears): call .eval()")

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_fold_cache", None)
        return state


# ---------------------------------------------------------------------------------------------
def compute_alpha_im(alpha_ij: torch.Tensor, I_HG: torch.Tensor, rel_rec: torch.Tensor, rel_send: torch.Tensor) -> torch.Tensor:
    """model/encoder.py:261-303.  alpha_ij (B, E) or (B, E, 1); I_HG (B, N, M) -> alpha_im (B, N, M)."""
    ops._require_cuda_f32(alpha_ij, "alpha_ij")
    ops._require_cuda_f32(I_HG, "I_HG")
    b, n, m = I_HG.shape
    if alpha_ij.dim() == 3:
        if alpha_ij.shape[-1] != 1:
            raise RuntimeError("compute_alpha_im: alpha_ij (B, E, H) broadcasts against (B, E, M) only for H == 1")
        alpha_ij = alpha_ij[..., 0]
    e = alpha_ij.shape[1]
    rec, stride = _rel(rel_rec, b, e, n, "rel_rec")
    snd, stride_s = _rel(rel_send, b, e, n, "rel_send")
    if stride != stride_s:
        snd = snd.expand(b, e, n).contiguous() if stride else snd
        rec = rec.expand(b, e, n).contiguous() if stride_s else rec
        stride = e * n
    out = torch.empty(b, n, m, dtype=torch.float32, device=I_HG.device)
    lib = _lib.load()
    with torch.cuda.device(I_HG.device):
        rc = lib.gn_fish_alpha_im(_ptr(alpha_ij.contiguous()), _ptr(I_HG.contiguous()), _ptr(rec), _ptr(snd), stride,
                                  b, e, n, m, _ptr(out), ops._stream_ptr(I_HG.device))
    _lib.check(rc, "gn_fish_alpha_im")
    return out


class MLPHGE(_FoldedModule):
    def __init__(self, n_in, n_hid, n_out, do_prob):
        super().__init__()
        self.fc1 = nn.Linear(n_in, n_hid)
        self.fc2 = nn.Linear(n_hid, n_hid)
        self.fc3 = nn.Linear(n_hid, n_out)
        self.bn = nn.BatchNorm1d(n_hid)
        self.bn2 = nn.BatchNorm1d(n_out)
        self.dropout_prob = do_prob
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight.data)
                m.bias.data.fill_(0.1)
            elif isinstance(m, nn.BatchNorm1d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _fold_all(self):
        return [_fold(self.fc1, self.bn) + (ACT_ELU, 0.0), _fold(self.fc2, self.bn) + (ACT_ELU, 0.0),
                _fold(self.fc3, self.bn2) + (ACT_ELU, 0.0)]

    def forward(self, alpha_im, V_CG):
        """alpha_im (B, N, M), V_CG (B, N, F) -> e_HG (B, M, n_out)  (:224-251)."""
        self._check_eval()
        ops._require_cuda_f32(alpha_im, "alpha_im")
        ops._require_cuda_f32(V_CG, "V_CG")
        b, n, m = alpha_im.shape
        f = V_CG.shape[-1]
        wn = _bmm_t(alpha_im.contiguous(), n * m, V_CG.contiguous(), b, n, m, f, norm_cols=True)      # (B, M, F)
        out = _run_mlp(wn.view(b * m, f), self._folded())
        return out.view(b, m, -1)


class HyperEdgeAttention(_FoldedModule):
    def __init__(self, input_dim_e, input_dim_v, hidden_dim, node_dim, alpha=0.2):
        super().__init__()
        self.W1 = nn.Linear(input_dim_e, hidden_dim, bias=False)
        self.W2 = nn.Linear(input_dim_v, hidden_dim, bias=False)
        self.attention_vector = nn.Parameter(torch.Tensor(hidden_dim * 2))
        self.leaky_relu = nn.LeakyReLU(alpha)
        nn.init.xavier_uniform_(self.W1.weight, gain=1.414)
        nn.init.xavier_uniform_(self.W2.weight, gain=1.414)
        nn.init.xavier_uniform_(self.attention_vector.unsqueeze(0), gain=1.414)
        self.f_HG_v = nn.Sequential(nn.Linear(input_dim_e, hidden_dim), nn.BatchNorm1d(hidden_dim),
                                    nn.LeakyReLU(negative_slope=0.01), nn.Linear(hidden_dim, node_dim),
                                    nn.BatchNorm1d(node_dim))
        self.f_HG_2 = nn.Sequential(nn.Linear(node_dim, hidden_dim), nn.BatchNorm1d(hidden_dim),
                                    nn.LeakyReLU(negative_slope=0.01), nn.Linear(hidden_dim, node_dim),
                                    nn.BatchNorm1d(node_dim))

    def _fold_all(self):
        a = float(self.leaky_relu.negative_slope)
        return {
            "w1": [_fold(self.W1, None) + (ACT_LEAKY, a)],
            "w2": [_fold(self.W2, None) + (ACT_LEAKY, a)],
            "v": [_fold(self.f_HG_v[0], self.f_HG_v[1]) + (ACT_LEAKY, 0.01), _fold(self.f_HG_v[3], self.f_HG_v[4]) + (ACT_LEAKY, a)],
            "e2": [_fold(self.f_HG_2[0], self.f_HG_2[1]) + (ACT_LEAKY, 0.01), _fold(self.f_HG_2[3], self.f_HG_2[4]) + (ACT_LEAKY, a)],
            "avec": self.attention_vector.detach().float().contiguous(),
        }

    def forward(self, e_HG, v_CG, I_HG):
        """e_HG (B, M, Fe), v_CG (B, N, Fv), I_HG (B, N, M) -> e_HG_2 (B, M, node_dim)  (:141-197)."""
        self._check_eval()
        for name, t in (("e_HG", e_HG), ("v_CG", v_CG), ("I_HG", I_HG)):
            ops._require_cuda_f32(t, name)
        b, n, m = I_HG.shape
        fe = e_HG.shape[-1]
        p = self._folded()
        e_hg = e_HG.contiguous()
        e_proj = _run_mlp(e_hg.view(b * m, fe), p["w1"])                       # leaky(W1 e_HG)          :160
        v_proj = _run_mlp(v_CG.contiguous().view(b * n, -1), p["w2"])          # leaky(W2 v_CG)          :161
        hd = e_proj.shape[1]
        v1 = torch.empty(b, n, fe, dtype=torch.float32, device=e_HG.device)
        lib = _lib.load()
        with torch.cuda.device(e_HG.device):
            rc = lib.gn_fish_hga_core(_ptr(e_proj), _ptr(v_proj), _ptr(p["avec"]), _ptr(I_HG.contiguous()), _ptr(e_hg),
                                      b, n, m, hd, fe, float(self.leaky_relu.negative_slope), _ptr(v1),
                                      ops._stream_ptr(e_HG.device))
        _lib.check(rc, "gn_fish_hga_core")
        v1 = _run_mlp(v1.view(b * n, fe), p["v"])                              # leaky(f_HG_v(.))        :181-182
        nd = v1.shape[1]
        e2 = _bmm_t(I_HG.contiguous(), n * m, v1, b, n, m, nd)                 # einsum('bnm,bnf->bmf')  :185
        e2 = _run_mlp(e2.view(b * m, nd), p["e2"])                             # leaky(f_HG_2(.))        :188-189
        return e2.view(b, m, nd)


class TemporalGATLayer(_FoldedModule):
    def __init__(self, out_dim, input_dim, hidden_dim, num_heads=1, concat_heads=True, alpha=0.2):
        super().__init__()
        if num_heads != 1:
            raise NotImplementedError("groupnet_b200.TemporalGATLayer: num_heads == 1 (the fish model's setting, "
                                      "test_fish.py:331) is the only head count built")
        self.num_heads = num_heads
        self.concat_heads = concat_heads
        self.out_dim = out_dim
        self.projection = nn.Linear(hidden_dim, self.out_dim * num_heads, bias=False)
        self.a_forward = nn.Parameter(torch.Tensor(num_heads, self.out_dim))
        self.a_backward = nn.Parameter(torch.Tensor(num_heads, self.out_dim))
        self.leaky_relu = nn.LeakyReLU(alpha)
        self.f_CG_e = nn.Sequential(nn.Linear(2 * self.out_dim, self.out_dim), nn.BatchNorm1d(out_dim),
                                    nn.LeakyReLU(negative_slope=0.01), nn.Linear(self.out_dim, self.out_dim),
                                    nn.BatchNorm1d(out_dim))
        self.f_CG_v = nn.Sequential(nn.Linear(self.out_dim, hidden_dim), nn.BatchNorm1d(hidden_dim),
                                    nn.LeakyReLU(negative_slope=0.01), nn.Linear(hidden_dim, hidden_dim),
                                    nn.BatchNorm1d(hidden_dim))
        nn.init.xavier_uniform_(self.projection.weight, gain=1.414)
        nn.init.xavier_uniform_(self.a_forward, gain=1.414)
        nn.init.xavier_uniform_(self.a_backward, gain=1.414)

    def _fold_all(self):
        a = float(self.leaky_relu.negative_slope)
        return {
            "proj": [_fold(self.projection, None) + (ACT_LEAKY, a)],
            "e": [_fold(self.f_CG_e[0], self.f_CG_e[1]) + (ACT_LEAKY, 0.01), _fold(self.f_CG_e[3], self.f_CG_e[4]) + (ACT_LEAKY, a)],
            "v": [_fold(self.f_CG_v[0], self.f_CG_v[1]) + (ACT_LEAKY, 0.01), _fold(self.f_CG_v[3], self.f_CG_v[4]) + (ACT_LEAKY, a)],
            "af": self.a_forward.detach().float().contiguous(), "ab": self.a_backward.detach().float().contiguous(),
        }

    def forward(self, v_self, rel_rec, rel_send):
        """v_self (B, N, F), rel_rec / rel_send (B, E, N) -> (v_social (B, N, H*hidden | hidden), alpha_ij (B, E, H))  (:385-467)."""
        self._check_eval()
        ops._require_cuda_f32(v_self, "v_self")
        b, n, f = v_self.shape
        h, d = self.num_heads, self.out_dim
        e = rel_rec.shape[-2]
        rec, stride = _rel(rel_rec, b, e, n, "rel_rec")
        snd, stride_s = _rel(rel_send, b, e, n, "rel_send")
        if stride != stride_s:
            snd = snd.expand(b, e, n).contiguous() if stride else snd
            rec = rec.expand(b, e, n).contiguous() if stride_s else rec
            stride = e * n
        p = self._folded()
        dev = v_self.device
        v_proj = _run_mlp(v_self.contiguous().view(b * n, f), p["proj"])                 # (B*N, H*D)    :404-405
        edge_input = torch.empty(b * e * h, 2 * d, dtype=torch.float32, device=dev)
        alpha_ij = torch.empty(b, e, h, dtype=torch.float32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            rc = lib.gn_fish_gat_edges(_ptr(v_proj), _ptr(rec), _ptr(snd), stride, _ptr(p["af"]), _ptr(p["ab"]), b, e, n, h, d,
                                       float(self.leaky_relu.negative_slope), _ptr(edge_input), _ptr(alpha_ij),
                                       ops._stream_ptr(dev))
        _lib.check(rc, "gn_fish_gat_edges")
        e_cg = _run_mlp(edge_input, p["e"])                                               # (B*E*H, D)    :447
        # v_social = einsum("behd,ben->bnhd", e_CG * alpha_ij, rel_rec): rows e, weights alpha_ij (H == 1) folded in   :451-454
        v_soc = _bmm_t(rec, stride, e_cg, b, e, n, d, roww=alpha_ij.view(b, e))
        v_soc = _run_mlp(v_soc.reshape(b * n * h, d), p["v"])                             # leaky(f_CG_v(.))  :456
        hid = v_soc.shape[1]
        v_soc = v_soc.view(b, n, h, hid)
        v_social = v_soc.reshape(b, n, h * hid) if self.concat_heads else v_soc.mean(dim=2)
        return v_social, alpha_ij


def build_dynamic_graph_and_hypergraph(z_CG, z_HG, rel_rec, rel_send, I_HG):
    """utilities/utils.py:191-244 -> (new_rel_rec, new_rel_send, new_I_HG, edge_types, hyperedge_types)."""
    for name, t in (("z_CG", z_CG), ("z_HG", z_HG), ("I_HG", I_HG)):
        ops._require_cuda_f32(t, name)
    b, e, lc = z_CG.shape
    _, m, lh = z_HG.shape
    n = rel_rec.shape[-1]
    rec, stride = _rel(rel_rec, b, e, n, "rel_rec")
    snd, stride_s = _rel(rel_send, b, e, n, "rel_send")
    if stride != stride_s:
        snd = snd.expand(b, e, n).contiguous() if stride else snd
        rec = rec.expand(b, e, n).contiguous() if stride_s else rec
        stride = e * n
    dev = z_CG.device
    new_rec = torch.empty(b, e, n, dtype=torch.float32, device=dev)
    new_snd = torch.empty(b, e, n, dtype=torch.float32, device=dev)
    new_i = torch.empty(b, n, m, dtype=torch.float32, device=dev)
    et = torch.empty(b, e, dtype=torch.int64, device=dev)
    ht = torch.empty(b, m, dtype=torch.int64, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.gn_fish_dynamic_graph(_ptr(z_CG.contiguous()), _ptr(z_HG.contiguous()), _ptr(rec), _ptr(snd), stride,
                                       _ptr(I_HG.contiguous()), b, e, n, m, lc, lh, _ptr(new_rec), _ptr(new_snd), _ptr(new_i),
                                       _ptr(et), _ptr(ht), ops._stream_ptr(dev))
    _lib.check(rc, "gn_fish_dynamic_graph")
    return new_rec, new_snd, new_i, et, ht
