"""Shared by tests/golden/make_golden_fish.py (reference side) and the fish-operator tests (drop-in side): seeded inputs
at the fish model's shapes and a deterministic randomisation of every BatchNorm1d (so eval-mode folding is exercised
with non-trivial statistics); weights themselves come from `torch.manual_seed(seed)` + the constructor."""
import hashlib

import numpy as np
import torch
import torch.nn as nn


def state_sha(module) -> str:
    hsh = hashlib.sha256()
    for k, v in module.state_dict().items():
        if k.endswith("num_batches_tracked"):
            continue
        hsh.update(k.encode())
        hsh.update(v.detach().cpu().numpy().astype(np.float32).tobytes())
    return hsh.hexdigest()


def randomize_bn(module, seed: int) -> None:
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, nn.BatchNorm1d):
                c = m.num_features
                m.weight.copy_(1.0 + 0.3 * torch.randn(c, generator=gen))
                m.bias.copy_(0.2 * torch.randn(c, generator=gen))
                m.running_mean.copy_(0.3 * torch.randn(c, generator=gen))
                m.running_var.copy_(0.5 + torch.rand(c, generator=gen))


def fish_inputs(b: int, n: int, m: int, f_v: int, n_hid: int, seed: int):
    """rel_rec / rel_send: the fully connected off-diagonal graph one-hot encoded and repeated per scene, as the fish
    scripts build it; I_HG: a hard (one-hot over M) membership per node, as gumbel_softmax(hard=True) yields."""
    gen = torch.Generator().manual_seed(seed + 5000)
    off = [(i, j) for i in range(n) for j in range(n) if i != j]
    e = len(off)
    rec = torch.zeros(e, n)
    snd = torch.zeros(e, n)
    for k, (i, j) in enumerate(off):
        rec[k, i] = 1.0
        snd[k, j] = 1.0
    grp = torch.randint(0, m, (b, n), generator=gen)
    i_hg = torch.zeros(b, n, m).scatter_(2, grp.unsqueeze(-1), 1.0)
    return {
        "rel_rec": rec.unsqueeze(0).repeat(b, 1, 1), "rel_send": snd.unsqueeze(0).repeat(b, 1, 1), "I_HG": i_hg,
        "v_self": torch.randn(b, n, n_hid, generator=gen), "v_combined": torch.randn(b, n, f_v, generator=gen),
        "z_CG": torch.softmax(torch.randn(b, e, 3, generator=gen), -1), "z_HG": torch.softmax(torch.randn(b, m, 3, generator=gen), -1),
    }
