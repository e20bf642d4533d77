"""Generate tests/golden/fish/*.npz by running the UNMODIFIED reference fish-model operators on CPU
(/root/reference/model/encoder.py:102-467, utilities/utils.py:191-244), eval mode.

Run in the build container only:  python tests/golden/make_golden_fish.py

Weights: `torch.manual_seed(seed)` + the reference constructor (the drop-ins reproduce that seeded initialisation), then
tests/fish_schema.py::randomize_bn gives the BatchNorm layers non-trivial affine parameters and running statistics from
a second seed; both are pinned by sha256.  Fixtures store inputs and outputs only.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GROUPNET_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))
if "networkx" not in sys.modules:
    try:
        import networkx  # noqa: F401
    except Exception:
        sys.modules["networkx"] = types.ModuleType("networkx")

from model.encoder import HyperEdgeAttention, MLPHGE, TemporalGATLayer, compute_alpha_im  # noqa: E402
from utilities.utils import build_dynamic_graph_and_hypergraph  # noqa: E402
from fish_schema import fish_inputs, randomize_bn, state_sha  # noqa: E402


def main():
    torch.set_num_threads(4)
    os.makedirs(os.path.join(HERE, "fish"), exist_ok=True)
    # the fish model's dimensions (test_fish.py:326-339: n_in 10, n_hid 128, n_out 5, one head, M 5) at N = 11 and N = 8
    for name, b, n, m, seed in (("nba11", 6, 11, 5, 91), ("fish8", 5, 8, 4, 92), ("n3_m1", 3, 3, 1, 93)):
        n_hid, n_fc = 128, 5
        f_v = n_hid + n_fc
        inp = fish_inputs(b, n, m, f_v, n_hid, seed)
        out = {}
        with torch.no_grad():
            torch.manual_seed(seed)
            gat = TemporalGATLayer(out_dim=n_hid, input_dim=10, hidden_dim=n_hid, num_heads=1, concat_heads=True)
            randomize_bn(gat, seed + 1000)
            gat.eval()
            v_social, alpha_ij = gat(inp["v_self"], inp["rel_rec"], inp["rel_send"])
            out.update(gat_sha=state_sha(gat), v_social=v_social.numpy(), alpha_ij=alpha_ij.numpy())
            alpha_im = compute_alpha_im(alpha_ij, inp["I_HG"], inp["rel_rec"], inp["rel_send"])
            out["alpha_im"] = alpha_im.numpy()
            torch.manual_seed(seed + 1)
            hge = MLPHGE(f_v, n_hid, n_fc * 3, 0.0)
            randomize_bn(hge, seed + 1001)
            hge.eval()
            e_hg = hge(alpha_im, inp["v_combined"])
            out.update(hge_sha=state_sha(hge), e_HG=e_hg.numpy())
            torch.manual_seed(seed + 2)
            hga = HyperEdgeAttention(n_fc * 3, f_v, n_hid, n_fc * 5)
            randomize_bn(hga, seed + 1002)
            hga.eval()
            e_hg2 = hga(e_hg, inp["v_combined"], inp["I_HG"])
            out.update(hga_sha=state_sha(hga), e_HG_2=e_hg2.numpy())
            nr, ns, ni, et, ht = build_dynamic_graph_and_hypergraph(inp["z_CG"], inp["z_HG"], inp["rel_rec"], inp["rel_send"],
                                                                    inp["I_HG"])
            out.update(new_rel_rec=nr.numpy(), new_rel_send=ns.numpy(), new_I_HG=ni.numpy(), edge_types=et.numpy(),
                       hyperedge_types=ht.numpy())
        np.savez_compressed(os.path.join(HERE, "fish", name + ".npz"), B=b, N=n, M=m, seed=seed, **out)
        print(f"{name}: v_social {tuple(v_social.shape)} alpha_im {tuple(alpha_im.shape)} e_HG {tuple(e_hg.shape)} "
              f"e_HG_2 {tuple(e_hg2.shape)} max|v_social| {v_social.abs().max():.4f}")


if __name__ == "__main__":
    main()
