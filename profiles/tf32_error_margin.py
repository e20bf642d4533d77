"""Error margin of the fp32-grade tensor-core path against the goldens: max|d| / max|ref| per output tensor (the parity
bar is 1e-5), printed per golden case.  Run on a B200 from the repo root."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import build_layer, golden_names, load_golden

worst = 0.0
with torch.no_grad():
    for name in golden_names():
        g = load_golden(name)
        m = build_layer(g).to("cuda").set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
        h = torch.from_numpy(g["h"]).cuda()
        noise = [torch.from_numpy(u).cuda() for u in g["noise"]]
        out = m(h, noise=noise) if g["kind"] == "pairwise" else m(h, torch.from_numpy(g["corr"]).cuda(), noise=noise)
        errs = []
        for got, ref in ((out[0], g["node_feat"]), (out[1], g["factors"])):
            ref = torch.from_numpy(ref)
            errs.append(float((got.cpu() - ref).abs().max() / ref.abs().max()))
        worst = max(worst, *errs)
        print(f"{name:32s} node {errs[0]:.2e}  factors {errs[1]:.2e}")
print(f"worst {worst:.2e} (bar 1e-5)")
