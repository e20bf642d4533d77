"""Stress script (not collected by pytest): re-run one fp32 forward many times with the
workspace and the output buffers NaN-poisoned before every call, optionally with a
different layer's kernels in between, and compare every result bit-for-bit with the first
one and with the oracle.  Back-to-back identical forwards cannot see a stale-read race
(the stale value equals the fresh one); this can.

    python tests/stress_stale_memory.py [iters]
"""
import sys
import pathlib

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import groupnet_b200 as gb                      # noqa: E402
from oracle import ms_hgnn_oracle as O          # noqa: E402

DEV = torch.device("cuda:0")


def make(kind, n, d, bo, scale, layers, b):
    torch.manual_seed(1000 + n + d + layers)
    if kind == "pairwise":
        m = gb.MS_HGNN_oridinary(16, d, 64, bo, batch_norm=0, nmp_layers=layers)
        e, t = n * n, 6
    else:
        m = gb.MS_HGNN_hyper(d, d, 64, bo, batch_norm=0, nmp_layers=layers, scale=scale)
        e, t = (1 if scale == n else n), 10
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(b)
    h = torch.randn(b, n, d, generator=gen)
    noise = [torch.rand(b, e, t, generator=gen) for _ in range(layers)]
    m = m.to(DEV)
    m.workspace_limit_bytes = 64 << 20
    with torch.no_grad():
        if kind == "pairwise":
            ref_node, ref_fac = O.forward_pairwise(sd, h, noise, nmp_layers=layers)
            inc = None
        else:
            corr = O.feature_correlation(h)
            ref_node, ref_fac, ref_h = O.forward_hyper(sd, h, corr, scale, noise, nmp_layers=layers)
            inc = ref_h.to(DEV)
    return dict(m=m, h=h.to(DEV), noise=[u.to(DEV) for u in noise], inc=inc, e=e, t=t, bo=bo,
                ref_node=ref_node.to(DEV), ref_fac=ref_fac.to(DEV), name=f"{kind}-{n}-{d}-{bo}-{scale}-{layers}-{b}")


def run(c, node, fac):
    with torch.no_grad():
        c["m"]._run(c["h"], c["inc"], c["e"], c["noise"], node_out=node, dist_out=fac)


def stress(c, other, iters, interleave):
    b, n, _ = c["h"].shape
    node = torch.empty(b, n, c["bo"], device=DEV)
    fac = torch.empty(b, c["e"], c["t"], device=DEV)
    onode = torch.empty(other["h"].shape[0], other["h"].shape[1], other["bo"], device=DEV)
    ofac = torch.empty(other["h"].shape[0], other["e"], other["t"], device=DEV)
    run(c, node, fac)
    torch.cuda.synchronize()
    node0, fac0 = node.clone(), fac.clone()
    e_f0 = ((fac0 - c["ref_fac"]).abs().max() / c["ref_fac"].abs().max()).item()
    e_n0 = ((node0 - c["ref_node"]).abs().max() / c["ref_node"].abs().max()).item()
    bad = 0
    for it in range(iters):
        c["m"]._ws.buf.fill_(0xFF)
        node.fill_(float("nan"))
        fac.fill_(float("nan"))
        if interleave:
            run(other, onode, ofac)
            c["m"]._ws.buf.fill_(0xFF)
        run(c, node, fac)
        same_f, same_n = torch.equal(fac, fac0), torch.equal(node, node0)
        if not (same_f and same_n):
            bad += 1
            if bad <= 8:
                df = torch.nan_to_num((fac - fac0).abs(), nan=1e30)
                dn = torch.nan_to_num((node - node0).abs(), nan=1e30)
                fi = [int(x) for x in torch.unravel_index(df.argmax(), df.shape)]
                ni = [int(x) for x in torch.unravel_index(dn.argmax(), dn.shape)]
                print(f"  MISMATCH it={it} fac: n_diff={(df > 0).sum().item()} max={df.max().item():.3e} at {fi}"
                      f" scenes={sorted(set((df > 0).nonzero()[:, 0].tolist()))[:12]}"
                      f" | node: n_diff={(dn > 0).sum().item()} max={dn.max().item():.3e} at {ni}"
                      f" scenes={sorted(set((dn > 0).nonzero()[:, 0].tolist()))[:12]}", flush=True)
    print(f"{c['name']:34s} interleave={int(interleave)} iters={iters} mismatches={bad} "
          f"first-run err vs oracle: fac {e_f0:.2e} node {e_n0:.2e}", flush=True)
    return bad


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    cases = [("pairwise", 11, 64, 64, 0, 1, 300), ("hyper", 11, 64, 64, 5, 1, 700), ("pairwise", 8, 64, 64, 0, 2, 130),
             ("hyper", 20, 64, 64, 8, 2, 200), ("pairwise", 5, 96, 72, 0, 1, 129)]
    built = [make(*c) for c in cases]
    total = 0
    total += stress(built[0], built[1], iters, False)
    total += stress(built[0], built[1], iters, True)
    total += stress(built[0], built[2], iters, True)
    for i in range(1, len(built)):
        total += stress(built[i], built[0], max(iters // 4, 100), True)
    print("TOTAL MISMATCHES", total)


if __name__ == "__main__":
    main()
