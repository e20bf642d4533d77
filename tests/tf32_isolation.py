"""Per-kernel isolation of the 3xTF32 chains (not collected by pytest; run on a B200):

    python tests/tf32_isolation.py            # every case in its own subprocess, one line per case
    python tests/tf32_isolation.py CASE KEEP  # one case in this process

The stage driver falls back to the FFMA kernel of a step whose tf_* weight stream is NULL, so clearing all but
one stream runs exactly one tensor-core chain inside an otherwise fp32 stage; its outputs are compared with the
all-FFMA run of the same inputs (bound 1e-5 * max|ref|, the fp32 parity criterion)."""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

STREAMS = ("tf_chain_w", "tf_pre_w", "tf_aggin_w", "tf_aggout_w", "tf_hagg_w", "tf_post_w", "tf_pagg_w")
CASES = {
    "pair11": ("pairwise", 11, 64, 64, 0, 300),
    "pair8": ("pairwise", 8, 64, 64, 0, 500),
    "pair20": ("pairwise", 20, 64, 64, 0, 40),
    "hyper5": ("hyper", 11, 64, 64, 5, 700),
    "hyper11": ("hyper", 11, 64, 64, 11, 700),
    "hyper20": ("hyper", 20, 64, 32, 8, 200),
}
KEEPS = ("tf_pre_w", "tf_pre_w+tf_aggin_w", "tf_pre_w+tf_pagg_w", "tf_chain_w", "tf_hagg_w", "tf_post_w", "tf_aggout_w+tf_post_w", "all")


def run_case(case: str, keep: str) -> None:
    import torch
    import groupnet_b200 as gb
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import rel_err
    kind, n, d, bo, scale, b = CASES[case]
    torch.manual_seed(7)
    if kind == "pairwise":
        m = gb.MS_HGNN_oridinary(16, d, 64, bo, batch_norm=0, nmp_layers=1)
        e, t = n * n, 6
    else:
        m = gb.MS_HGNN_hyper(d, d, 64, bo, batch_norm=0, nmp_layers=1, scale=scale)
        e, t = (1 if scale == n else n), 10
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(b)
    h = torch.randn(b, n, d, generator=gen).cuda()
    u = [torch.rand(b, e, t, generator=gen).cuda()]
    corr = None
    if kind != "pairwise":
        hn = torch.nn.functional.normalize(h, p=2, dim=2)
        corr = hn @ hn.transpose(1, 2)
    with torch.no_grad():
        ref = m(h, noise=u) if corr is None else m(h, corr, noise=u)
        torch.cuda.synchronize()
        m.set_precision("tf32")
        stages = m._packs.get(m, h.device)
        kept = STREAMS if keep == "all" else tuple(keep.split("+"))
        for st in stages:
            for name in STREAMS:
                if name not in kept:
                    setattr(st.struct, name, C.c_void_p(0))
        got = m(h, noise=u) if corr is None else m(h, corr, noise=u)
        torch.cuda.synchronize()
    en, ef = rel_err(got[0], ref[0]), rel_err(got[1], ref[1])
    ok = en <= 1e-5 and ef <= 1e-5
    print(f"RESULT {case:8s} keep={keep:24s} node {en:.3e} factors {ef:.3e} {'ok' if ok else 'FAIL'}", flush=True)


def main() -> None:
    if len(sys.argv) == 3:
        run_case(sys.argv[1], sys.argv[2])
        return
    for case, spec in CASES.items():
        for keep in KEEPS:
            pw = spec[0] == "pairwise"
            if ("agg" in keep and "hagg" not in keep and not pw) or ("hagg" in keep and pw):
                continue
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), case, keep], capture_output=True,
                                   text=True, timeout=240)
                lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")]
                if lines:
                    print(lines[-1], flush=True)
                else:
                    tail = (r.stderr or r.stdout).strip().splitlines()[-3:]
                    print(f"RESULT {case:8s} keep={keep:24s} CRASH rc={r.returncode} :: {' | '.join(tail)}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"RESULT {case:8s} keep={keep:24s} TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
