"""Occupancy sweep of the two hyper gather / scatter kernels at the NBA shape (kernel times from the library's profiling
hook).  python profiles/hbm_kernel_sweep.py  — each configuration in its own process (the knobs are read once)."""
import json
import os
import subprocess
import sys

if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import groupnet_b200 as gb
    from groupnet_b200 import _lib
    torch.manual_seed(0)
    b, n = 65536, 11
    x = torch.randn(b, n, 64, device="cuda")
    hn = torch.nn.functional.normalize(x, p=2, dim=2)
    corr = hn @ hn.transpose(1, 2)
    m = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=1, scale=5).cuda().eval().set_precision("tf32").set_rng("philox", seed=1)
    m.workspace_limit_bytes = 9 << 30
    with torch.no_grad():
        for _ in range(3):
            m(x, corr)
        _lib.profile_enable(True)
        for _ in range(5):
            m(x, corr)
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
    print(json.dumps({k: round(t / c, 4) for k, (t, c) in prof.items() if "hyper" in k and "agg" not in k}))
else:
    for cap_n2e in (72, 36, 18):
        for cap_e2n, ctas in ((48, 4), (24, 8), (12, 12), (8, 16)):
            env = dict(os.environ, GN_N2E_SMEM_KB=str(cap_n2e), GN_E2N_SMEM_KB=str(cap_e2n), GN_E2N_CTAS=str(ctas))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True)
            print(f"n2e cap {cap_n2e} KB | e2n cap {cap_e2n} KB x {ctas} CTAs/SM:", (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)
