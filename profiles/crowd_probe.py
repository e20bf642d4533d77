"""Per-kernel profile of the synthetic crowd config (BASELINE.json configs[3]):
N=64 agents, h_dim 256, hyper scales {2,4,8,16}; fused corr+top-k for the 4 scales
+ 4x MS_HGNN_hyper on one chunk of scenes.  Usage: python profiles/crowd_probe.py [scenes] [precision]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import groupnet_b200 as gb
from groupnet_b200 import _lib, ops

b = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
n, d, scales = 64, 256, (2, 4, 8, 16)
dev = torch.device("cuda")
torch.manual_seed(1234)
layers = [gb.MS_HGNN_hyper(d, d, 64, d, batch_norm=0, nmp_layers=1, scale=s).to(dev).eval() for s in scales]
for i, l in enumerate(layers):
    l.set_rng("philox", seed=i).set_precision(prec)
    l.workspace_limit_bytes = 24 << 30
x = torch.randn(b, n, d, generator=torch.Generator().manual_seed(0)).to(dev)
hcat = torch.empty(b, sum(ops.incidence_rows(n, s) for s in scales), n, device=dev)
feat = torch.empty(b, n, d * len(scales), device=dev)

def step():
    hs = ops.corr_topk_h_into(x, list(scales), hcat)
    for i, l in enumerate(layers):
        l(x, H=hs[i], out=feat[:, :, i * d:(i + 1) * d], want_factors=False)

with torch.no_grad():
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    _lib.profile_enable(True)
    step(); torch.cuda.synchronize()
    prof = _lib.profile_collect(); _lib.profile_enable(False)
print(json.dumps({"scenes": b, "precision": prec, "ms_per_step": ms, "scenes_per_s": b / ms * 1e3,
                  "tflops_at_451MFLOP_per_scene": 451e6 * b / ms / 1e9}))
for k, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:22s} {t:9.3f} ms  n={c}")
