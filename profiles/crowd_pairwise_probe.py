"""The pairwise layer at the crowd shape (N=64 -> 4,096 edges per scene, h_dim 256), reported separately at a
reduced batch as SURVEY.md 8d asks.  Usage: python profiles/crowd_pairwise_probe.py [scenes] [precision]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import groupnet_b200 as gb
from groupnet_b200 import _lib

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
n, d = 64, 256
dev = torch.device("cuda")
torch.manual_seed(1234)
layer = gb.MS_HGNN_oridinary(16, d, 64, d, batch_norm=0, nmp_layers=1).to(dev).eval()
layer.set_rng("philox", seed=0).set_precision(prec)
layer.workspace_limit_bytes = 40 << 30
x = torch.randn(b, n, d, generator=torch.Generator().manual_seed(0)).to(dev)
out = torch.empty(b, n, d, device=dev)
with torch.no_grad():
    for _ in range(2):
        layer(x, out=out, want_factors=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        layer(x, out=out, want_factors=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    _lib.profile_enable(True)
    layer(x, out=out, want_factors=False); torch.cuda.synchronize()
    prof = _lib.profile_collect(); _lib.profile_enable(False)
print(json.dumps({"scenes": b, "precision": prec, "ms_per_step": ms, "scenes_per_s": b / ms * 1e3,
                  "tflops_at_3.56GFLOP_per_scene": 3.564e9 * b / ms / 1e9}))
for k, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:22s} {t:9.3f} ms  n={c}")
