"""Phase timeline of the fused pairwise aggregation (csrc/gn_pair_agg_tf32.cu): clock64 stamps of block 0 / thread 0
for its first tiles.  Run on a B200: python profiles/trace_pair_agg_tf32.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GN_TRACE_KERNEL"] = "pair_agg_tf32"
import torch
import groupnet_b200 as gb
from groupnet_b200 import _lib

TR_TILES, TR_SLOTS = 6, 128
torch.manual_seed(1234)
m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1).cuda().eval().set_precision("tf32").set_rng("philox", seed=1)
x = torch.randn(8192, 11, 64, device="cuda")
lib = _lib.load()
buf = torch.zeros(TR_TILES * TR_SLOTS, dtype=torch.int64, device="cuda")
with torch.no_grad():
    m(x)
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(buf.data_ptr()))
    m(x)
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(0))
t = buf.cpu().view(TR_TILES, TR_SLOTS)
for it in range(1, TR_TILES):
    r = t[it].tolist()
    base = r[0]
    print(f"tile iter {it}: previous tile started {base - t[it - 1][0].item()} clk earlier")
    print("  prologue: h staged %d | s table written %d | barrier %d | S done %d" % tuple(r[i] - base for i in (1, 2, 3, 4)))
    for u in range(12):
        s = r[8 + 8 * u: 8 + 8 * u + 7]
        print(f"  u{u:2d}: start {s[0]-base:6d} | partial u-2 +{s[1]-s[0]:5d} | relu-sum +{s[2]-s[1]:5d} | barrier +{s[3]-s[2]:5d}"
              f"   || drain warp: G(u-1) -> TMEM done at {s[4]-base if s[4] else 0:6d}, P(u+1) -> smem done at {s[5]-base if s[5] else 0:6d}")
    print("  epilogue: start %d | G(U-1) written %d | bias done %d | stored %d | end %d" % tuple(r[i] - base for i in (120, 121, 122, 123, 124)))
