"""Phase timeline (clock64) of hyper_fused64_tc_kernel, block 0 (build csrc with -DGN_ENABLE_TRACE)."""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import groupnet_b200 as gb
from groupnet_b200 import _lib, ops
lib = _lib.load()
b, n, d = 65536, 11, 64
torch.manual_seed(1234)
m = gb.MS_HGNN_hyper(d, d, 64, d, batch_norm=0, nmp_layers=1, scale=5).cuda().eval().set_precision("bf16").set_rng("philox", 0)
m.workspace_limit_bytes = 24 << 30
x = torch.randn(b, n, d, device="cuda")
hcat = torch.empty(b, n, n, device="cuda")
buf = torch.zeros(8 * 16, dtype=torch.int64, device="cuda")
names = ["start", "raw H + Hblk", "hT + arrive", "eo_full", "eo drained", "main loop", "ef_full", "raw H + HblkT + efT",
         "inc h part", "agg_full", "inc agg part", "o1_full", "o1 drained", "out_full", "out stored"]
with torch.no_grad():
    hs = ops.corr_topk_h_into(x, [5], hcat)
    for _ in range(2):
        m(x, H=hs[0], want_factors=False)
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(buf.data_ptr()))
    m(x, H=hs[0], want_factors=False)
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(0))
t = buf.cpu().view(8, 16)
for it in range(2, 5):
    row = t[it]
    print(f"tile {it}: total {int(t[it + 1, 0] - row[0])}")
    for i in range(1, 15):
        print(f"   {names[i]:22s} +{int(row[i] - row[i - 1]):7d}")
