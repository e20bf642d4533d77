"""Phase timeline (clock64) of hyper_fused_tc_kernel, block 0, via gn_profile_set_trace."""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import groupnet_b200 as gb
from groupnet_b200 import _lib, ops
lib = _lib.load()
b, n, d = 8192, 64, 256
torch.manual_seed(1234)
m = gb.MS_HGNN_hyper(d, d, 64, d, batch_norm=0, nmp_layers=1, scale=8).cuda().eval().set_precision("bf16").set_rng("philox", 0)
m.workspace_limit_bytes = 24 << 30
x = torch.randn(b, n, d, device="cuda")
hcat = torch.empty(b, n, n, device="cuda")
buf = torch.zeros(8 * 16 + 8 * 4, dtype=torch.int64, device="cuda")
names = ["start", "Hblk", "hT+arrive", "eo_full", "eo drained", "main loop", "ef_full", "HblkT", "ef transposed",
         "A_h staged", "agg_full", "A_agg", "o1_full", "o1 drained", "out_full", "out stored"]
with torch.no_grad():
    hs = ops.corr_topk_h_into(x, [8], hcat)
    for _ in range(2):
        m(x, H=hs[0], want_factors=False)
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(buf.data_ptr()))
    m(x, H=hs[0], want_factors=False)
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(0))
w = buf.cpu()[128:].view(8, 4)
t = buf.cpu()[:128].view(8, 16)
for it in range(1, 5):
    row = t[it]
    print(f"tile {it}: total {int(t[it + 1, 0] - row[0])}")
    for i in range(1, 16):
        print(f"   {names[i]:16s} +{int(row[i] - row[i - 1]):7d}")
    print("   issuer waits in main loop: hid_free %d | w_full(G1) %d | a2_full %d | w_full(G2) %d" % tuple(int(v) for v in w[it]))
