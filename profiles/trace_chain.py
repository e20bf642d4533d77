import sys, ctypes as C, torch
sys.path.insert(0, ".")
import groupnet_b200 as gb
from groupnet_b200 import _lib
lib = _lib.load()
torch.manual_seed(1234)
m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1).cuda().set_precision("bf16").set_rng("philox", 0)
x = torch.randn(65536, 11, 64, device="cuda")
buf = torch.zeros(2 * 8 * 16, dtype=torch.int64, device="cuda")
with torch.no_grad():
    for _ in range(2): m(x)
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(buf.data_ptr()))
    m(x)
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(0))
t = buf.cpu().view(2, 8, 16)
names = {0: "start", 1: "staged", 2: "G1 done", 4: "D1+G2 done", 6: "D2+G3 done", 8: "D3a+G4a done", 10: "D3b+G4b done", 11: "epi4 done"}
pts = [0, 1, 2, 4, 6, 8, 10, 11]
for g in range(2):
    for it in range(2, 6):
        row = t[g, it]
        d = [int(row[pts[i + 1]] - row[pts[i]]) for i in range(len(pts) - 1)]
        nxt = int(t[g, it + 1, 0] - row[11]) if it + 1 < 8 else -1
        print(f"grp{g} tile{it}: stage {d[0]:6d} | G1 {d[1]:6d} | D1+G2 {d[2]:6d} | D2+G3 {d[3]:6d} | D3a+G4a {d[4]:6d} | D3b+G4b {d[5]:6d} | epi4 {d[6]:6d} | ->next {nxt:6d} | total {int(row[11]-row[0]):6d}")
