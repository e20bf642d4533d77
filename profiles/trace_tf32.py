"""Phase timeline of the 3xTF32 chain engine (csrc/gn_chain_tf32.cu): clock64 stamps of block 0's first tiles for the
pairwise edge chain (or `python profiles/trace_tf32.py hyper`), printed per op for the issuer and per event for row
thread 0.  Run on a B200."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import groupnet_b200 as gb
from groupnet_b200 import _lib

MAX_OPS, MAX_EV, TR_TILES = 64, 72, 6
TR_ROWS, TR_MAXCH = 3 * MAX_OPS, 40
TR_CHUNK = TR_ROWS + 3 * MAX_EV
TR_STAGE = TR_CHUNK + 4 * TR_MAXCH
TR_SLOTS = TR_STAGE + 8
kind = sys.argv[1] if len(sys.argv) > 1 else "pair"
keep = sys.argv[2] if len(sys.argv) > 2 else "all"
os.environ["GN_TRACE_KERNEL"] = sys.argv[3] if len(sys.argv) > 3 else ("edge_chain_pair_tf32" if kind == "pair" else "edge_chain_tf32")
torch.manual_seed(1234)
if kind == "pair":
    m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1)
else:
    m = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=1, scale=5)
m = m.cuda().eval().set_precision("tf32").set_rng("philox", seed=1)
b, n = 8192, 11
x = torch.randn(b, n, 64, device="cuda")
hn = torch.nn.functional.normalize(x, p=2, dim=2)
corr = hn @ hn.transpose(1, 2)
lib = _lib.load()
buf = torch.zeros(TR_TILES * TR_SLOTS, dtype=torch.int64, device="cuda")
STREAMS = ("tf_chain_w", "tf_pre_w", "tf_aggin_w", "tf_aggout_w", "tf_hagg_w", "tf_post_w", "tf_pagg_w")
with torch.no_grad():
    run = (lambda: m(x)) if kind == "pair" else (lambda: m(x, corr))
    run()
    for st in m._packs.get(m, x.device):          # only the traced chain on the tensor cores: one engine launch writes the buffer
        for name in STREAMS:
            if keep != "all" and name not in keep.split("+"):
                setattr(st.struct, name, C.c_void_p(0))
    run()
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(buf.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.gn_profile_set_trace(C.c_void_p(0))
t = buf.cpu().view(TR_TILES, TR_SLOTS)
print(f"trace of {kind} / {keep}: clk relative to the tile's first stamp; issuer: (operands ready, weights landed, issued); rows: (start, wait done, end)")
for it in range(1, TR_TILES):
    row = t[it]
    base = int(min(v for v in row.tolist() if v > 0))
    iss = row[:TR_ROWS].view(MAX_OPS, 3)
    ev = row[TR_ROWS:TR_CHUNK].view(MAX_EV, 3)
    ch = row[TR_CHUNK:TR_STAGE].view(TR_MAXCH, 4)
    stg = row[TR_STAGE:]
    print(f"tile iter {it}: previous tile started {base - int(min(v for v in t[it - 1].tolist() if v > 0))} clk earlier")
    print("  issuer:", " | ".join(f"op{o}: {int(iss[o,0])-base:6d} {int(iss[o,1])-base:6d} {int(iss[o,2])-base:6d}"
                                   for o in range(MAX_OPS) if int(iss[o, 0]) > 0))
    print("  rows  :", " | ".join(f"ev{e}: {int(ev[e,0])-base:6d} {int(ev[e,1])-base:6d} {int(ev[e,2])-base:6d}"
                                   for e in range(MAX_EV) if int(ev[e, 0]) > 0))
    print("  chunks (producer copy issued | issuer wait start, wait end, MMAs issued):")
    print("   ", " | ".join(f"c{c}: {int(ch[c,0])-base:6d} / {int(ch[c,1])-base:6d} {int(ch[c,2])-base:6d} {int(ch[c,3])-base:6d}"
                            for c in range(TR_MAXCH) if int(ch[c, 1]) > 0))
    if int(stg[0]) > 0:
        print("  staging sub-stamps (nodes landed, partials written, barrier, weights, A0 written, fenced, barrier, prefetch issued):",
              " ".join(f"{int(v)-base:6d}" for v in stg.tolist()))
