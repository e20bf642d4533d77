"""Poor man's memcheck (not collected by pytest; compute-sanitizer is not available on the pool).

Every buffer one forward touches -- packed weights, h, noise, H, workspace, both outputs -- is
re-homed inside a larger allocation whose 64 KiB margins hold 0xFF bytes (NaN as fp32 and as
bf16).  An out-of-bounds WRITE shows up as a damaged margin; an out-of-bounds READ that reaches
arithmetic shows up as NaN (or any bit difference) against the same forward on ordinary buffers.

    python tests/stress_guard_bands.py
"""
import ctypes as C
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import groupnet_b200 as gb                      # noqa: E402
from groupnet_b200 import _lib, ops             # noqa: E402

DEV = torch.device("cuda:0")
G = 1 << 16


class Guarded:
    def __init__(self, t: torch.Tensor, name: str):
        t = t.contiguous()
        self.name = name
        self.nbytes = t.numel() * t.element_size()
        self.raw = torch.full((2 * G + self.nbytes,), 0xFF, dtype=torch.uint8, device=DEV)
        self.view = self.raw[G:G + self.nbytes].view(t.dtype).view(t.shape)
        self.view.copy_(t)

    def intact(self) -> bool:
        return bool((self.raw[:G] == 0xFF).all().item() and (self.raw[G + self.nbytes:] == 0xFF).all().item())


def build(kind, n, d, bo, scale, layers, b, precision):
    torch.manual_seed(7 + n + d + layers + b)
    if kind == "pairwise":
        m = gb.MS_HGNN_oridinary(16, d, 64, bo, batch_norm=0, nmp_layers=layers)
        e, t = n * n, 6
    else:
        m = gb.MS_HGNN_hyper(d, d, 64, bo, batch_norm=0, nmp_layers=layers, scale=scale)
        e, t = (1 if scale == n else n), 10
    m = m.to(DEV).set_precision(precision)
    h = torch.randn(b, n, d, device=DEV)
    noise = [torch.rand(b, e, t, device=DEV) for _ in range(layers)]
    inc = None
    if kind == "hyper":
        inc = ops.corr_topk_h(h, [scale])[0].contiguous()
    return m, h, noise, inc, e, t


def one_case(kind, n, d, bo, scale, layers, b, precision, limit):
    tag = f"{kind}-{n}-{d}-{bo}-{scale}-{layers}-{b}-{precision}-ws{limit >> 20}M"
    m, h, noise, inc, e, t = build(kind, n, d, bo, scale, layers, b, precision)
    m.workspace_limit_bytes = limit
    with torch.no_grad():
        node0, fac0 = m._run(h, inc, e, noise)
    torch.cuda.synchronize()
    guards = []
    for si, st in enumerate(m._packs.get(m, DEV)):
        for name in list(st.tensors):
            tens = st.tensors[name]
            if tens is None or tens.numel() == 0:
                continue
            g = Guarded(tens, f"w{si}.{name}")
            guards.append(g)
            st.tensors[name] = g.view
            if name in _lib.StageWeights.FIELDS:
                setattr(st.struct, name, C.c_void_p(g.view.data_ptr()))
    gh = Guarded(h, "h")
    gu = [Guarded(u, f"u{i}") for i, u in enumerate(noise)]
    gi = Guarded(inc, "H") if inc is not None else None
    gws = Guarded(torch.empty(m._ws.buf.numel(), dtype=torch.uint8, device=DEV), "ws")
    gws.view.fill_(0xFF)
    m._ws.buf = gws.view
    gn = Guarded(torch.empty_like(node0), "node_out")
    gf = Guarded(torch.empty_like(fac0), "dist_out")
    gn.view.fill_(float("nan"))
    gf.view.fill_(float("nan"))
    guards += [gh, gws, gn, gf] + gu + ([gi] if gi is not None else [])
    with torch.no_grad():
        m._run(gh.view, gi.view if gi is not None else None, e, [g.view for g in gu], node_out=gn.view, dist_out=gf.view)
    torch.cuda.synchronize()
    damaged = [g.name for g in guards if not g.intact()]
    same = torch.equal(gn.view, node0) and torch.equal(gf.view, fac0)
    nan = bool(torch.isnan(gn.view).any().item() or torch.isnan(gf.view).any().item())
    ok = same and not damaged
    print(f"{'ok ' if ok else 'BAD'} {tag:44s} bitwise_same={same} nan={nan} damaged_margins={damaged}", flush=True)
    return ok


def topk_case(b, n, d, scales):
    x = torch.randn(b, n, d, device=DEV)
    rows = sum(ops.incidence_rows(n, s) for s in scales)
    ref = torch.empty(b, rows, n, device=DEV)
    ops.corr_topk_h_into(x, scales, ref)
    gx = Guarded(x, "x")
    go = Guarded(torch.empty_like(ref), "H")
    go.view.fill_(float("nan"))
    ops.corr_topk_h_into(gx.view, scales, go.view)
    torch.cuda.synchronize()
    damaged = [g.name for g in (gx, go) if not g.intact()]
    same = torch.equal(go.view, ref)
    ok = same and not damaged
    print(f"{'ok ' if ok else 'BAD'} topk-{b}-{n}-{d}-{scales} bitwise_same={same} damaged_margins={damaged}", flush=True)
    return ok


def main():
    cases = [("pairwise", 11, 64, 64, 0, 1, 300), ("hyper", 11, 64, 64, 5, 1, 700), ("hyper", 11, 64, 64, 11, 1, 700),
             ("pairwise", 8, 64, 64, 0, 2, 130), ("hyper", 20, 64, 64, 8, 2, 200), ("hyper", 64, 256, 256, 4, 1, 9),
             ("pairwise", 5, 96, 72, 0, 1, 129), ("hyper", 9, 128, 64, 3, 1, 257), ("pairwise", 11, 64, 64, 0, 1, 1),
             ("hyper", 11, 64, 64, 5, 1, 1), ("pairwise", 11, 64, 64, 0, 1, 4099), ("hyper", 11, 64, 64, 5, 1, 4099),
             ("hyper", 11, 64, 64, 11, 1, 4099), ("hyper", 64, 256, 256, 16, 1, 37), ("hyper", 64, 256, 256, 64, 1, 37),
             ("pairwise", 20, 64, 64, 0, 1, 77), ("hyper", 8, 64, 64, 4, 1, 1001), ("hyper", 64, 64, 32, 8, 1, 33)]
    bad = 0
    for c in cases:
        for precision in ("fp32", "bf16"):
            for limit in (64 << 20, 4 << 30):
                try:
                    bad += 0 if one_case(*c, precision, limit) else 1
                except Exception as exc:            # shape not eligible for a precision etc.
                    print(f"ERR {c} {precision}: {type(exc).__name__}: {exc}", flush=True)
                    bad += 1
    for tk in [(300, 11, 64, [5, 11]), (1, 11, 64, [5]), (4099, 11, 64, [5, 11]), (37, 64, 256, [2, 4, 8, 16]),
               (129, 20, 64, [4, 10, 20]), (77, 8, 64, [2, 4, 8]), (5, 64, 32, [7, 64])]:
        bad += 0 if topk_case(*tk) else 1
    print("BAD CASES", bad)


if __name__ == "__main__":
    main()
