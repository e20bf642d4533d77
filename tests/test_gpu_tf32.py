"""fp32-grade tensor-core path (precision "tf32": 3xTF32 tcgen05 chains, csrc/gn_chain_tf32.cu) against the
reference outputs (goldens) and the oracle at the fp32 criterion: max|d| <= 1e-5 * max|ref| per output tensor."""
import pytest
import torch

import groupnet_b200 as gb
from oracle import ms_hgnn_oracle as O
from helpers import FP32_REL, assert_close, build_layer, diagnose_oracle_mismatch, golden_names, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _inference_mode():
    with torch.no_grad():
        yield


def _run_layer(g):
    m = build_layer(g).to(DEV).set_precision("tf32")
    h = torch.from_numpy(g["h"]).to(DEV)
    noise = [torch.from_numpy(u).to(DEV) for u in g["noise"]]
    if g["kind"] == "pairwise":
        node, fac = m(h, noise=noise)
        hinc = None
    else:
        node, fac, hinc = m(h, torch.from_numpy(g["corr"]).to(DEV), noise=noise)
    torch.cuda.synchronize()
    return node, fac, hinc


@pytest.mark.parametrize("name", golden_names())
def test_layer_forward_tf32_vs_golden(name):
    g = load_golden(name)
    node, fac, hinc = _run_layer(g)
    if hinc is not None:
        assert torch.equal(hinc.cpu(), torch.from_numpy(g["H"]))
    assert_close(fac, g["factors"], FP32_REL, f"{name} factors")
    assert_close(node, g["node_feat"], FP32_REL, f"{name} node_feat")


@pytest.mark.parametrize("kind,n,d,bo,scale,layers,b", [
    ("pairwise", 11, 64, 64, 0, 1, 300), ("hyper", 11, 64, 64, 5, 1, 700), ("hyper", 11, 64, 64, 11, 1, 700),
    ("pairwise", 8, 64, 64, 0, 2, 130), ("hyper", 20, 64, 64, 8, 2, 200), ("hyper", 64, 256, 256, 4, 1, 9),
    ("pairwise", 5, 96, 72, 0, 1, 129), ("hyper", 9, 128, 64, 3, 1, 257), ("pairwise", 20, 64, 128, 0, 1, 33),
    ("pairwise", 13, 32, 32, 0, 1, 77), ("hyper", 7, 32, 96, 2, 3, 150),
])
def test_layer_forward_tf32_vs_oracle(kind, n, d, bo, scale, layers, b):
    torch.manual_seed(2000 + n + d + layers)
    if kind == "pairwise":
        m = gb.MS_HGNN_oridinary(16, d, 64, bo, batch_norm=0, nmp_layers=layers)
        e, t = n * n, 6
    else:
        m = gb.MS_HGNN_hyper(d, d, 64, bo, batch_norm=0, nmp_layers=layers, scale=scale)
        e, t = (1 if scale == n else n), 10
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(b)
    h = torch.randn(b, n, d, generator=gen)
    noise = [torch.rand(b, e, t, generator=gen) for _ in range(layers)]
    m = m.to(DEV).set_precision("tf32")
    corr = None if kind == "pairwise" else O.feature_correlation(h)
    ref_h = None if kind == "pairwise" else O.incidence_topk(corr, scale)

    def run_oracle(dtype=torch.float32, threads=None):
        keep = torch.get_num_threads()
        if threads:
            torch.set_num_threads(threads)
        try:
            sdd = {k: v.to(dtype) for k, v in sd.items()}
            nz = [u.to(dtype) for u in noise]
            if kind == "pairwise":
                return O.forward_pairwise(sdd, h.to(dtype), nz, nmp_layers=layers)
            return O.forward_hyper(sdd, h.to(dtype), corr.to(dtype), scale, nz, nmp_layers=layers, h_inc=ref_h.to(dtype))[:2]
        finally:
            torch.set_num_threads(keep)

    def run_gpu():
        out = m(h.to(DEV), noise=noise) if kind == "pairwise" else m(h.to(DEV), corr.to(DEV), noise=noise)
        torch.cuda.synchronize()
        return tuple(o.cpu() for o in out)

    ref_node, ref_fac = run_oracle()
    got = run_gpu()
    if kind != "pairwise":
        assert torch.equal(got[2], ref_h)
    try:
        assert_close(got[1], ref_fac, FP32_REL, "factors")
        assert_close(got[0], ref_node, FP32_REL, "node_feat")
    except AssertionError as first:
        diagnose_oracle_mismatch(first, run_oracle, run_gpu, (ref_node, ref_fac), got[:2])


@pytest.mark.parametrize("kind,b", [("pairwise", 2500), ("hyper5", 30000), ("hyper11", 40000)])
def test_tf32_many_tiles_vs_fp32_path(kind, b):
    """More tiles than SMs (persistent CTAs loop, ring / barrier phases wrap many times); Philox noise on the
    device so both precisions draw the same stream."""
    torch.manual_seed(11)
    n, d = 11, 64
    if kind == "pairwise":
        m = gb.MS_HGNN_oridinary(16, d, 64, 64, batch_norm=0, nmp_layers=1)
    else:
        m = gb.MS_HGNN_hyper(d, d, 64, 64, batch_norm=0, nmp_layers=1, scale=5 if kind == "hyper5" else 11)
    m = m.to(DEV).eval()
    h = torch.randn(b, n, d, device=DEV)
    hn = torch.nn.functional.normalize(h, p=2, dim=2)
    corr = hn @ hn.transpose(1, 2)
    outs = {}
    for prec in ("fp32", "tf32"):
        m.set_precision(prec).set_rng("philox", seed=5)
        outs[prec] = m(h) if kind == "pairwise" else m(h, corr)[:2]
    torch.cuda.synchronize()
    assert_close(outs["tf32"][1], outs["fp32"][1], FP32_REL, "factors")
    assert_close(outs["tf32"][0], outs["fp32"][0], FP32_REL, "node_feat")
    # determinism of the tensor-core path
    m.set_precision("tf32").set_rng("philox", seed=5)
    again = m(h) if kind == "pairwise" else m(h, corr)[:2]
    assert torch.equal(again[0], outs["tf32"][0]) and torch.equal(again[1], outs["tf32"][1])


@pytest.mark.parametrize("precision,tol", [("tf32", FP32_REL), ("bf16", 2e-2)])
@pytest.mark.parametrize("kind,b", [("pairwise", 300), ("hyper5", 1800), ("hyper11", 19200)])
def test_tensor_core_paths_vs_oracle_at_more_tiles_than_sms(kind, b, precision, tol):
    """Both tensor-core paths against the ORACLE (not the repo's own fp32 path) at NBA shape with more 128-row tiles
    than the GPU has SMs, so every persistent CTA loops: pairwise 300 x 121 = 284 edge tiles; scale 5: 1,800 x 11 =
    155 tiles of edge / node rows; scale 11: 19,200 node rows = 150 tiles (one all-ones hyperedge per scene)."""
    torch.manual_seed(4000 + b)
    n, d = 11, 64
    if kind == "pairwise":
        m = gb.MS_HGNN_oridinary(16, d, 64, 64, batch_norm=0, nmp_layers=1)
        e, t, scale = n * n, 6, 0
    else:
        scale = 5 if kind == "hyper5" else 11
        m = gb.MS_HGNN_hyper(d, d, 64, 64, batch_norm=0, nmp_layers=1, scale=scale)
        e, t = (1 if scale == n else n), 10
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(b)
    h = torch.randn(b, n, d, generator=gen)
    noise = [torch.rand(b, e, t, generator=gen)]
    m = m.to(DEV).set_precision(precision)
    corr = None if kind == "pairwise" else O.feature_correlation(h)
    ref_h = None if kind == "pairwise" else O.incidence_topk(corr, scale)

    def run_oracle(dtype=torch.float32, threads=None):
        keep = torch.get_num_threads()
        if threads:
            torch.set_num_threads(threads)
        try:
            sdd = {k: v.to(dtype) for k, v in sd.items()}
            nz = [u.to(dtype) for u in noise]
            if kind == "pairwise":
                return O.forward_pairwise(sdd, h.to(dtype), nz)
            return O.forward_hyper(sdd, h.to(dtype), corr.to(dtype), scale, nz, h_inc=ref_h.to(dtype))[:2]
        finally:
            torch.set_num_threads(keep)

    def run_gpu():
        out = m(h.to(DEV), noise=noise) if kind == "pairwise" else m(h.to(DEV), corr.to(DEV), noise=noise)
        torch.cuda.synchronize()
        return tuple(o.cpu() for o in out)

    ref_node, ref_fac = run_oracle()
    got = run_gpu()
    if kind != "pairwise":
        assert torch.equal(got[2], ref_h)
    try:
        assert_close(got[1], ref_fac, tol, f"{kind} {precision} factors")
        assert_close(got[0], ref_node, tol, f"{kind} {precision} node_feat")
    except AssertionError as first:
        diagnose_oracle_mismatch(first, run_oracle, run_gpu, (ref_node, ref_fac), got[:2], tol=tol)
