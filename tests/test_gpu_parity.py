"""Parity of the CUDA path (through the C ABI) against the golden vectors from
the live reference and against the CPU oracle.  Run on the B200 box: -m gpu."""
import numpy as np
import pytest
import torch

import groupnet_b200 as gb
from groupnet_b200 import _lib
from helpers import (BF16_REL, FP32_REL, assert_close, build_layer, build_past_encoder, golden_names, diagnose_oracle_mismatch,
                     load_golden, rel_err)
from oracle import ms_hgnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _inference_mode(request):
    """Forward parity tests run like the reference's eval scripts (torch.set_grad_enabled(False),
    test_nba.py:571); tests marked `train` keep autograd on."""
    if "train" in request.keywords:
        yield
    else:
        with torch.no_grad():
            yield


def test_native_library_is_loaded():
    lib = _lib.load()
    assert lib.gn_abi_version() == _lib.ABI_VERSION == 6
    maps = open("/proc/self/maps").read()
    assert "libgroupnet_b200.so" in maps


# ---- T1: H from the reference's corr, bit-exact ---------------------------------
@pytest.mark.parametrize("name", [n for n in golden_names() if "hyper" in n])
def test_topk_h_bit_exact_vs_golden(name):
    g = load_golden(name)
    h = gb.topk_h(torch.from_numpy(g["corr"]).to(DEV), g["scale"])
    assert h.shape == g["H"].shape
    assert torch.equal(h.cpu(), torch.from_numpy(g["H"]))


@pytest.mark.parametrize("n,d,scales", [(11, 64, [5, 11]), (8, 64, [3, 5, 8]), (20, 64, [5, 8]),
                                        (64, 256, [2, 4, 8, 16]), (5, 32, [0, 1, 5]), (33, 128, [7, 32]),
                                        (2, 8, [1, 2]), (1, 64, [1])])
def test_topk_h_vs_oracle_random(n, d, scales):
    gen = torch.Generator().manual_seed(n * 1000 + d)
    x = torch.randn(257, n, d, generator=gen)
    corr = O.feature_correlation(x)
    for s in scales:
        ref = O.incidence_topk(corr, s)
        got = gb.topk_h(corr.to(DEV), s).cpu()
        gap = O.topk_gap(corr, s)
        ok = (got == ref).all(dim=2)
        if ref.shape[1] == n:
            bad = ~ok & (gap > 0)
            assert not bad.any(), f"{int(bad.sum())} tie-free rows differ (n={n}, scale={s})"
            assert torch.equal(got, torch.from_numpy(O.incidence_topk_lowest_index(corr.numpy(), s)))
        else:
            assert ok.all()


# ---- T2: constructed ties ---------------------------------------------------------
def test_topk_ties_lowest_index():
    corr = torch.zeros(3, 6, 6)
    corr[0] = 0.5                      # everything tied
    corr[1, :, 3:] = 1.0               # tie among the top three
    corr[2] = torch.tensor([0.1, 0.9, 0.9, 0.1, 0.9, 0.9]).repeat(6, 1)
    for k in (1, 2, 3, 4):
        got = gb.topk_h(corr.to(DEV), k).cpu().numpy()
        ref = O.incidence_topk_lowest_index(corr.numpy(), k)
        assert np.array_equal(got, ref)
    # zero feature rows normalise to 0 and tie with everything: lowest indices win
    x = torch.zeros(2, 5, 16)
    hs = gb.corr_topk_h(x.to(DEV), [2])
    assert hs[0].cpu()[0, 0].tolist() == [1, 1, 0, 0, 0]


def test_scale_errors_and_edge_cases():
    corr = torch.eye(4).repeat(2, 1, 1).to(DEV)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        gb.topk_h(corr, 5)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        gb.corr_topk_h(torch.randn(2, 4, 8, device=DEV), [2, 9])
    assert gb.topk_h(corr, 4).shape == (2, 1, 4)
    assert torch.equal(gb.topk_h(corr, 0), gb.topk_h(corr, 1))
    assert torch.equal(gb.topk_h(corr, 1).cpu(), torch.eye(4).repeat(2, 1, 1))
    empty = gb.topk_h(torch.empty(0, 4, 4, device=DEV), 2)
    assert empty.shape == (0, 4, 4)
    with pytest.raises(RuntimeError, match="Float"):
        gb.topk_h(corr.double(), 2)


# ---- T3: fused corr + top-k --------------------------------------------------------
@pytest.mark.parametrize("n,d,scales,b", [(11, 64, [5, 11], 1000), (8, 64, [3, 5, 8], 333),
                                          (20, 64, [5, 8], 130), (64, 256, [2, 4, 8, 16], 70),
                                          (7, 36, [2, 7], 50), (3, 4, [1, 2], 9)])
def test_fused_corr_topk(n, d, scales, b):
    gen = torch.Generator().manual_seed(b)
    x = torch.randn(b, n, d, generator=gen)
    (hs, full), corr = gb.corr_topk_h(x.to(DEV), scales, concat=True, return_corr=True)
    corr = corr.cpu()
    ref_corr = O.feature_correlation(x)
    assert (corr - ref_corr).abs().max().item() <= 2e-6
    assert full.shape == (b, sum(1 if s == n else n for s in scales), n)
    off = 0
    for s, hm in zip(scales, hs):
        hm = hm.cpu()
        # self-consistency: H == topk(own corr) with the documented tie rule
        assert np.array_equal(hm.numpy(), O.incidence_topk_lowest_index(corr.numpy(), s))
        # vs the reference corr: only rows whose k/k+1 gap is within the corr error may differ
        ref = O.incidence_topk(ref_corr, s)
        if hm.shape[1] == n:
            bad = ~(hm == ref).all(dim=2)
            assert (O.topk_gap(ref_corr, s)[bad] < 4e-6).all()
            assert hm.sum(2).eq(max(s, 1)).all()
        assert torch.equal(full[:, off:off + hm.shape[1]].cpu(), hm)
        off += hm.shape[1]
    sep = gb.corr_topk_h(x.to(DEV), scales)
    for a, c in zip(sep, hs):
        assert torch.equal(a, c.contiguous())


# ---- T4: layer forward vs golden (reference outputs) ---------------------------------
def _run_layer(g, precision="fp32"):
    m = build_layer(g).to(DEV).set_precision(precision)
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
def test_layer_forward_fp32_vs_golden(name):
    g = load_golden(name)
    node, fac, hinc = _run_layer(g)
    if hinc is not None:
        assert torch.equal(hinc.cpu(), torch.from_numpy(g["H"]))
    assert node.shape == g["node_feat"].shape and fac.shape == g["factors"].shape
    assert_close(fac, g["factors"], FP32_REL, f"{name} factors")
    assert_close(node, g["node_feat"], FP32_REL, f"{name} node_feat")
    assert torch.allclose(fac.sum(-1), torch.ones_like(fac[..., 0]), atol=1e-5)


# ---- larger seeded batches vs the oracle (sizes the oracle finishes in seconds) -------
@pytest.mark.parametrize("kind,n,d,bo,scale,layers,b", [
    ("pairwise", 11, 64, 64, 0, 1, 300), ("hyper", 11, 64, 64, 5, 1, 700), ("hyper", 11, 64, 64, 11, 1, 700),
    ("pairwise", 8, 64, 64, 0, 2, 130), ("hyper", 20, 64, 64, 8, 2, 200), ("hyper", 64, 256, 256, 4, 1, 9),
    ("pairwise", 5, 96, 72, 0, 1, 129), ("hyper", 9, 128, 64, 3, 1, 257),
])
def test_layer_forward_fp32_vs_oracle(kind, n, d, bo, scale, layers, b):
    torch.manual_seed(1000 + n + d + layers)
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
    m = m.to(DEV)
    m.workspace_limit_bytes = 64 << 20          # force batch chunking
    corr = None if kind == "pairwise" else O.feature_correlation(h)
    ref_h = None if kind == "pairwise" else O.incidence_topk(corr, scale)

    def run_oracle(dtype=torch.float32, threads=None):
        keep = torch.get_num_threads()
        if threads:
            torch.set_num_threads(threads)
        try:
            with torch.no_grad():
                sdd = {k: v.to(dtype) for k, v in sd.items()}
                nz = [u.to(dtype) for u in noise]
                if kind == "pairwise":
                    return O.forward_pairwise(sdd, h.to(dtype), nz, nmp_layers=layers)
                return O.forward_hyper(sdd, h.to(dtype), corr.to(dtype), scale, nz, nmp_layers=layers,
                                       h_inc=ref_h.to(dtype))[:2]
        finally:
            torch.set_num_threads(keep)

    def run_gpu():
        with torch.no_grad():
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
        # say which side moved (float64 evaluation of the oracle as the arbiter)
        diagnose_oracle_mismatch(first, run_oracle, run_gpu, (ref_node, ref_fac), got[:2])


# ---- RNG contracts ------------------------------------------------------------------
def test_cpu_compat_rng_consumes_reference_stream():
    g = load_golden("l2_pairwise")
    m = build_layer(g).to(DEV).set_rng("cpu-compat")
    h = torch.from_numpy(g["h"]).to(DEV)
    torch.manual_seed(4242)
    node, fac = m(h)
    torch.manual_seed(4242)
    noise = [torch.rand(s) for s in O.noise_shapes(g["B"], g["N"], True, 0, g["L"])]
    after = torch.rand(1)
    node2, fac2 = m(h, noise=noise)
    assert torch.equal(node, node2) and torch.equal(fac, fac2)
    torch.manual_seed(4242)
    m(h)
    assert torch.equal(torch.rand(1), after)     # exactly L draws of (B,E,T) were consumed


def test_philox_statistics_and_sharding_invariance():
    g = load_golden("nba_hyper5")
    m = build_layer(g).to(DEV)
    gen = torch.Generator().manual_seed(5)
    h = torch.randn(4096, 11, 64, generator=gen).to(DEV)
    corr = O.feature_correlation(h.cpu()).to(DEV)
    m.set_rng("philox", seed=77)
    node, fac, _ = m(h, corr)
    m.set_rng("philox", seed=77)
    node_b, fac_b, _ = m(h, corr)
    assert torch.equal(node, node_b) and torch.equal(fac, fac_b)     # deterministic
    m.set_rng("philox", seed=78)
    assert not torch.equal(m(h, corr)[1], fac)                        # seed matters
    # T8: the same global batch processed as 1, 2 or 4 shards gives bit-identical scenes
    for shards in (2, 4):
        parts_n, parts_f = [], []
        step = 4096 // shards
        for r in range(shards):
            m.set_rng("philox", seed=77, scene_offset=r * step)
            a, f, _ = m(h[r * step:(r + 1) * step], corr[r * step:(r + 1) * step])
            parts_n.append(a); parts_f.append(f)
        assert torch.equal(torch.cat(parts_n), node) and torch.equal(torch.cat(parts_f), fac)
    # the implied uniforms are uniform: recover the categorical's entropy range and row sums
    assert torch.allclose(fac.sum(-1), torch.ones_like(fac[..., 0]), atol=1e-5)
    # successive calls draw fresh noise
    m.set_rng("philox", seed=77)
    f1 = m(h, corr)[1]
    f2 = m(h, corr)[1]
    assert torch.equal(f1, fac) and not torch.equal(f2, f1)


def test_philox_matches_distribution_of_reference_noise():
    """Mean categorical under device Philox noise agrees with the mean under torch.rand noise."""
    g = load_golden("nba_pairwise")
    m = build_layer(g).to(DEV)
    gen = torch.Generator().manual_seed(9)
    h = torch.randn(2048, 11, 64, generator=gen).to(DEV)
    m.set_rng("philox", seed=3)
    f_dev = m(h)[1].mean(dim=(0, 1))
    m.set_rng("cpu-compat")
    torch.manual_seed(3)
    f_cpu = m(h)[1].mean(dim=(0, 1))
    assert (f_dev - f_cpu).abs().max().item() < 5e-3


# ---- full-size properties (BASELINE.json shapes) ----------------------------------------
def test_full_size_nba_properties():
    b, n, d = 65536, 11, 64
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(b, n, d, generator=gen).to(DEV)
    (hs, full), corr = gb.corr_topk_h(x, [5, 11], concat=True, return_corr=True)
    assert full.shape == (b, 12, 11)
    assert hs[0].sum(2).eq(5).all() and hs[1].eq(1).all()
    assert torch.diagonal(hs[0], dim1=1, dim2=2).eq(1).all()          # corr[i,i] is the row max
    assert (corr - corr.transpose(1, 2)).abs().max().item() == 0.0    # Gram symmetry is exact
    assert torch.equal(gb.topk_h(corr, 5), hs[0].contiguous())        # idempotent: topk(own corr)
    torch.manual_seed(1234)
    pair = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1).to(DEV).set_rng("philox", 0)
    hyp = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=1, scale=5).to(DEV).set_rng("philox", 0)
    node, fac = pair(x)
    assert node.shape == (b, n, 64) and fac.shape == (b, 121, 6)
    assert torch.isfinite(node).all() and torch.allclose(fac.sum(-1), torch.ones_like(fac[..., 0]), atol=1e-5)
    node_h, fac_h, hm = hyp(x, corr, H=hs[0])
    assert torch.isfinite(node_h).all() and torch.allclose(fac_h.sum(-1), torch.ones_like(fac_h[..., 0]), atol=1e-5)
    # permutation equivariance over scenes: reversing the batch reverses the outputs (injected noise)
    sl = slice(0, 512)
    u = torch.rand(512, 121, 6, generator=gen).to(DEV)
    a, fa = pair(x[sl], noise=[u])
    bwd, fb = pair(x[sl].flip(0).contiguous(), noise=[u.flip(0).contiguous()])
    assert torch.equal(a, bwd.flip(0)) and torch.equal(fa, fb.flip(0))


def test_inputs_not_mutated_and_fresh_outputs():
    g = load_golden("fish8_hyper3")
    m = build_layer(g).to(DEV)
    h = torch.from_numpy(g["h"]).to(DEV)
    corr = torch.from_numpy(g["corr"]).to(DEV)
    h0, c0 = h.clone(), corr.clone()
    noise = [torch.from_numpy(u).to(DEV) for u in g["noise"]]
    a1, f1, hm1 = m(h, corr, noise=noise)
    a2, f2, hm2 = m(h, corr, noise=noise)
    assert torch.equal(h, h0) and torch.equal(corr, c0)
    assert a1.data_ptr() != a2.data_ptr() and torch.equal(a1, a2) and torch.equal(f1, f2)
    assert not hm1.requires_grad
    with pytest.raises(RuntimeError, match="Float"):
        m(h.double(), corr.double())


# ---- the block the encoders run: MultiScaleInteraction (GroupNet_nba.py:284-309) -------
def test_multiscale_interaction_matches_layers_and_host_pipeline():
    torch.manual_seed(1234)
    m = gb.MultiScaleInteraction(64, (5, 11)).to(DEV).eval()
    gen = torch.Generator().manual_seed(3)
    b, n = 1000, 11
    x = torch.randn(b, n, 64, generator=gen)
    noise = [torch.rand(b, 121, 6, generator=gen), torch.rand(b, 11, 10, generator=gen),
             torch.rand(b, 1, 10, generator=gen)]
    feat, new_h = m(x.to(DEV), noise=[u.to(DEV) for u in noise])
    assert feat.shape == (b, n, 256) and new_h.shape == (b, 12, 11)
    # against the oracle, layer by layer, concatenated the way PastEncoder does
    sds = [{k: v.detach().cpu() for k, v in l.state_dict().items()} for l in m.layers()]
    with torch.no_grad():
        corr = O.feature_correlation(x)
        a, _ = O.forward_pairwise(sds[0], x, [noise[0]])
        h5, _, H5 = O.forward_hyper(sds[1], x, corr, 5, [noise[1]])
        h11, _, H11 = O.forward_hyper(sds[2], x, corr, 11, [noise[2]])
    ref = torch.cat((x, a, h5, h11), dim=-1)
    assert torch.equal(new_h.cpu(), torch.cat((H5, H11), dim=1))
    assert torch.equal(feat[:, :, :64].cpu(), x)
    assert_close(feat, ref, FP32_REL, "final_feature")
    # host pipeline: chunking must not change anything (philox keyed by global scene index)
    m.set_rng("philox", seed=11)
    f_dev, h_dev = m(x.to(DEV))
    for cs, mode in ((1000, "auto"), (256, "host"), (77, "device"), (300, "device")):
        m.set_rng("philox", seed=11)
        f_host, h_host = m.forward_host(x.pin_memory(), chunk_scenes=cs, input_slice=mode)
        # no synchronize here: forward_host returns when its outputs are complete on the host
        assert torch.equal(f_host, f_dev.cpu()) and torch.equal(h_host, h_dev.cpu()), (cs, mode)
    m.set_rng("philox", seed=11)
    f_host, h_host, done = m.forward_host(x.pin_memory(), chunk_scenes=300, sync=False)
    done.synchronize()
    assert torch.equal(f_host, f_dev.cpu()) and torch.equal(h_host, h_dev.cpu())
    assert m.launches_per_forward(b, n) == 1 + 5 + 6 + 6
    # reference RNG contract through the host pipeline: the whole batch's uniforms are drawn first, per layer in
    # the reference's order, so chunking does not change which draw a scene gets
    m.set_rng("cpu-compat")
    torch.manual_seed(99)
    f_one, _ = m(x.to(DEV))
    for cs in (1000, 256):
        torch.manual_seed(99)
        f_chunked, _ = m.forward_host(x.pin_memory(), chunk_scenes=cs)
        assert torch.equal(f_chunked, f_one.cpu()), cs


def test_pack_cache_invalidation_and_module_copies():
    """writes through .data are invisible to the (data_ptr, _version) key -> invalidate_packs();
    deepcopy after a forward must work and give an independent module."""
    import copy
    torch.manual_seed(4)
    m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1).to(DEV).eval()
    h = torch.randn(5, 6, 64, device=DEV)
    u = torch.rand(5, 36, 6, device=DEV)
    out1, _ = m(h, noise=[u])
    w = m.nmp_mlp_end.layers[1].weight
    w.data.mul_(2.0)                                    # does not bump w._version
    m.invalidate_packs()
    out2, _ = m(h, noise=[u])
    assert not torch.equal(out1, out2)
    c = copy.deepcopy(m)                                 # the original holds ctypes structs by now
    out3, _ = c(h, noise=[u])
    assert torch.equal(out2, out3)
    with torch.no_grad():
        c.nmp_mlp_end.layers[1].weight.mul_(0.5)         # autograd-visible in-place write: cache key changes
    out4, _ = c(h, noise=[u])
    assert torch.equal(out4, out1)                        # x2 then x0.5 is exact
    assert torch.equal(m(h, noise=[u])[0], out2)         # the copy's update did not touch the original
    sd = {k: v.clone() for k, v in c.state_dict().items()}
    m.load_state_dict(sd)                                # post-hook invalidates
    assert torch.equal(m(h, noise=[u])[0], out4)


# ---- bf16 tensor-core (tcgen05) path: 2e-2 (BASELINE.json north_star) ------------------
@pytest.mark.parametrize("name", golden_names())
def test_layer_forward_bf16_tc_vs_golden(name):
    g = load_golden(name)
    node, fac, hinc = _run_layer(g, precision="bf16")
    if hinc is not None:
        assert torch.equal(hinc.cpu(), torch.from_numpy(g["H"]))      # membership stays bit-exact
    assert_close(fac, g["factors"], BF16_REL, f"{name} factors (bf16)")
    assert_close(node, g["node_feat"], BF16_REL, f"{name} node_feat (bf16)")
    assert torch.allclose(fac.sum(-1), torch.ones_like(fac[..., 0]), atol=1e-5)


def test_bf16_tc_large_batch_vs_fp32_path():
    """Many tiles per CTA, ragged last tile: tcgen05 path vs the fp32 path on the same inputs."""
    torch.manual_seed(77)
    m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1).to(DEV)
    gen = torch.Generator().manual_seed(8)
    b = 3001                                   # 3001 * 121 rows: not a multiple of 128
    h = torch.randn(b, 11, 64, generator=gen).to(DEV)
    u = torch.rand(b, 121, 6, generator=gen).to(DEV)
    n32, f32 = m(h, noise=[u])
    m.set_precision("bf16")
    n16, f16 = m(h, noise=[u])
    assert_close(f16, f32, BF16_REL, "factors bf16 vs fp32")
    assert_close(n16, n32, BF16_REL, "node_feat bf16 vs fp32")
    n16b, f16b = m(h, noise=[u])
    assert torch.equal(n16, n16b) and torch.equal(f16, f16b)          # deterministic


@pytest.mark.parametrize("n,b", [(12, 50), (13, 301), (20, 77), (36, 9), (8, 1000), (37, 5)])
def test_bf16_pairwise_scene_aligned_tiles_vs_fp32_path(n, b):
    """Pairwise layer on the tensor-core path for N^2 >= 128: the fused node2edge + MLP chain uses scene-aligned
    tiles (ceil(N^2 / 128) per scene, ragged last tile); N = 37 exceeds the staged node block and takes the
    unfused chain; N = 8 keeps the linear tiles."""
    torch.manual_seed(78)
    m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1).to(DEV)
    gen = torch.Generator().manual_seed(10)
    h = torch.randn(b, n, 64, generator=gen).to(DEV)
    u = torch.rand(b, n * n, 6, generator=gen).to(DEV)
    n32, f32 = m(h, noise=[u])
    m.set_precision("bf16")
    n16, f16 = m(h, noise=[u])
    assert_close(f16, f32, BF16_REL, "factors bf16 vs fp32")
    assert_close(n16, n32, BF16_REL, "node_feat bf16 vs fp32")
    assert torch.allclose(f16.sum(-1), torch.ones_like(f16[..., 0]), atol=1e-5)
    m.set_rng("philox", seed=3)
    a1, _ = m(h)
    a2, _ = m(torch.cat((h[b // 2:], h[:b // 2])))        # Philox is keyed by the element index, not the tile
    assert a1.shape == a2.shape


@pytest.mark.parametrize("d,n,scale,bo,b", [(256, 64, 4, 256, 37), (256, 64, 16, 96, 301), (256, 20, 5, 256, 45),
                                            (256, 32, 8, 80, 19), (256, 7, 3, 128, 130),
                                            (64, 11, 5, 64, 3001), (64, 20, 8, 32, 77), (64, 8, 3, 64, 1000),
                                            (64, 64, 16, 64, 9), (64, 3, 2, 64, 4099)])
def test_bf16_wide_hyper_fused_vs_fp32_path(d, n, scale, bo, b):
    """h_dim 256 / 64 hyper layers run the fused gather + T MLPs + scatter (+ closing MLP) kernels
    (csrc/gn_hyper_fused_tc.cu, gn_hyper_fused64_tc.cu): many tiles per CTA, ragged last tile (b % (128 // n) != 0), scene
    counts per tile that do not fill 128 rows (n = 20, 7), closing MLP fused (bo % 32 == 0) or not (bo = 80)."""
    torch.manual_seed(91)
    m = gb.MS_HGNN_hyper(d, d, 64, bo, batch_norm=0, nmp_layers=1, scale=scale).to(DEV)
    gen = torch.Generator().manual_seed(9)
    h = torch.randn(b, n, d, generator=gen).to(DEV)
    corr = torch.bmm(torch.nn.functional.normalize(h, dim=2), torch.nn.functional.normalize(h, dim=2).transpose(1, 2))
    u = torch.rand(b, n, 10, generator=gen).to(DEV)
    n32, f32, h32 = m(h, corr, noise=[u])
    m.set_precision("bf16")
    n16, f16, h16 = m(h, corr, noise=[u])
    assert torch.equal(h32, h16)
    assert_close(f16, f32, BF16_REL, "factors bf16 vs fp32")
    assert_close(n16, n32, BF16_REL, "node_feat bf16 vs fp32")
    n16b, _, _ = m(h, corr, noise=[u])
    assert torch.equal(n16, n16b)                                      # deterministic
    # incidence read in place from a concatenated new_H (scene stride > E*N, model/GroupNet_nba.py:296-299)
    hcat = torch.zeros(b, 2 * n + 1, n, device=DEV)
    hcat[:, n + 1:] = h16
    n16c, _, _ = m(h, corr, noise=[u], H=hcat[:, n + 1:])
    assert torch.equal(n16c, n16)
    # strided output (the concatenated feature tensor of PastEncoder.forward, :301-309)
    wide = torch.zeros(b, n, bo + 64, device=DEV)
    m(h, corr, noise=[u], out=wide[:, :, 32:32 + bo], want_factors=False)
    assert torch.equal(wide[:, :, 32:32 + bo], n16) and float(wide[:, :, :32].abs().max()) == 0.0


# ---- T7: backward (gn_stage_bwd) vs torch autograd through the CPU oracle ---------------------
@pytest.mark.train
@pytest.mark.parametrize("kind,n,d,bo,scale,layers,b", [
    ("pairwise", 5, 64, 64, 0, 1, 7), ("hyper", 6, 64, 64, 3, 1, 9), ("hyper", 6, 64, 64, 6, 1, 9),
    ("pairwise", 11, 64, 64, 0, 1, 33), ("hyper", 11, 64, 64, 5, 1, 33),
    ("pairwise", 4, 32, 48, 0, 2, 5), ("hyper", 7, 128, 64, 2, 2, 6),
    ("pairwise", 1, 64, 64, 0, 1, 5), ("pairwise", 2, 64, 32, 0, 1, 300), ("pairwise", 20, 64, 64, 0, 1, 3),
    ("pairwise", 6, 128, 64, 0, 1, 17),
])
def test_backward_matches_oracle_autograd(kind, n, d, bo, scale, layers, b):
    torch.manual_seed(500 + n + d + layers)
    if kind == "pairwise":
        m = gb.MS_HGNN_oridinary(16, d, 64, bo, batch_norm=0, nmp_layers=layers)
        e, t = n * n, 6
    else:
        m = gb.MS_HGNN_hyper(d, d, 64, bo, batch_norm=0, nmp_layers=layers, scale=scale)
        e, t = (1 if scale == n else n), 10
    gen = torch.Generator().manual_seed(b)
    h = torch.randn(b, n, d, generator=gen)
    noise = [torch.rand(b, e, t, generator=gen) for _ in range(layers)]
    w_node = torch.randn(b, n, bo, generator=gen)
    w_fac = torch.randn(b, e, t, generator=gen)
    # reference gradients: torch autograd through the CPU oracle on the same weights
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    h_ref = h.clone().requires_grad_(True)
    if kind == "pairwise":
        node_r, fac_r = O.forward_pairwise(sd, h_ref, noise, nmp_layers=layers)
    else:
        corr = O.feature_correlation(h)
        node_r, fac_r, _ = O.forward_hyper(sd, h_ref, corr, scale, noise, nmp_layers=layers)
    ((node_r * w_node).sum() + (fac_r * w_fac).sum()).backward()
    # ours
    m = m.to(DEV).train()
    h_dev = h.to(DEV).requires_grad_(True)
    if kind == "pairwise":
        node, fac = m(h_dev, noise=[u.to(DEV) for u in noise])
    else:
        node, fac, hm = m(h_dev, corr.to(DEV), noise=[u.to(DEV) for u in noise])
        assert not hm.requires_grad
    assert node.requires_grad and fac.requires_grad
    assert_close(node, node_r.detach(), FP32_REL, "train-mode node_feat")
    ((node * w_node.to(DEV)).sum() + (fac * w_fac.to(DEV)).sum()).backward()
    assert_close(h_dev.grad, h_ref.grad, 1e-4, "d h_states", atol=1e-7)
    used, unused = 0, 0
    for name, p in m.named_parameters():
        ref_g = sd[name].grad
        if ref_g is None:
            assert p.grad is None, f"{name}: never-used parameter must keep grad None"
            unused += 1
        else:
            assert p.grad is not None, f"{name}: missing gradient"
            assert_close(p.grad, ref_g, 1e-4, f"grad {name}", atol=1e-7)
            used += 1
    assert used > 20 and unused >= 4          # edge_aggregation.mlp (+ spatial_* for the hyper layer)


@pytest.mark.train
def test_backward_accumulates_and_optimizer_step_repacks():
    torch.manual_seed(3)
    m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1).to(DEV)
    opt = torch.optim.SGD([p for p in m.parameters()], lr=0.1)
    h = torch.randn(4, 5, 64, device=DEV)
    u = torch.rand(4, 25, 6, device=DEV)
    out1, _ = m(h, noise=[u])
    out1.sum().backward()
    g1 = m.nmp_mlp_end.layers[1].bias.grad.clone()
    out1b, _ = m(h, noise=[u])
    out1b.sum().backward()                                            # grads accumulate like torch
    assert torch.allclose(m.nmp_mlp_end.layers[1].bias.grad, 2 * g1)
    opt.step()
    with torch.no_grad():
        out2, _ = m(h, noise=[u])
    assert not torch.equal(out1.detach(), out2)                       # packed weights were refreshed


@pytest.mark.train
def test_flat_grad_arena_takes_the_gradients_in_place():
    """ddp.FlatGradBucket.zero_grad(): .grad are views into one flat buffer, the backward accumulates into them and
    the all-reduce runs on the buffer itself; values equal the plain backward (wgrad uses float atomics: 1e-5)."""
    from groupnet_b200.ddp import FlatGradBucket
    x = torch.randn(6, 11, 64, device=DEV)
    wgt = torch.randn(6, 11, 256, device=DEV)
    noise = [torch.rand(6, 121, 6, device=DEV), torch.rand(6, 11, 10, device=DEV), torch.rand(6, 1, 10, device=DEV)]

    def run(arena):
        torch.manual_seed(77)
        m = gb.MultiScaleInteraction(64, (5, 11)).to(DEV).train()
        bucket = FlatGradBucket(m.parameters())
        for _ in range(2):                                            # the second step reuses the installed views
            if arena:
                bucket.zero_grad()
            else:
                for p in m.parameters():
                    p.grad = None
            feat, _ = m(x, noise=noise)
            (feat * wgt).sum().backward()
            bucket.allreduce_mean()
        return m, bucket

    plain, _ = run(False)
    arena, bucket = run(True)
    assert bucket._views_installed()
    off = 0
    for (name, p), (_, q) in zip(plain.named_parameters(), arena.named_parameters()):
        assert q.grad is not None and q.grad.data_ptr() == bucket.flat.data_ptr() + 4 * off
        off += q.numel()
        if p.grad is None:
            assert q.grad.abs().max().item() == 0.0, name              # never-used parameter: zeros in the arena
        else:
            assert_close(q.grad, p.grad, 1e-5, f"arena grad {name}", atol=1e-7)


# ---- T6 / §8(f) rank 1: the whole PastEncoder against the reference's own PastEncoder --------
@pytest.mark.parametrize("name", golden_names("pastenc"))
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_REL), ("tf32", FP32_REL), ("bf16", BF16_REL)])
def test_past_encoder_vs_reference_golden(name, precision, tol):
    g = load_golden(name)
    enc = build_past_encoder(g).to(DEV)
    for l in enc.layers():
        l.set_precision(precision)
    inputs = torch.from_numpy(g["inputs"]).to(DEV)
    feat, new_h = enc(inputs, g["B"], g["N"], noise=[torch.from_numpy(u).to(DEV) for u in g["noise"]])
    assert feat.shape == g["output_feature"].shape and new_h.shape == g["new_H"].shape
    assert torch.equal(new_h.cpu(), torch.from_numpy(g["new_H"]))            # hyperedge membership bit-exact
    ref = torch.from_numpy(g["output_feature"])
    assert_close(feat[:, :64], ref[:, :64], 1e-5, "ftraj_input (folded front-end)")
    for i, part in enumerate(("pairwise", "hyper5", "hyper11")):
        assert_close(feat[:, 64 * (i + 1):64 * (i + 2)], ref[:, 64 * (i + 1):64 * (i + 2)], tol, part)


def test_past_encoder_errors_like_reference():
    import types
    enc = gb.PastEncoder(types.SimpleNamespace(hidden_dim=64, hyper_scales=[3], past_length=5)).to(DEV).eval()
    with pytest.raises(IndexError):                       # add_category indexes agent 10 (:264)
        enc(torch.randn(2 * 8, 5, 4, device=DEV), 2, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.randn(2 * 11, 5, 4), 2, 11)


# ---- graph-replayed rollout forward ---------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_past_encoder_matches_eager_and_consumes_the_same_rng_stream(precision):
    import types
    args = types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], past_length=5)
    torch.manual_seed(31)
    enc = gb.PastEncoder(args).to(DEV).eval()
    enc._interaction_block().set_precision(precision)
    for b in (1, 3):
        x = torch.randn(b * 11, 5, 4)
        g = gb.GraphedPastEncoder(enc, b, 11, 5)
        for trial in range(3):
            xt = x + trial
            torch.manual_seed(100 + trial)
            f0, h0 = enc(xt.to(DEV), b, 11)
            after0 = torch.rand(1)
            torch.manual_seed(100 + trial)
            f1, h1 = g(xt)
            after1 = torch.rand(1)
            assert torch.equal(f0, f1) and torch.equal(h0, h1)      # fresh noise per replay, same kernels
            assert torch.equal(after0, after1)                       # the same draws left the CPU generator
        big = torch.randn(64 * 11, 5, 4, device=DEV)                 # eager call that grows the layers' workspaces
        enc(big, 64, 11)
        torch.manual_seed(7)
        f0, _ = enc(x.to(DEV), b, 11)
        torch.manual_seed(7)
        f1, _ = g(x, clone=True)
        assert torch.equal(f0, f1)
    with pytest.raises(RuntimeError):
        g(torch.zeros(5, 5, 4))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_philox_device_seed_mode_equals_philox_and_replays_from_a_graph(precision):
    import types
    # layer level, eager: the device-resident seed walks the same sequence as the by-value seed (chunked batch too)
    for kind, scale in (("pairwise", 0), ("hyper", 5), ("hyper", 11)):
        torch.manual_seed(3)
        if kind == "pairwise":
            m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=2)
        else:
            m = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=2, scale=scale)
        m = m.to(DEV).set_precision(precision)
        m.workspace_limit_bytes = 8 << 20
        h = torch.randn(150, 11, 64, device=DEV)
        q = torch.nn.functional.normalize(h, dim=2)
        args = (h,) if kind == "pairwise" else (h, torch.bmm(q, q.transpose(1, 2)))
        with torch.no_grad():
            m.set_rng("philox", seed=(1 << 63) + 12345)
            ref = [tuple(t.clone() for t in m(*args)[:2]) for _ in range(3)]
            m.set_rng("philox-device", seed=(1 << 63) + 12345)
            got = [tuple(t.clone() for t in m(*args)[:2]) for _ in range(3)]
        for r, g in zip(ref, got):
            assert torch.equal(r[0], g[0]) and torch.equal(r[1], g[1])
        assert not torch.equal(ref[0][1], ref[1][1])                 # the noise does change from call to call
    # encoder level, graph replay k == eager philox call k
    args = types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], past_length=5)
    torch.manual_seed(31)
    enc = gb.PastEncoder(args).to(DEV).eval()
    block = enc._interaction_block()
    block.set_precision(precision)
    x = torch.randn(2 * 11, 5, 4, device=DEV)
    block.set_rng("philox", seed=77)
    ref = [tuple(t.clone() for t in enc(x, 2, 11)) for _ in range(3)]
    g = gb.GraphedPastEncoder(enc, 2, 11, 5, rng="philox", seed=77)
    for k in range(3):
        f, hh = g(x, clone=True)
        assert torch.equal(f, ref[k][0]) and torch.equal(hh, ref[k][1])
    layer_args = (h, torch.bmm(q, q.transpose(1, 2)))
    with torch.enable_grad(), pytest.raises(RuntimeError, match="philox-device"):
        m.set_rng("philox-device", 1)
        m(*layer_args)                                               # parameters require grad -> training path
