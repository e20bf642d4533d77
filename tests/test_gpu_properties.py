"""T9: randomized property tests (hypothesis) of the CUDA path against the CPU oracle over the
shape space the library accepts: N in [1, 64], D in {32..256}, scale in [0, N], ragged batch sizes."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import groupnet_b200 as gb
from helpers import FP32_REL, BF16_REL, assert_close
from oracle import ms_hgnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
COMMON = dict(deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


@settings(max_examples=40, **COMMON)
@given(n=st.integers(1, 64), d4=st.integers(1, 64), b=st.integers(1, 300), seed=st.integers(0, 2**20),
       data=st.data())
def test_topk_and_fused_corr_random_shapes(n, d4, b, seed, data):
    d = 4 * d4
    scales = sorted(set(data.draw(st.lists(st.integers(0, n), min_size=1, max_size=4))))
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(b, n, d, generator=gen)
    with torch.no_grad():
        (hs, full), corr = gb.corr_topk_h(x.to(DEV), scales, concat=True, return_corr=True)
    corr = corr.cpu()
    ref_corr = O.feature_correlation(x)
    assert (corr - ref_corr).abs().max().item() <= 2e-6
    for s, hm in zip(scales, hs):
        hm = hm.cpu()
        assert np.array_equal(hm.numpy(), O.incidence_topk_lowest_index(corr.numpy(), s))
        assert torch.equal(gb.topk_h(corr.to(DEV), s).cpu(), hm)
        if hm.shape[1] == n:
            ref = O.incidence_topk(ref_corr, s)
            bad = ~(hm == ref).all(dim=2)
            assert (O.topk_gap(ref_corr, s)[bad] < 4e-6).all()


@settings(max_examples=14, **COMMON)
@given(pairwise=st.booleans(), n=st.integers(1, 24), d=st.sampled_from([32, 64, 128, 256]),
       bo=st.sampled_from([32, 48, 64, 96, 128]), layers=st.integers(1, 2), b=st.integers(1, 40),
       seed=st.integers(0, 2**20), data=st.data())
def test_layer_forward_random_shapes(pairwise, n, d, bo, layers, b, seed, data):
    if pairwise and n > 12:
        n = 12                                       # keep the oracle's (B,E,N,128) tensor small
    scale = 0 if pairwise else data.draw(st.integers(0, n))
    torch.manual_seed(seed)
    if pairwise:
        m = gb.MS_HGNN_oridinary(16, d, 64, bo, batch_norm=0, nmp_layers=layers)
        e, t = n * n, 6
    else:
        m = gb.MS_HGNN_hyper(d, d, 64, bo, batch_norm=0, nmp_layers=layers, scale=scale)
        e, t = (1 if scale == n else n), 10
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(seed + 1)
    h = torch.randn(b, n, d, generator=gen)
    noise = [torch.rand(b, e, t, generator=gen) for _ in range(layers)]
    m = m.to(DEV)
    with torch.no_grad():
        if pairwise:
            ref_node, ref_fac = O.forward_pairwise(sd, h, noise, nmp_layers=layers)
            outs = {p: m.set_precision(p)(h.to(DEV), noise=noise) for p in ("fp32", "bf16")}
        else:
            corr = O.feature_correlation(h)
            ref_node, ref_fac, ref_h = O.forward_hyper(sd, h, corr, scale, noise, nmp_layers=layers)
            outs = {p: m.set_precision(p)(h.to(DEV), corr.to(DEV), noise=noise) for p in ("fp32", "bf16")}
            assert torch.equal(outs["fp32"][2].cpu(), ref_h)
    for prec, tol in (("fp32", FP32_REL), ("bf16", BF16_REL)):
        assert_close(outs[prec][1], ref_fac, tol, f"{prec} factors")
        assert_close(outs[prec][0], ref_node, tol, f"{prec} node_feat")
