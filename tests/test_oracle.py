"""The CPU oracle against the golden vectors generated from the live reference
(tests/golden/make_golden.py), and the C restatement of the index work."""
import numpy as np
import pytest
import torch

from helpers import build_layer, golden_names, load_golden, rel_err
from oracle import c_oracle
from oracle import ms_hgnn_oracle as O


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    g = load_golden(name)
    sd = build_layer(g).state_dict()
    h = torch.from_numpy(g["h"])
    noise = [torch.from_numpy(u) for u in g["noise"]]
    with torch.no_grad():
        if g["kind"] == "pairwise":
            node, fac = O.forward_pairwise(sd, h, noise, nmp_layers=g["L"])
        else:
            node, fac, hinc = O.forward_hyper(sd, h, torch.from_numpy(g["corr"]), g["scale"], noise,
                                              nmp_layers=g["L"])
            assert np.array_equal(hinc.numpy(), g["H"]), "incidence must be bit-exact"
    assert rel_err(node, g["node_feat"]) <= 2e-6
    assert rel_err(fac, g["factors"]) <= 2e-6
    assert torch.allclose(fac.sum(-1), torch.ones(fac.shape[:-1]), atol=1e-5)


@pytest.mark.parametrize("name", [n for n in golden_names() if "hyper" in n])
def test_c_oracle_topk_matches_golden(name):
    if not c_oracle.available():
        pytest.skip("oracle/libgn_oracle.so not built (run __graft_entry__.build())")
    g = load_golden(name)
    got = c_oracle.topk_h(g["corr"], g["scale"])
    assert np.array_equal(got, g["H"])
    assert np.array_equal(O.incidence_topk_lowest_index(g["corr"], g["scale"]), g["H"])


def test_c_oracle_corr_close():
    if not c_oracle.available():
        pytest.skip("oracle/libgn_oracle.so not built")
    g = load_golden("nba_hyper5")
    got = c_oracle.corr(g["h"])
    assert np.abs(got - g["corr"]).max() <= 2e-6


def test_tie_rule_lowest_index():
    # constructed ties at the k boundary: equal values, lower index must win
    corr = np.array([[[1.0, 0.5, 0.5, 0.5],
                      [0.2, 1.0, 0.2, 0.2],
                      [0.0, 0.0, 0.0, 0.0],
                      [0.3, 0.9, 0.9, 1.0]]], dtype=np.float32)
    h = O.incidence_topk_lowest_index(corr, 2)
    assert h[0].tolist() == [[1, 1, 0, 0], [1, 1, 0, 0], [1, 1, 0, 0], [0, 1, 0, 1]]
    if c_oracle.available():
        assert np.array_equal(c_oracle.topk_h(corr, 2), h)


def test_scale_rules():
    corr = O.feature_correlation(torch.randn(2, 5, 8, generator=torch.Generator().manual_seed(0)))
    assert O.incidence_topk(corr, 5).shape == (2, 1, 5)          # scale == N: one all-ones edge
    assert torch.equal(O.incidence_topk(corr, 0), O.incidence_topk(corr, 1))  # scale < 1 -> k = 1
    with pytest.raises(RuntimeError):
        O.incidence_topk(corr, 6)
    with pytest.raises(RuntimeError):
        O.incidence_topk_lowest_index(corr.numpy(), 6)


def test_pairwise_incidence_self_loops():
    h = O.pairwise_incidence(3, 1)[0]
    assert h.shape == (9, 3)
    assert h[0].tolist() == [2, 0, 0] and h[1].tolist() == [1, 1, 0] and h[5].tolist() == [0, 1, 1]
    assert float(h.sum()) == 18.0
