"""Live differential test: oracle restatement vs the UNMODIFIED reference module.
Runs only where the reference tree exists (the build container)."""
import sys

import pytest
import torch

from helpers import REFERENCE_DIR, have_reference, rel_err
from oracle import ms_hgnn_oracle as O

pytestmark = pytest.mark.skipif(not have_reference(), reason="reference tree not present")


def _ref():
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import model.MS_HGNN_batch as ref
    return ref


def _with_noise(ref, fn, noise):
    it = iter(noise)
    orig = ref.sample_gumbel
    ref.sample_gumbel = lambda shape, eps=1e-10: -torch.log(eps - torch.log(next(it) + eps))
    try:
        with torch.no_grad():
            return fn()
    finally:
        ref.sample_gumbel = orig


@pytest.mark.parametrize("n,d,bo,layers", [(11, 64, 64, 1), (8, 64, 32, 2), (5, 128, 64, 3)])
def test_pairwise_live(n, d, bo, layers):
    ref = _ref()
    torch.manual_seed(100 + n)
    mod = ref.MS_HGNN_oridinary(16, d, 64, bo, batch_norm=0, nmp_layers=layers).eval()
    h = torch.randn(4, n, d)
    noise = [torch.rand(4, n * n, 6) for _ in range(layers)]
    a, b = _with_noise(ref, lambda: mod(h), noise)
    with torch.no_grad():
        a2, b2 = O.forward_pairwise(mod.state_dict(), h, noise, nmp_layers=layers)
    assert rel_err(a2, a) <= 2e-6 and rel_err(b2, b) <= 2e-6


@pytest.mark.parametrize("n,scale,layers", [(11, 5, 1), (11, 11, 1), (8, 3, 2), (20, 8, 1), (6, 0, 1)])
def test_hyper_live(n, scale, layers):
    ref = _ref()
    torch.manual_seed(200 + n + scale)
    mod = ref.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=layers, scale=scale).eval()
    h = torch.randn(4, n, 64)
    corr = O.feature_correlation(h)
    e = 1 if scale == n else n
    noise = [torch.rand(4, e, 10) for _ in range(layers)]
    a, b, hm = _with_noise(ref, lambda: mod(h, corr), noise)
    with torch.no_grad():
        a2, b2, hm2 = O.forward_hyper(mod.state_dict(), h, corr, scale, noise, nmp_layers=layers)
    assert torch.equal(hm, hm2)
    assert rel_err(a2, a) <= 2e-6 and rel_err(b2, b) <= 2e-6


def test_seeded_rng_order_matches_reference():
    """cpu-compat contract: one torch.rand(B,E,T) per MLP_dict_softmax call, in call order."""
    ref = _ref()
    torch.manual_seed(7)
    mod = ref.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=2).eval()
    h = torch.randn(3, 4, 64)
    torch.manual_seed(99)
    with torch.no_grad():
        a, b = mod(h)
    torch.manual_seed(99)
    noise = [torch.rand(s) for s in O.noise_shapes(3, 4, True, 0, 2)]
    with torch.no_grad():
        a2, b2 = O.forward_pairwise(mod.state_dict(), h, noise, nmp_layers=2)
    assert rel_err(a2, a) <= 2e-6 and rel_err(b2, b) <= 2e-6
