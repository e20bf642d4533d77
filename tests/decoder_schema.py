"""`torch.manual_seed(s)` followed by `DecoderSchema(...)` regenerates, in the reference's construction and
initialisation order (model/GroupNet_nba.py:13-46, :441-456), the weights the decoder fixtures were computed
with (pinned by sha256), so the fixtures do not have to carry 1.4 M parameters.  The container is the drop-in
`groupnet_b200.Decoder` itself."""
import types

import groupnet_b200 as gb


def DecoderSchema(hidden_dim, hyper_scales, zdim, past_length, future_length, num_decompose):
    return gb.Decoder(types.SimpleNamespace(hidden_dim=hidden_dim, hyper_scales=list(hyper_scales), zdim=zdim,
                                            past_length=past_length, future_length=future_length,
                                            num_decompose=num_decompose))
