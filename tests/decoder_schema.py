"""Parameter containers with the reference decoder's schema (model/GroupNet_nba.py:13-46, :441-456).

Test infrastructure: `torch.manual_seed(s)` followed by `DecoderSchema(...)` regenerates, in the reference's
construction and initialisation order, the weights the decoder fixtures were computed with (pinned by sha256),
so the fixtures do not have to carry 1.4 M parameters.  No forward: the math lives in oracle/decoder_oracle.py.
"""
import torch.nn as nn

import groupnet_b200 as gb


class DecomposeBlockSchema(nn.Module):
    def __init__(self, past_len, future_len, input_dim):
        super().__init__()
        self.conv_past = nn.Conv1d(2, 32, 3, stride=1, padding=1)
        self.encoder_past = nn.GRU(32, 96, 1, batch_first=True)
        self.decoder_y = gb.MLP(96 + input_dim, future_len * 2, hidden_size=(512, 256))
        self.decoder_x = gb.MLP(96 + input_dim, past_len * 2, hidden_size=(512, 256))
        nn.init.kaiming_normal_(self.conv_past.weight)                     # init_parameters (:38-46)
        nn.init.kaiming_normal_(self.encoder_past.weight_ih_l0)
        nn.init.kaiming_normal_(self.encoder_past.weight_hh_l0)
        nn.init.zeros_(self.conv_past.bias)
        nn.init.zeros_(self.encoder_past.bias_ih_l0)
        nn.init.zeros_(self.encoder_past.bias_hh_l0)


class DecoderSchema(nn.Module):
    def __init__(self, hidden_dim, hyper_scales, zdim, past_length, future_length, num_decompose):
        super().__init__()
        input_dim = (2 + len(hyper_scales)) * hidden_dim + zdim            # :447-450
        self.decompose = nn.ModuleList(
            DecomposeBlockSchema(past_length, future_length, input_dim) for _ in range(num_decompose))
