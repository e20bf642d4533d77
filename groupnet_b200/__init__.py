"""groupnet_b200 — B200-native (sm_100a) multiscale hypergraph message-passing layers.

Drop-in for `model/MS_HGNN_batch.py` of TaliMotzkin/GroupNet: same nn.Module API,
hand-written CUDA kernels behind a C ABI (include/groupnet_b200.h).
"""
from .layers import (MLP, MLP_dict, MLP_dict_softmax, MS_HGNN_hyper, MS_HGNN_oridinary,
                     edge_aggregation, encode_onehot, gumbel_softmax, gumbel_softmax_sample, make_mlp, my_softmax,
                     sample_gumbel)
from .encoder import PastEncoder, PositionalAgentEncoding
from .interaction import MultiScaleInteraction
from .decoder import Decoder, DecomposeBlock
from .rollout import GraphedInference, GraphedPastEncoder, inference_simulator
from .fish import (HyperEdgeAttention, MLPHGE, TemporalGATLayer, build_dynamic_graph_and_hypergraph,
                   compute_alpha_im)
from .ops import corr_topk_h, topk_h
from ._lib import GroupNetLibraryError, LIB_PATH

__all__ = [
    "MS_HGNN_oridinary", "MS_HGNN_hyper", "MLP", "MLP_dict", "MLP_dict_softmax",
    "edge_aggregation", "encode_onehot", "make_mlp", "sample_gumbel", "gumbel_softmax_sample", "gumbel_softmax",
    "my_softmax", "MultiScaleInteraction", "PastEncoder", "PositionalAgentEncoding", "GraphedPastEncoder", "GraphedInference", "inference_simulator", "Decoder", "DecomposeBlock", "MLPHGE", "HyperEdgeAttention", "TemporalGATLayer", "compute_alpha_im",
    "build_dynamic_graph_and_hypergraph", "corr_topk_h", "topk_h", "GroupNetLibraryError", "LIB_PATH",
]
__version__ = "0.1.0"
