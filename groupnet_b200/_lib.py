"""ctypes binding of libgroupnet_b200.so (the C ABI declared in include/groupnet_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or was
built for another ABI version every op raises `GroupNetLibraryError`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgroupnet_b200.so")
ABI_VERSION = 6

GN_MAX_AGENTS = 64
GN_MAX_SCALES = 8
GN_FP32 = 0
GN_BF16_TC = 1
GN_TF32X3 = 2
GN_NOISE_GIVEN = 0
GN_NOISE_PHILOX = 1
GN_NOISE_PHILOX_DEVICE_SEED = 2

# every symbol include/groupnet_b200.h declares
EXPORTS = (
    "gn_abi_version",
    "gn_error_string",
    "gn_corr_topk_h",
    "gn_topk_h",
    "gn_stage_workspace_bytes",
    "gn_stage_fwd",
    "gn_stage_launch_count",
    "gn_past_frontend",
    "gn_decoder_workspace_bytes",
    "gn_decoder_fwd",
    "gn_decoder_tc_workspace_bytes",
    "gn_decoder_fwd_tc",
    "gn_stage_saved_offsets",
    "gn_stage_bwd_workspace_bytes",
    "gn_stage_bwd",
    "gn_profile_enable",
    "gn_profile_collect",
    "gn_profile_set_trace",
    "gn_fish_alpha_im",
    "gn_fish_bmm_t",
    "gn_fish_mlp",
    "gn_fish_hga_core",
    "gn_fish_gat_edges",
    "gn_fish_dynamic_graph",
)


class GroupNetLibraryError(RuntimeError):
    pass


class StageWeights(C.Structure):
    """struct gn_stage_weights (43 device pointers, header order)."""
    FIELDS = (
        "node_w0t", "node_b0", "node_w1t", "node_b1",
        "att_wpqt", "att_b0", "att_w1", "att_b1",
        "init_w0t", "init_b0", "init_w1t", "init_b1",
        "df_w0t", "df_b0", "df_w1", "df_b1",
        "agg_w0t", "agg_b0", "agg_w1t", "agg_b1",
        "post_w0t", "post_b0", "post_w1t", "post_b1",
        "tc_init_w0", "tc_init_w1", "tc_df_w0", "tc_df_w1",
        "tc_node_w0", "tc_node_w1", "tc_att_wpq", "tc_agg_w0", "tc_agg_w1", "tc_post_w0", "tc_post_w1", "tc_hfuse_w", "tc_npre_w",
        "tf_chain_w", "tf_pre_w", "tf_aggin_w", "tf_aggout_w", "tf_hagg_w", "tf_post_w", "tf_pagg_w",
    )
    _fields_ = [(name, C.c_void_p) for name in FIELDS]


class DecoderWeights(C.Structure):
    """struct gn_decoder_weights (17 device pointers, header order)."""
    FIELDS = ("conv_w", "conv_b", "gru_wx", "gru_wh", "gru_b",
              "x_w0", "x_b0", "x_w1", "x_b1", "x_w2", "x_b2",
              "y_w0", "y_b0", "y_w1", "y_b1", "y_w2", "y_b2")
    _fields_ = [(name, C.c_void_p) for name in FIELDS]


class DecoderTcWeights(C.Structure):
    """struct gn_decoder_tc_weights (16 device pointers, header order)."""
    FIELDS = ("conv_w", "conv_b", "gru_w", "gru_b", "w0", "b0",
              "x_w1", "x_b1", "x_w2", "x_b2", "y_w1", "y_b1", "y_w2", "y_b2", "mlp_stream", "mlp_bias")
    _fields_ = [(name, C.c_void_p) for name in FIELDS]


class StageCfg(C.Structure):
    """struct gn_stage_cfg."""
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("D", C.c_int32), ("Dout", C.c_int32),
        ("E", C.c_int32), ("T", C.c_int32), ("pairwise", C.c_int32), ("precision", C.c_int32),
        ("noise_mode", C.c_int32), ("stage_index", C.c_int32),
        ("out_ld", C.c_int32), ("h_stride", C.c_int32),
        ("seed", C.c_uint64), ("scene_offset", C.c_int64),
    ]


class Lin(C.Structure):
    """struct gn_lin."""
    _fields_ = [("W", C.c_void_p), ("b", C.c_void_p), ("dW", C.c_void_p), ("db", C.c_void_p),
                ("N", C.c_int32), ("K", C.c_int32)]


class TrainParams(C.Structure):
    """struct gn_train_params."""
    _fields_ = [("node0", Lin), ("node1", Lin), ("attpq", Lin),
                ("att_b0", C.c_void_p), ("att_w1", C.c_void_p), ("att_b1", C.c_void_p),
                ("d_att_b0", C.c_void_p), ("d_att_w1", C.c_void_p), ("d_att_b1", C.c_void_p),
                ("init0", Lin), ("init1", Lin), ("dist0", Lin), ("dist1", Lin), ("fac0", Lin), ("fac1", Lin),
                ("agg0", Lin * 15), ("agg1", Lin * 15), ("post0", Lin), ("post1", Lin)]


_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the library once; raise loudly if it is absent or mismatched."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GroupNetLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C groupnet_b200/csrc`. groupnet_b200 has no CPU / PyTorch fallback.")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as exc:  # missing libcudart etc.
            raise GroupNetLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
        for name in EXPORTS:
            if not hasattr(lib, name):
                raise GroupNetLibraryError(f"{LIB_PATH} does not export {name}")
        lib.gn_abi_version.restype = C.c_int
        lib.gn_error_string.restype = C.c_char_p
        lib.gn_error_string.argtypes = [C.c_int]
        lib.gn_corr_topk_h.restype = C.c_int
        lib.gn_corr_topk_h.argtypes = [
            C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32,
            C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]
        lib.gn_topk_h.restype = C.c_int
        lib.gn_topk_h.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_int64, C.c_void_p]
        lib.gn_stage_workspace_bytes.restype = C.c_size_t
        lib.gn_stage_workspace_bytes.argtypes = [C.POINTER(StageCfg)]
        lib.gn_stage_launch_count.restype = C.c_int
        lib.gn_stage_launch_count.argtypes = [C.POINTER(StageCfg)]
        lib.gn_stage_fwd.restype = C.c_int
        lib.gn_stage_fwd.argtypes = [
            C.POINTER(StageCfg), C.POINTER(StageWeights), C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.gn_past_frontend.restype = C.c_int
        lib.gn_past_frontend.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        lib.gn_decoder_workspace_bytes.restype = C.c_size_t
        lib.gn_decoder_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
        lib.gn_decoder_fwd.restype = C.c_int
        lib.gn_decoder_fwd.argtypes = [C.POINTER(DecoderWeights), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.gn_decoder_tc_workspace_bytes.restype = C.c_size_t
        lib.gn_decoder_tc_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        lib.gn_decoder_fwd_tc.restype = C.c_int
        lib.gn_decoder_fwd_tc.argtypes = [C.POINTER(DecoderTcWeights), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.gn_stage_saved_offsets.restype = C.c_int
        lib.gn_stage_saved_offsets.argtypes = [C.POINTER(StageCfg), C.POINTER(C.c_size_t)]
        lib.gn_stage_bwd_workspace_bytes.restype = C.c_size_t
        lib.gn_stage_bwd_workspace_bytes.argtypes = [C.POINTER(StageCfg)]
        lib.gn_stage_bwd.restype = C.c_int
        lib.gn_stage_bwd.argtypes = [C.POINTER(StageCfg), C.POINTER(TrainParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.c_void_p]
        lib.gn_profile_set_trace.restype = None
        lib.gn_profile_set_trace.argtypes = [C.c_void_p]
        lib.gn_profile_enable.restype = None
        lib.gn_profile_enable.argtypes = [C.c_int]
        lib.gn_profile_collect.restype = C.c_int
        lib.gn_profile_collect.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]
        vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        for name, args in (
            ("gn_fish_alpha_im", [vp, vp, vp, vp, i64, i32, i32, i32, i32, vp, vp]),
            ("gn_fish_bmm_t", [vp, i64, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
            ("gn_fish_mlp", [vp, i64, i64, i32, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(i32), C.POINTER(i32),
                             C.POINTER(f32), vp, i64, vp]),
            ("gn_fish_hga_core", [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp]),
            ("gn_fish_gat_edges", [vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp]),
            ("gn_fish_dynamic_graph", [vp, vp, vp, vp, i64, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
        ):
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = args
        got = lib.gn_abi_version()
        if got != ABI_VERSION:
            raise GroupNetLibraryError(f"{LIB_PATH} has ABI version {got}, expected {ABI_VERSION}")
        _lib = lib
        return _lib


def check(code: int, what: str) -> None:
    """Map a C return code to the exception the reference would raise."""
    if code == 0:
        return
    msg = load().gn_error_string(code).decode()
    if code == -3:
        # torch.topk in the reference (MS_HGNN_batch.py:382) raises RuntimeError with this text
        raise RuntimeError("selected index k out of range")
    if code < 0:
        raise ValueError(f"{what}: {msg} (gn_error {code})")
    raise RuntimeError(f"{what}: CUDA error {code}: {msg}")


def profile_enable(on: bool) -> None:
    load().gn_profile_enable(1 if on else 0)


def profile_collect() -> dict:
    """{kernel name: (total_ms, launches)} since the last collect (synchronises)."""
    lib = load()
    names = C.create_string_buffer(4096)
    ms = (C.c_float * 64)()
    cnt = (C.c_int * 64)()
    n = lib.gn_profile_collect(names, 4096, ms, cnt, 64)
    keys = [k for k in names.value.decode().split(";") if k]
    return {keys[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}
