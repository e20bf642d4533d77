"""Host-side boundary checks that need no GPU: the C-ABI library loads and
exports every symbol include/groupnet_b200.h declares, the drop-in modules keep
the reference's state_dict schema, packing is a pure permutation, and the
product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import sys

import pytest
import torch

import groupnet_b200 as gb
from groupnet_b200 import _lib, packing
from helpers import REFERENCE_DIR, ROOT, build_layer, golden_names, have_reference, load_golden


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "groupnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _declared_symbols()
    assert set(syms) == set(_lib.EXPORTS), (syms, _lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"libgroupnet_b200.so does not export {s}"
    assert _lib.load().gn_abi_version() == _lib.ABI_VERSION
    assert _lib.load().gn_error_string(-3) == b"selected index k out of range"


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.StageWeights) == 44 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.StageCfg) == 12 * 4 + 8 + 8
    text = open(os.path.join(ROOT, "include", "groupnet_b200.h")).read()
    body = text[text.index("typedef struct gn_stage_weights"):text.index("} gn_stage_weights;")]
    fields = re.findall(r"const (?:float|void)\*\s+(\w+);", body)
    assert tuple(fields) == _lib.StageWeights.FIELDS


@pytest.mark.skipif(not have_reference(), reason="reference tree not present")
@pytest.mark.parametrize("layers", [1, 2, 3])
def test_state_dict_matches_reference(layers):
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import model.MS_HGNN_batch as ref
    for ctor_r, ctor_m, kw in (
            (ref.MS_HGNN_oridinary, gb.MS_HGNN_oridinary, dict(embedding_dim=16)),
            (ref.MS_HGNN_hyper, gb.MS_HGNN_hyper, dict(embedding_dim=64, scale=5))):
        torch.manual_seed(5)
        r = ctor_r(h_dim=64, mlp_dim=64, bottleneck_dim=48, batch_norm=0, nmp_layers=layers, **kw)
        torch.manual_seed(5)
        m = ctor_m(h_dim=64, mlp_dim=64, bottleneck_dim=48, batch_norm=0, nmp_layers=layers, **kw)
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys())
        assert [n for n, _ in r.named_parameters()] == [n for n, _ in m.named_parameters()]
        for k in rs:
            assert rs[k].shape == ms[k].shape and torch.equal(rs[k], ms[k]), k
        m.load_state_dict(rs, strict=True)
        r.load_state_dict(ms, strict=True)


def test_default_constructor_signature():
    import inspect
    sig = inspect.signature(gb.MS_HGNN_oridinary.__init__)
    assert list(sig.parameters)[1:] == ["embedding_dim", "h_dim", "mlp_dim", "bottleneck_dim", "activation",
                                        "batch_norm", "dropout", "nmp_layers", "vis"]
    sig = inspect.signature(gb.MS_HGNN_hyper.__init__)
    assert list(sig.parameters)[1:] == ["embedding_dim", "h_dim", "mlp_dim", "bottleneck_dim", "activation",
                                        "batch_norm", "dropout", "nmp_layers", "scale", "vis", "actor_number"]
    m = gb.MS_HGNN_hyper(h_dim=64, bottleneck_dim=64, nmp_layers=1)
    assert m.scale == 2 and m.edge_types == 10 and m.listall is False
    assert sum(p.numel() for p in m.parameters()) == 283340
    m = gb.MS_HGNN_oridinary(h_dim=64, bottleneck_dim=64, nmp_layers=1)
    assert sum(p.numel() for p in m.parameters()) == 212168


@pytest.mark.parametrize("name", ["nba_pairwise", "l2_hyper", "crowd_hyper2"])
def test_golden_weights_regenerate(name):
    build_layer(load_golden(name))


@pytest.mark.parametrize("tn", [64, 128])
def test_chunk_permutation_is_permutation(tn):
    src = packing.chunk_permutation(tn)
    assert sorted(src.tolist()) == list(range(tn))
    w = torch.arange(3 * 2 * tn, dtype=torch.float32).reshape(3, 2 * tn)
    p = packing.permute_cols(w, tn)
    for tx in range(16):
        for j in range(tn // 16):
            pos = (j // 4) * 64 + tx * 4 + (j % 4)
            assert torch.equal(p[:, tn + pos], w[:, tn + tx + 16 * j])


def test_pack_stage_shapes_and_content():
    m = build_layer(load_golden("l2_hyper"))      # D=64, Bo=96, L=2, T=10
    for s in range(2):
        t = packing.pack_stage(m, s, torch.device("cpu"))
        assert t["node_w0t"].shape == (64, 256) and t["agg_w0t"].shape == (64, 1280)
        assert t["agg_w1t"].shape == (1280, 64) and t["post_w0t"].shape == (128, 128)
        assert t["post_w1t"].shape == (128, 64 if s == 0 else 128)
        assert t["post_b1"].numel() == (64 if s == 0 else 96)
        assert t["df_w1"].shape == (256, 16) and float(t["df_w1"][:128, 10:].abs().sum()) == 0.0
    # un-permuting recovers the transposed Linear weight
    t = packing.pack_stage(m, 0, torch.device("cpu"))
    src = packing.chunk_permutation(128)
    w = m.nmp_mlp_start.init_MLP.layers[0].weight.detach()           # (128, 64)
    un = torch.empty(64, 128)
    un[:, src] = t["init_w0t"]
    assert torch.equal(un, w.t())


def test_no_cpu_fallback():
    m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 3, 64))
    h = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=1, scale=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        h(torch.randn(2, 3, 64), torch.randn(2, 3, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gb.corr_topk_h(torch.randn(2, 3, 64), [2])
    with pytest.raises(RuntimeError):
        m.nmp_mlp_end(torch.randn(2, 128))       # containers do not compute


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "groupnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), f"{f} mentions the oracle"


def test_all_goldens_present():
    assert len(golden_names()) >= 20


@pytest.mark.skipif(not have_reference(), reason="reference tree not present")
def test_past_encoder_state_dict_matches_reference():
    import types
    for missing in ("tkinter", "glob2"):
        if missing not in sys.modules:
            stub = types.ModuleType(missing)
            stub.TRUE = True
            sys.modules[missing] = stub
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    from model.GroupNet_nba import PastEncoder as RefPastEncoder
    for scales in ([5, 11], [3, 5, 8], []):
        args = types.SimpleNamespace(hidden_dim=64, hyper_scales=scales, past_length=5)
        torch.manual_seed(9)
        r = RefPastEncoder(args)
        torch.manual_seed(9)
        m = gb.PastEncoder(args)
        rs, ms = r.state_dict(), m.state_dict()
        assert list(rs.keys()) == list(ms.keys())
        assert all(torch.equal(rs[k], ms[k]) for k in rs)
        m.load_state_dict(rs, strict=True)
        # the folded affine front-end equals the reference's module chain (eval mode)
        if len(scales) == 2:
            r.eval()
            inp = torch.randn(3 * 11, 5, 4)
            with torch.no_grad():
                tf = r.pos_encoder(r.input_fc(inp).view(33, 5, 64), num_a=33).view(3, 11, 5, 64)
                f = r.input_fc3(r.add_category(r.input_fc2(tf.contiguous().view(3, 11, 320))))
            mt, bias = m.folded_frontend(11, 5, torch.device("cpu"))
            mine = (inp.view(33, 20) @ mt + bias.repeat(3, 1)).view(3, 11, 64)
            assert (mine - f).abs().max().item() <= 1e-6


def _canon_ref(m):
    """byte(n, k) = (k/8)*(N*16) + n*16 + (k%8)*2 of the canonical operand, as a flat bf16 tensor (include/groupnet_b200.h)."""
    n, k = m.shape
    out = torch.empty(n * k, dtype=torch.bfloat16)
    mb = m.to(torch.bfloat16)
    for kk in range(k):
        for nn_ in range(n):
            out[(kk // 8) * (n * 8) + nn_ * 8 + (kk % 8)] = mb[nn_, kk]
    return out


def test_fused_kernel_weight_streams_follow_the_header_layout():
    """tc_hfuse_w / tc_npre_w are linear streams of canonical operands in MMA consumption order
    (include/groupnet_b200.h); spot-check sizes and a few chunks against a literal restatement of the layout."""
    torch.manual_seed(3)
    # h_dim 64 hyper layer: [W0_s | b0] (128 x 80), [W1_{s-1} | b1] (64 x 144) ..., closing [W0 | b0] (128 x 144), W1, b1 block
    l64 = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=1, scale=3)
    t = packing.pack_stage(l64, 0, torch.device("cpu"))
    T = 10
    assert t["tc_hfuse_w"].dtype == torch.bfloat16 and "tc_npre_w" not in t
    assert t["tc_hfuse_w"].numel() * 2 == T * (128 * 80 * 2 + 64 * 144 * 2) + 128 * 144 * 2 + 64 * 128 * 2 + 64 * 16 * 2
    agg = l64.edge_aggregation_list[0].agg_mlp
    blk = torch.zeros(128, 80)
    blk[:, :64] = agg[0].layers[0].weight.detach()
    b = agg[0].layers[0].bias.detach()
    hi = b.to(torch.bfloat16).float()
    blk[:, 64], blk[:, 65] = hi, (b - hi).to(torch.bfloat16).float()
    assert torch.equal(t["tc_hfuse_w"][:128 * 80], _canon_ref(blk))
    # second chunk of the stream is W0 of step 1, third is [W1_0 | b1_0]
    off = 2 * 128 * 80
    blk1 = torch.zeros(64, 144)
    blk1[:, :128] = agg[0].layers[1].weight.detach()
    b1 = agg[0].layers[1].bias.detach()
    hi1 = b1.to(torch.bfloat16).float()
    blk1[:, 128], blk1[:, 129], blk1[:, 130] = hi1, (b1 - hi1).to(torch.bfloat16).float(), hi1
    assert torch.equal(t["tc_hfuse_w"][off:off + 64 * 144], _canon_ref(blk1))
    # h_dim 256: sizes of both streams
    l256 = gb.MS_HGNN_hyper(256, 256, 64, 256, batch_norm=0, nmp_layers=1, scale=4)
    t = packing.pack_stage(l256, 0, torch.device("cpu"))
    assert t["tc_hfuse_w"].numel() * 2 == T * (128 * 128 * 2 + 128 * 144 * 2 + 256 * 80 * 2 + 256 * 64 * 2) \
        + 4 * 128 * 128 * 2 + 128 * 16 * 2 + 256 * 64 * 2 + 256 * 16 * 2 + 256 * 64 * 2
    assert t["tc_npre_w"].numel() * 2 == 4 * 256 * 64 * 2 + 256 * 16 * 2 + 64 * 272 * 2 + 64 * 64 * 2
    # h_dim 256 stream, chunk by chunk for step 0 / 1: W0_0[:, :128] | [W0_0[:, 128:] | b0] | W0_1 ... | [W1_0[:, :64] | b1] | W1_0[:, 64:]
    agg = l256.edge_aggregation_list[0].agg_mlp

    def hi_lo(v):
        hi_ = v.to(torch.bfloat16).float()
        return hi_, (v - hi_).to(torch.bfloat16).float()

    st = t["tc_hfuse_w"]
    w0 = agg[0].layers[0].weight.detach()
    assert torch.equal(st[:128 * 128], _canon_ref(w0[:, :128]))
    blk = torch.zeros(128, 144)
    blk[:, :128] = w0[:, 128:]
    blk[:, 128], blk[:, 129] = hi_lo(agg[0].layers[0].bias.detach())
    assert torch.equal(st[128 * 128:128 * 128 + 128 * 144], _canon_ref(blk))
    off = 2 * (128 * 128 + 128 * 144)                      # after the W0 chunks of steps 0 and 1
    w1 = agg[0].layers[1].weight.detach()
    blk = torch.zeros(256, 80)
    blk[:, :64] = w1[:, :64]
    hi, lo = hi_lo(agg[0].layers[1].bias.detach())
    blk[:, 64], blk[:, 65], blk[:, 66] = hi, lo, hi
    assert torch.equal(st[off:off + 256 * 80], _canon_ref(blk))
    assert torch.equal(st[off + 256 * 80:off + 256 * 80 + 256 * 64], _canon_ref(w1[:, 64:]))
    # closing MLP chunks start after the T aggregation steps
    post = T * (128 * 128 + 128 * 144 + 256 * 80 + 256 * 64)
    pw0 = l256.nmp_mlp_end.layers[0].weight.detach()
    assert torch.equal(st[post:post + 128 * 128], _canon_ref(pw0[:, :128]))
    node = l256.node2edge_start_mlp[0].layers
    assert torch.equal(t["tc_npre_w"][:256 * 64], _canon_ref(node[0].weight.detach()[:, :64]))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm, no GPU needed) prints one JSON line with the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["workload"].startswith("nba_synth")


def test_graphed_past_encoder_refuses_a_cpu_encoder():
    import types
    import groupnet_b200 as gb
    enc = gb.PastEncoder(types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], past_length=5)).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        gb.GraphedPastEncoder(enc, 1, 11, 5)


def test_binding_constants_match_the_header_enums():
    """groupnet_b200/_lib.py restates the header's enumerators and limits: keep them in step."""
    hdr = open(os.path.join(ROOT, "include", "groupnet_b200.h")).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(GN_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    defines = {k: int(v) for k, v in re.findall(r"#define\s+(GN_[A-Z0-9_]+)\s+(\d+)\b", hdr)}
    for name in ("GN_FP32", "GN_BF16_TC", "GN_NOISE_GIVEN", "GN_NOISE_PHILOX", "GN_NOISE_PHILOX_DEVICE_SEED"):
        assert getattr(_lib, name) == enums[name], name
    assert _lib.ABI_VERSION == defines["GN_ABI_VERSION"]
    for name in ("GN_MAX_AGENTS", "GN_MAX_SCALES"):
        if name in defines:
            assert getattr(_lib, name) == defines[name], name


def test_device_seed_walk_matches_the_by_value_philox_seeds():
    """rng="philox-device" keeps the per-call seed in an int64 tensor advanced by `add_` (two's-complement wrap);
    the eager "philox" mode computes (seed + k * stride) mod 2^64 on the host.  Same sequence of 64-bit patterns."""
    from groupnet_b200 import layers as L
    for seed in (0, 12345, (1 << 63) + 12345, (1 << 64) - 1):
        t = torch.full((1,), L._as_int64(seed), dtype=torch.int64)
        for k in range(6):
            want = (seed + L._PHILOX_CALL_STRIDE * k) & L._MASK64
            assert (int(t.item()) & L._MASK64) == want
            t.add_(L._as_int64(L._PHILOX_CALL_STRIDE))
    m = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1)
    assert m.set_rng("philox-device", 7).rng == "philox-device"
    with pytest.raises(ValueError):
        m.set_rng("philox-host")


def test_pack_cache_is_reused_until_a_parameter_changes():
    """Packed weights are rebuilt exactly when a parameter is modified in place (optimizer step, load_state_dict),
    moved / re-typed, or replaced — and not otherwise."""
    torch.manual_seed(2)
    m = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=2, scale=3)
    cpu = torch.device("cpu")
    first = m._packs.get(m, cpu)
    assert m._packs.get(m, cpu) is first                                     # unchanged -> cached
    names = [n for n, _ in m.named_parameters()]
    flat = packing.PackCache._fingerprint(m, cpu)
    assert len(flat) == 1 + 2 * len(names)                                   # every parameter, registration order
    assert list(flat[1::2]) == [p.data_ptr() for p in m.parameters()]
    w0_before = first[-1].tensors["post_w0t"].clone()                        # a transformed (K-major) copy
    with torch.no_grad():
        m.nmp_mlp_end.layers[0].weight.mul_(2.0)                             # in-place update (optimizer step)
    second = m._packs.get(m, cpu)
    assert second is not first
    assert torch.equal(second[-1].tensors["post_w0t"], 2.0 * w0_before)
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})     # copies in place -> versions bump
    third = m._packs.get(m, cpu)
    assert third is not second
    m.nmp_mlp_end.layers[0].weight = torch.nn.Parameter(torch.zeros_like(m.nmp_mlp_end.layers[0].weight))
    fourth = m._packs.get(m, cpu)                                            # Parameter object replaced
    assert fourth is not third and torch.count_nonzero(fourth[-1].tensors["post_w0t"]) == 0
    m.double().float()                                                       # storage replaced, same values
    assert m._packs.get(m, cpu) is not fourth


# ---- the reference's unused module-level helpers stay importable ----------------------
@pytest.mark.skipif(not have_reference(), reason="reference tree not present")
def test_unused_helpers_match_the_reference():
    import sys
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from model import MS_HGNN_batch as R
            assert (gb.encode_onehot([3, 1, 3, 2]) == R.encode_onehot([3, 1, 3, 2])).all()
            torch.manual_seed(5)
            a = gb.make_mlp([4, 8, 3], activation='leakyrelu', batch_norm=True, dropout=0.1)
            torch.manual_seed(5)
            b = R.make_mlp([4, 8, 3], activation='leakyrelu', batch_norm=True, dropout=0.1)
            assert [type(m) for m in a] == [type(m) for m in b]
            assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))
            torch.manual_seed(6)
            g1 = gb.sample_gumbel((3, 4, 5))
            torch.manual_seed(6)
            assert torch.equal(g1, R.sample_gumbel((3, 4, 5)))
            x = torch.randn(2, 3, 5)
            for axis in (1, 2, -1):
                assert torch.allclose(gb.my_softmax(x, axis), R.my_softmax(x, axis), atol=1e-7)
            assert torch.allclose(gb.my_softmax(x[0], 1), R.my_softmax(x[0], 1), atol=1e-7)   # the 2-D quirk (axis 0 wins)
            for hard in (False, True):
                torch.manual_seed(7)
                y1 = gb.gumbel_softmax(x, tau=0.5, hard=hard)
                torch.manual_seed(7)
                y2 = R.gumbel_softmax(x, tau=0.5, hard=hard)
                assert torch.allclose(y1, y2, atol=1e-6)
            torch.manual_seed(8)
            y1 = gb.gumbel_softmax_sample(x, tau=0.5)
            torch.manual_seed(8)
            assert torch.allclose(y1, R.gumbel_softmax_sample(x, tau=0.5), atol=1e-7)
    finally:
        sys.path.remove(REFERENCE_DIR)


# ---- runtime caches never travel with a copy of the module ----------------------------------
def test_modules_deepcopy_and_pickle_without_their_runtime_caches():
    import copy
    import io
    import types
    from groupnet_b200.packing import PackCache
    m = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=1, scale=3)
    m._packs._stages = [ctypes.c_void_p(1234)]            # what a forward leaves behind: ctypes objects do not pickle
    m._packs._key = ("stale",)
    c = copy.deepcopy(m)
    assert isinstance(c._packs, PackCache) and c._packs is not m._packs and c._packs._key is None
    assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), m.state_dict().values()))
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert r._packs._key is None and r.scale == 3 and list(r.state_dict()) == list(m.state_dict())
    # explicit invalidation and the load_state_dict hook
    m.invalidate_packs()
    assert m._packs._key is None
    m._packs._key = ("stale",)
    m.load_state_dict(c.state_dict())
    assert m._packs._key is None
    # DataParallel-style replicas get their own caches
    rep = m._replicate_for_data_parallel()
    assert rep._packs is not m._packs and rep._ws is not m._ws
    enc = gb.PastEncoder(types.SimpleNamespace(hidden_dim=64, hyper_scales=[5], past_length=5))
    enc._fold_key = ("stale",)
    e2 = copy.deepcopy(enc)
    assert e2._fold_key is None and e2._block is None
    enc.invalidate_packs()
    assert enc._fold_key is None
    dec = gb.Decoder(types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], zdim=32, past_length=5,
                                           future_length=10, num_decompose=1))
    dec._pack_key = ("stale",)
    assert copy.deepcopy(dec)._pack_key is None
