"""SURVEY.md §8(f) rank 2 groundwork: the decoder oracle (oracle/decoder_oracle.py) is pinned against fixtures
generated from the unmodified reference and — where /root/reference exists — against the live reference module.
CPU only; the CUDA path for this row is not built yet."""
import glob
import os
import sys
import types

import numpy as np
import pytest
import torch

from decoder_schema import DecoderSchema
from helpers import FP32_REL, GOLDEN_DIR, REFERENCE_DIR, assert_close, have_reference, state_sha
from oracle import decoder_oracle as DO

DEC_DIR = os.path.join(GOLDEN_DIR, "decoder")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(DEC_DIR, "*.npz")))


def _load(name):
    z = np.load(os.path.join(DEC_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    for k in ("hidden_dim", "zdim", "past_length", "future_length", "num_decompose", "batch", "agents",
              "sample_num", "weight_seed"):
        g[k] = int(g[k])
    g["hyper_scales"] = [int(s) for s in g["hyper_scales"]]
    g["mode"], g["weight_sha256"] = str(g["mode"]), str(g["weight_sha256"])
    return g


def _schema(g):
    torch.manual_seed(g["weight_seed"])
    m = DecoderSchema(g["hidden_dim"], g["hyper_scales"], g["zdim"], g["past_length"], g["future_length"],
                      g["num_decompose"])
    assert state_sha(m) == g["weight_sha256"], "regenerated decoder weights differ from the golden run"
    return m


def _run_oracle(g, sd):
    pf = torch.from_numpy(g["past_feature_per_agent"]).repeat_interleave(g["sample_num"], dim=0)
    return DO.decoder_forward(sd, pf, torch.from_numpy(g["z"]), g["batch"], g["agents"],
                              torch.from_numpy(g["past_traj"]), torch.from_numpy(g["cur_location"]),
                              g["sample_num"], past_len=g["past_length"], future_len=g["future_length"],
                              num_decompose=g["num_decompose"], mode=g["mode"])


def test_decoder_fixtures_exist():
    assert {"nba_inference", "nba_train", "fish_three_blocks", "single_block_single_row"} <= set(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_decoder_oracle_vs_golden(name):
    g = _load(name)
    sd = {k: v.detach() for k, v in _schema(g).state_dict().items()}
    with torch.no_grad():
        out_seq, recover = _run_oracle(g, sd)
    assert out_seq.shape == g["out_seq"].shape and recover.shape == g["recover_pre_seq"].shape
    assert_close(out_seq, g["out_seq"], FP32_REL, f"{name} out_seq")
    assert_close(recover, g["recover_pre_seq"], FP32_REL, f"{name} recover_pre_seq")


def test_written_out_gru_equals_nn_gru():
    torch.manual_seed(5)
    gru = torch.nn.GRU(32, 96, 1, batch_first=True)
    sd = {f"b.encoder_past.{k}": v.detach() for k, v in gru.state_dict().items()}
    for rows, steps in ((1, 1), (7, 5), (33, 12)):
        x = torch.randn(rows, steps, 32)
        with torch.no_grad():
            _, ref = gru(x)
            got = DO.gru_last_state(sd, "b", x)
        assert_close(got, ref.squeeze(0), FP32_REL, f"gru {rows}x{steps}")


def test_decoder_sums_blocks_and_adds_the_current_location():
    """Structure of Decoder.forward (:490-502): with one block the output is that block's y_hat + cur_location,
    and shifting cur_location shifts out_seq by the same amount and nothing else."""
    g = _load("single_block_single_row")
    sd = {k: v.detach() for k, v in _schema(g).state_dict().items()}
    with torch.no_grad():
        out, rec = _run_oracle(g, sd)
        x_true = torch.from_numpy(g["past_traj"])
        f = torch.cat((torch.from_numpy(g["past_feature_per_agent"]), torch.from_numpy(g["z"])), dim=1)
        xh, yh = DO.decompose_block(sd, "decompose.0", x_true, torch.zeros_like(x_true), f, 5, 10)
        assert torch.equal(rec, xh) and torch.equal(out, yh + torch.from_numpy(g["cur_location"]))
        g2 = dict(g)
        g2["cur_location"] = g["cur_location"] + 3.0
        out2, rec2 = _run_oracle(g2, sd)
        assert torch.equal(rec2, rec)
        assert_close(out2 - 3.0, out, 1e-6, "shifted out_seq")


@pytest.mark.skipif(not have_reference(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("cfg", [
    dict(hidden_dim=64, hyper_scales=[5, 11], zdim=32, past_length=5, future_length=10, num_decompose=2,
         batch=2, agents=11, sample_num=20, mode="inference"),
    dict(hidden_dim=64, hyper_scales=[5, 11], zdim=32, past_length=5, future_length=10, num_decompose=2,
         batch=4, agents=11, sample_num=1, mode="train"),
    dict(hidden_dim=16, hyper_scales=[2, 3, 4], zdim=8, past_length=3, future_length=4, num_decompose=4,
         batch=1, agents=5, sample_num=3, mode="inference"),
])
def test_decoder_oracle_vs_live_reference(cfg):
    sys.path.insert(0, REFERENCE_DIR)
    for missing in ("tkinter", "glob2"):                 # model/GroupNet_nba.py:2, model/utils.py:9
        if missing not in sys.modules:
            stub = types.ModuleType(missing)
            stub.TRUE = True
            sys.modules[missing] = stub
    from model.GroupNet_nba import Decoder
    args = types.SimpleNamespace(**{k: cfg[k] for k in ("hidden_dim", "hyper_scales", "zdim", "past_length",
                                                        "future_length", "num_decompose")})
    torch.manual_seed(11)
    ref = Decoder(args).eval()
    torch.manual_seed(11)
    mine = DecoderSchema(cfg["hidden_dim"], cfg["hyper_scales"], cfg["zdim"], cfg["past_length"],
                         cfg["future_length"], cfg["num_decompose"])
    # same keys, shapes, order and (seeded) values as the reference module
    assert [(k, tuple(v.shape)) for k, v in ref.state_dict().items()] == \
           [(k, tuple(v.shape)) for k, v in mine.state_dict().items()]
    assert state_sha(ref) == state_sha(mine)
    a = cfg["batch"] * cfg["agents"]
    gen = torch.Generator().manual_seed(12)
    pf = torch.randn(a, (2 + len(cfg["hyper_scales"])) * cfg["hidden_dim"], generator=gen) \
        .repeat_interleave(cfg["sample_num"], dim=0)
    z = torch.randn(a * cfg["sample_num"], cfg["zdim"], generator=gen)
    past = torch.randn(a, cfg["past_length"], 2, generator=gen)
    cur = torch.randn(a, 1, 2, generator=gen)
    with torch.no_grad():
        r_out, r_rec = ref(pf, z, cfg["batch"], cfg["agents"], past, cur, cfg["sample_num"], mode=cfg["mode"])
        o_out, o_rec = DO.decoder_forward({k: v.detach() for k, v in ref.state_dict().items()}, pf, z, cfg["batch"],
                                          cfg["agents"], past, cur, cfg["sample_num"], past_len=cfg["past_length"],
                                          future_len=cfg["future_length"], num_decompose=cfg["num_decompose"],
                                          mode=cfg["mode"])
    assert o_out.shape == r_out.shape and o_rec.shape == r_rec.shape
    assert_close(o_out, r_out, FP32_REL, "out_seq")
    assert_close(o_rec, r_rec, FP32_REL, "recover_pre_seq")


# ---- host side of the CUDA path: packing (no GPU needed) -------------------------------------------
def _unpermute(packed, tn):
    """Inverse of packing.permute_cols: K-major matrix with its original column order."""
    from groupnet_b200.packing import chunk_permutation
    src = chunk_permutation(tn)
    out = torch.empty_like(packed)
    for c0 in range(0, packed.shape[1], tn):
        out[:, c0 + src] = packed[:, c0:c0 + tn]
    return out


def _emulate_block(p, feat, x_true, x_hat, past_len, future_len):
    """csrc/gn_decoder_simt.cu's algorithm on the PACKED tensors (padded gates, padded K, padded outputs)."""
    rows = feat.shape[0]
    res = x_true - x_hat
    xpad = torch.nn.functional.pad(res, (0, 0, 1, 1))                       # time padding 1
    wx, wh = _unpermute(p["gru_wx"], 128), _unpermute(p["gru_wh"], 128)     # (32,384), (96,384)
    br, bz, b_in, b_hn = p["gru_b"]
    h = torch.zeros(rows, 128)
    for t in range(past_len):
        win = xpad[:, t:t + 3]                                              # (R, 3, 2): steps t-1, t, t+1
        e = torch.relu(torch.einsum("rkc,ock->ro", win, p["conv_w"]) + p["conv_b"])
        r = torch.sigmoid(e @ wx[:, :128] + h[:, :96] @ wh[:, :128] + br)
        n = torch.tanh(e @ wx[:, 256:] + b_in + r * (h[:, :96] @ wh[:, 256:] + b_hn))
        zg = torch.sigmoid(e @ wx[:, 128:256] + h[:, :96] @ wh[:, 128:256] + bz)
        h = (1 - zg) * n + zg * h
    assert torch.count_nonzero(h[:, 96:]) == 0                              # padded state columns stay exactly 0
    kp = p["x_w0"].shape[0]
    full = torch.zeros(rows, kp)
    full[:, :feat.shape[1]] = feat
    full[:, feat.shape[1]:feat.shape[1] + 96] = h[:, :96]
    outs = []
    for tag, width in (("x", 2 * past_len), ("y", 2 * future_len)):
        h1 = torch.relu(full @ _unpermute(p[f"{tag}_w0"], 128) + p[f"{tag}_b0"])
        h2 = torch.relu(h1 @ _unpermute(p[f"{tag}_w1"], 128) + p[f"{tag}_b1"])
        o = h2 @ _unpermute(p[f"{tag}_w2"], 64) + p[f"{tag}_b2"]
        assert torch.count_nonzero(o[:, width:]) == 0
        outs.append(o[:, :width])
    return outs[0].reshape(rows, past_len, 2), outs[1].reshape(rows, future_len, 2)


@pytest.mark.parametrize("name", NAMES)
def test_packed_decoder_weights_reproduce_the_oracle(name):
    from groupnet_b200.packing import pack_decoder_block
    g = _load(name)
    m = _schema(g)
    for blk in m.decompose:                      # the fixtures' GRU / conv biases are zero-initialised: exercise them
        torch.nn.init.normal_(blk.encoder_past.bias_ih_l0, std=0.3)
        torch.nn.init.normal_(blk.encoder_past.bias_hh_l0, std=0.3)
        torch.nn.init.normal_(blk.conv_past.bias, std=0.3)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    s = g["sample_num"]
    feat = torch.cat((torch.from_numpy(g["past_feature_per_agent"]).repeat_interleave(s, dim=0),
                      torch.from_numpy(g["z"])), dim=1)
    x_true = torch.from_numpy(g["past_traj"]).repeat_interleave(s, dim=0)
    x_hat = torch.zeros_like(x_true)
    with torch.no_grad():
        for i, blk in enumerate(m.decompose):
            p = pack_decoder_block(blk, torch.device("cpu"))
            assert p["gru_wx"].shape == (32, 384) and p["gru_wh"].shape == (96, 384) and p["gru_b"].shape == (4, 128)
            ref_x, ref_y = DO.decompose_block(sd, f"decompose.{i}", x_true, x_hat, feat, g["past_length"],
                                              g["future_length"])
            got_x, got_y = _emulate_block(p, feat, x_true, x_hat, g["past_length"], g["future_length"])
            assert_close(got_x, ref_x, FP32_REL, f"{name} block {i} x_hat")
            assert_close(got_y, ref_y, FP32_REL, f"{name} block {i} y_hat")
            x_hat = ref_x


def test_decoder_dropin_refuses_cpu_tensors_and_autograd():
    g = _load("single_block_single_row")
    m = _schema(g)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 256), torch.zeros(1, 32), 1, 1, torch.zeros(1, 5, 2), torch.zeros(1, 1, 2), 1)


# ---- numeric plan of the tensor-core decoder path, checked before the kernel exists -------
def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _lin_bf16(x, w, b):
    """bf16 operands, fp32 accumulate, fp32 bias: what a tcgen05 kind::f16 GEMM computes."""
    return _bf16(x) @ _bf16(w).t() + b


@pytest.mark.parametrize("name", NAMES)
def test_bf16_operand_plan_for_the_decoder_stays_inside_the_bf16_tolerance(name):
    """Every Linear (GRU gates included) with bf16 operands and fp32 accumulation, the GRU state carried in fp32 and
    only its operand copy rounded, conv1d in fp32: 3e-3 .. 5e-3 of max|ref| on the fixtures, against the 2e-2 bar."""
    import torch.nn.functional as F
    g = _load(name)
    sd = {k: v.detach() for k, v in _schema(g).state_dict().items()}
    s, tp, tf = g["sample_num"], g["past_length"], g["future_length"]
    f = torch.cat((torch.from_numpy(g["past_feature_per_agent"]).repeat_interleave(s, 0), torch.from_numpy(g["z"])), 1)
    x_true = torch.from_numpy(g["past_traj"]).repeat_interleave(s, 0)
    x_hat, pred, rec = torch.zeros_like(x_true), 0.0, 0.0
    with torch.no_grad():
        for i in range(g["num_decompose"]):
            p = f"decompose.{i}"
            e = torch.relu(F.conv1d((x_true - x_hat).transpose(1, 2), sd[f"{p}.conv_past.weight"],
                                    sd[f"{p}.conv_past.bias"], padding=1)).transpose(1, 2)
            h = torch.zeros(x_true.shape[0], 96)
            for t in range(tp):
                gi = _lin_bf16(e[:, t], sd[f"{p}.encoder_past.weight_ih_l0"], sd[f"{p}.encoder_past.bias_ih_l0"])
                gh = _lin_bf16(h, sd[f"{p}.encoder_past.weight_hh_l0"], sd[f"{p}.encoder_past.bias_hh_l0"])
                rg = torch.sigmoid(gi[:, :96] + gh[:, :96])
                zg = torch.sigmoid(gi[:, 96:192] + gh[:, 96:192])
                h = (1 - zg) * torch.tanh(gi[:, 192:] + rg * gh[:, 192:]) + zg * h
            feat, outs = torch.cat((f, h), 1), []
            for tag, width in (("decoder_x", tp), ("decoder_y", tf)):
                a = feat
                for li in range(3):
                    a = _lin_bf16(a, sd[f"{p}.{tag}.layers.{li}.weight"], sd[f"{p}.{tag}.layers.{li}.bias"])
                    a = torch.relu(a) if li < 2 else a
                outs.append(a.view(-1, width, 2))
            x_hat = outs[0]
            pred, rec = pred + outs[1], rec + x_hat
        out = pred + torch.from_numpy(g["cur_location"]).repeat_interleave(s, 0)
    assert_close(out, torch.from_numpy(g["out_seq"]).reshape(out.shape), 1e-2, f"{name} out_seq (bf16 plan)")
    assert_close(rec, g["recover_pre_seq"], 1e-2, f"{name} recover_pre_seq (bf16 plan)")


def _uncanon(flat, n, k):
    """flat canonical bf16 operand [K/8][N][8] -> (N, K) fp32."""
    return flat.view(k // 8, n, 8).permute(1, 0, 2).reshape(n, k).float()


@pytest.mark.parametrize("name", NAMES)
def test_tensor_core_decoder_packs_evaluate_to_the_reference(name):
    """The host side of gn_decoder_fwd_tc: the algorithm of csrc/gn_decoder_tc.cu (gate matrix [r | z | n_x | n_h] over
    K = [e | h], both first Linears as one 1024-row operand, zero-padded last Linears) evaluated on the PACKED bf16
    tensors with bf16 operand rounding must reproduce the reference fixtures inside the bf16 bar."""
    import torch.nn.functional as F
    from groupnet_b200.packing import pack_decoder_block_tc
    g = _load(name)
    m = _schema(g)
    s, tp, tf = g["sample_num"], g["past_length"], g["future_length"]
    bf = lambda v: v.to(torch.bfloat16).float()
    f = torch.cat((torch.from_numpy(g["past_feature_per_agent"]).repeat_interleave(s, 0), torch.from_numpy(g["z"])), 1)
    x_true = torch.from_numpy(g["past_traj"]).repeat_interleave(s, 0)
    x_hat, pred, rec = torch.zeros_like(x_true), 0.0, 0.0
    kf = f.shape[1] + 96
    with torch.no_grad():
        for blk in m.decompose:
            p = pack_decoder_block_tc(blk, torch.device("cpu"))
            assert p["gru_w"].dtype == torch.bfloat16 and p["gru_w"].numel() == 2 * 192 * 128
            wg = torch.cat([_uncanon(p["gru_w"][:192 * 128], 192, 128), _uncanon(p["gru_w"][192 * 128:], 192, 128)])
            e = torch.relu(F.conv1d((x_true - x_hat).transpose(1, 2), p["conv_w"], p["conv_b"], padding=1)).transpose(1, 2)
            h = torch.zeros(x_true.shape[0], 96)
            for t in range(tp):
                acc = bf(torch.cat([e[:, t], h], 1)) @ wg.t()                  # (R, 384): r | z | n_x | n_h
                rg = torch.sigmoid(acc[:, :96] + p["gru_b"][0, :96])
                zg = torch.sigmoid(acc[:, 96:192] + p["gru_b"][1, :96])
                ng = torch.tanh(acc[:, 192:288] + p["gru_b"][2, :96] + rg * (acc[:, 288:] + p["gru_b"][3, :96]))
                h = (1 - zg) * ng + zg * h
            feat = bf(torch.cat((f, h), 1))
            hid1 = bf(torch.relu(feat @ _uncanon(p["w0"], 1024, kf).t() + p["b0"]))
            outs = []
            for i, (tag, width) in enumerate((("x", tp), ("y", tf))):
                hid2 = bf(torch.relu(hid1[:, 512 * i:512 * (i + 1)] @ _uncanon(p[f"{tag}_w1"], 256, 512).t() + p[f"{tag}_b1"]))
                pad = p[f"{tag}_b2"].numel()
                assert pad % 16 == 0 and pad >= 2 * width
                o = hid2 @ _uncanon(p[f"{tag}_w2"], pad, 256).t() + p[f"{tag}_b2"]
                assert torch.all(o[:, 2 * width:] == 0)
                outs.append(o[:, :2 * width].reshape(-1, width, 2))
            x_hat = outs[0]
            pred, rec = pred + outs[1], rec + x_hat
        out = pred + torch.from_numpy(g["cur_location"]).repeat_interleave(s, 0)
    assert_close(out, torch.from_numpy(g["out_seq"]).reshape(out.shape), 1e-2, f"{name} out_seq (tc packs)")
    assert_close(rec, g["recover_pre_seq"], 1e-2, f"{name} recover_pre_seq (tc packs)")


@pytest.mark.parametrize("name", NAMES)
def test_fused_mlp_stage_stream_holds_the_weights_in_the_issue_order(name):
    """decoder_mlp_stream: walking the 16 KB stages in the fused kernel's issue order (G1(0) G1(1) G2(0) G1(2) G2(1) G1(3)
    G2(2) G2(3) G3 per MLP, csrc/gn_decoder_tc.cu) must reassemble bf16(W0), bf16(W1) and the zero-padded bf16(W2) of
    decoder_x and decoder_y exactly, and the bias block must hold b0 | b1 | b2 at the offsets the drain reads."""
    from groupnet_b200.packing import decoder_mlp_stream
    g = _load(name)
    m = _schema(g)
    for blk in m.decompose:
        stream, bias = decoder_mlp_stream(blk, torch.device("cpu"))
        kf = blk.decoder_x.layers[0].weight.shape[1]
        if kf % 64 or kf > 384:
            assert stream.numel() == 0
            continue
        nk1, slot = kf // 64, 8192
        assert stream.dtype == torch.bfloat16 and stream.numel() == 2 * (4 * nk1 + 20) * slot
        pos = 0

        def take(rows):
            nonlocal pos
            flat = stream[pos * slot:(pos + 1) * slot]
            pos += 1
            assert torch.all(flat[rows * 64:] == 0)
            return _uncanon(flat[:rows * 64], rows, 64)

        for i, mlp in enumerate((blk.decoder_x, blk.decoder_y)):
            w0 = torch.zeros(512, kf)
            w1 = torch.zeros(256, 512)
            w2 = torch.zeros(32, 256)

            def g1(c):
                for ks in range(nk1):
                    w0[c * 128:(c + 1) * 128, ks * 64:(ks + 1) * 64] = take(128)

            def g2(c):
                for kh in range(2):
                    for nh in range(2):
                        w1[nh * 128:(nh + 1) * 128, c * 128 + kh * 64:c * 128 + (kh + 1) * 64] = take(128)

            g1(0); g1(1); g2(0); g1(2); g2(1); g1(3); g2(2); g2(3)
            for ks in range(4):
                w2[:, ks * 64:(ks + 1) * 64] = take(32)
            l0, l1, l2 = mlp.layers
            bf = lambda v: v.detach().to(torch.bfloat16).float()
            assert torch.equal(w0, bf(l0.weight)) and torch.equal(w1, bf(l1.weight))
            n_out = l2.weight.shape[0]
            assert torch.equal(w2[:n_out], bf(l2.weight)) and torch.all(w2[n_out:] == 0)
            assert torch.equal(bias[i * 512:(i + 1) * 512], l0.bias.detach())
            assert torch.equal(bias[1024 + i * 256:1024 + (i + 1) * 256], l1.bias.detach())
            assert torch.equal(bias[1536 + i * 32:1536 + i * 32 + n_out], l2.bias.detach())
            assert torch.all(bias[1536 + i * 32 + n_out:1536 + (i + 1) * 32] == 0)
        assert pos * slot == stream.numel()


# ---- randomized differential pin against the live reference (build container only) ---------------------
try:
    from hypothesis import HealthCheck, given, settings, strategies as st
    _HAVE_HYP = True
except Exception:                                   # pragma: no cover
    _HAVE_HYP = False

if _HAVE_HYP:
    @pytest.mark.skipif(not have_reference(), reason="reference tree not present (GPU box)")
    @settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(hidden=st.sampled_from([8, 16, 32]), n_scales=st.integers(0, 3), zdim=st.sampled_from([4, 8, 32]),
           tp=st.integers(1, 7), tf=st.integers(1, 9), blocks=st.integers(1, 3), batch=st.integers(1, 3),
           agents=st.integers(1, 6), s=st.integers(1, 4), inference=st.booleans())
    def test_decoder_oracle_vs_live_reference_random_configs(hidden, n_scales, zdim, tp, tf, blocks, batch, agents, s,
                                                             inference):
        sys.path.insert(0, REFERENCE_DIR)
        for missing in ("tkinter", "glob2"):
            if missing not in sys.modules:
                stub = types.ModuleType(missing)
                stub.TRUE = True
                sys.modules[missing] = stub
        from model.GroupNet_nba import Decoder
        scales = list(range(2, 2 + n_scales))
        torch.manual_seed(hidden + tp + tf)
        ref = Decoder(types.SimpleNamespace(hidden_dim=hidden, hyper_scales=scales, zdim=zdim, past_length=tp,
                                            future_length=tf, num_decompose=blocks)).eval()
        for blk in ref.decompose:                    # zero-initialised by the reference: make them count
            torch.nn.init.normal_(blk.encoder_past.bias_ih_l0, std=0.3)
            torch.nn.init.normal_(blk.encoder_past.bias_hh_l0, std=0.3)
            torch.nn.init.normal_(blk.conv_past.bias, std=0.3)
        a = batch * agents
        gen = torch.Generator().manual_seed(a + s)
        pf = torch.randn(a, (2 + n_scales) * hidden, generator=gen).repeat_interleave(s, dim=0)
        z = torch.randn(a * s, zdim, generator=gen)
        past, cur = torch.randn(a, tp, 2, generator=gen), torch.randn(a, 1, 2, generator=gen)
        mode = "inference" if inference else "train"
        with torch.no_grad():
            r_out, r_rec = ref(pf, z, batch, agents, past, cur, s, mode=mode)
            o_out, o_rec = DO.decoder_forward({k: v.detach() for k, v in ref.state_dict().items()}, pf, z, batch,
                                              agents, past, cur, s, past_len=tp, future_len=tf,
                                              num_decompose=blocks, mode=mode)
        assert o_out.shape == r_out.shape and o_rec.shape == r_rec.shape
        assert_close(o_out, r_out, FP32_REL, "out_seq")
        assert_close(o_rec, r_rec, FP32_REL, "recover_pre_seq")
