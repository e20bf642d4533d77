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
