"""Trajectory decoder on the GPU (gn_decoder_fwd through groupnet_b200.Decoder) against the fixtures generated
from the reference and against the CPU oracle.  fp32 criterion: 1e-5 * max|ref|."""

import pytest
import torch

from helpers import BF16_REL, FP32_REL, assert_close
from oracle import decoder_oracle as DO
from test_decoder_oracle import NAMES, _load, _schema

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


@pytest.mark.parametrize("name", NAMES)
def test_decoder_vs_golden(name):
    g = _load(name)
    m = _schema(g).to(DEV)
    s = g["sample_num"]
    pf = torch.from_numpy(g["past_feature_per_agent"]).repeat_interleave(s, dim=0).to(DEV)
    out_seq, recover = m(pf, torch.from_numpy(g["z"]).to(DEV), g["batch"], g["agents"],
                         torch.from_numpy(g["past_traj"]).to(DEV), torch.from_numpy(g["cur_location"]).to(DEV),
                         s, mode=g["mode"])
    assert tuple(out_seq.shape) == g["out_seq"].shape and tuple(recover.shape) == g["recover_pre_seq"].shape
    assert_close(out_seq, g["out_seq"], FP32_REL, f"{name} out_seq")
    assert_close(recover, g["recover_pre_seq"], FP32_REL, f"{name} recover_pre_seq")


@pytest.mark.parametrize("batch,agents,s,blocks,hidden,scales,zdim,tp,tf", [
    (7, 11, 20, 2, 64, [5, 11], 32, 5, 10),       # 1,540 rows: 24 full tiles + a ragged one
    (1, 13, 5, 3, 32, [3], 16, 8, 12),            # 65 rows: one full tile + 1 row
    (2, 8, 4, 1, 64, [2, 4, 8, 16], 32, 1, 32),   # single step GRU, widest output, feature width 416
    (300, 11, 1, 2, 64, [5, 11], 32, 5, 10),      # training-style call, more tiles than SMs? (52) no; many agents
])
def test_decoder_vs_oracle(batch, agents, s, blocks, hidden, scales, zdim, tp, tf):
    import groupnet_b200 as gb
    import types
    torch.manual_seed(batch + agents + s)
    m = gb.Decoder(types.SimpleNamespace(hidden_dim=hidden, hyper_scales=scales, zdim=zdim, past_length=tp,
                                         future_length=tf, num_decompose=blocks))
    for blk in m.decompose:                        # non-zero conv / GRU biases (zero-initialised by the reference)
        torch.nn.init.normal_(blk.encoder_past.bias_ih_l0, std=0.3)
        torch.nn.init.normal_(blk.encoder_past.bias_hh_l0, std=0.3)
        torch.nn.init.normal_(blk.conv_past.bias, std=0.3)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    a = batch * agents
    gen = torch.Generator().manual_seed(3)
    pf = torch.randn(a, (2 + len(scales)) * hidden, generator=gen).repeat_interleave(s, dim=0)
    z = torch.randn(a * s, zdim, generator=gen)
    past = torch.randn(a, tp, 2, generator=gen)
    cur = torch.randn(a, 1, 2, generator=gen)
    ref_out, ref_rec = DO.decoder_forward(sd, pf, z, batch, agents, past, cur, s, past_len=tp, future_len=tf,
                                          num_decompose=blocks, mode="inference")
    m = m.to(DEV)
    out, rec = m(pf.to(DEV), z.to(DEV), batch, agents, past.to(DEV), cur.to(DEV), s, mode="inference")
    assert tuple(out.shape) == (a, s, tf, 2)
    assert_close(out, ref_out, FP32_REL, "out_seq")
    assert_close(rec, ref_rec, FP32_REL, "recover_pre_seq")
    out2, rec2 = m(pf.to(DEV), z.to(DEV), batch, agents, past.to(DEV), cur.to(DEV), s, mode="inference")
    assert torch.equal(out, out2) and torch.equal(rec, rec2)          # the running sums restart on every call


@pytest.mark.parametrize("name", NAMES)
def test_decoder_bf16_tensor_core_path_vs_golden(name):
    """gn_decoder_fwd_tc (tcgen05: GRU step as one MMA chain per time step, MLPs as row-tile GEMMs) against the
    fixtures of the unmodified reference; bf16 criterion 2e-2 * max|ref|."""
    g = _load(name)
    m = _schema(g).to(DEV).set_precision("bf16")
    s = g["sample_num"]
    pf = torch.from_numpy(g["past_feature_per_agent"]).repeat_interleave(s, dim=0).to(DEV)
    args = (pf, torch.from_numpy(g["z"]).to(DEV), g["batch"], g["agents"], torch.from_numpy(g["past_traj"]).to(DEV),
            torch.from_numpy(g["cur_location"]).to(DEV), s)
    out_seq, recover = m(*args, mode=g["mode"])
    assert tuple(out_seq.shape) == g["out_seq"].shape and tuple(recover.shape) == g["recover_pre_seq"].shape
    assert_close(out_seq, g["out_seq"], BF16_REL, f"{name} out_seq (bf16)")
    assert_close(recover, g["recover_pre_seq"], BF16_REL, f"{name} recover_pre_seq (bf16)")
    # the two paths share one module: switching back restores the 1e-5 path
    out32, rec32 = m.set_precision("fp32")(*args, mode=g["mode"])
    assert_close(out32, g["out_seq"], FP32_REL, f"{name} out_seq (fp32 after bf16)")
    assert_close(rec32, g["recover_pre_seq"], FP32_REL, f"{name} recover_pre_seq (fp32 after bf16)")


@pytest.mark.parametrize("batch,agents,s,blocks,hidden,scales,zdim,tp,tf", [
    (7, 11, 20, 2, 64, [5, 11], 32, 5, 10),       # 1,540 rows: 12 full tiles + a ragged one
    (1, 13, 5, 3, 32, [3], 16, 8, 12),            # 65 rows, three blocks, feature width 208 (K tail of 80)
    (2, 8, 4, 1, 64, [2, 4, 8, 16], 32, 1, 32),   # single step GRU, widest output (64 columns), feature width 512
    (180, 11, 10, 2, 64, [5, 11], 32, 5, 10),     # 19,800 rows = 155 tiles: more tiles than SMs
])
def test_decoder_bf16_tensor_core_path_vs_oracle(batch, agents, s, blocks, hidden, scales, zdim, tp, tf):
    import groupnet_b200 as gb
    import types
    torch.manual_seed(batch + agents + s)
    m = gb.Decoder(types.SimpleNamespace(hidden_dim=hidden, hyper_scales=scales, zdim=zdim, past_length=tp,
                                         future_length=tf, num_decompose=blocks))
    for blk in m.decompose:
        torch.nn.init.normal_(blk.encoder_past.bias_ih_l0, std=0.3)
        torch.nn.init.normal_(blk.encoder_past.bias_hh_l0, std=0.3)
        torch.nn.init.normal_(blk.conv_past.bias, std=0.3)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    a = batch * agents
    gen = torch.Generator().manual_seed(3)
    pf = torch.randn(a, (2 + len(scales)) * hidden, generator=gen).repeat_interleave(s, dim=0)
    z = torch.randn(a * s, zdim, generator=gen)
    past = torch.randn(a, tp, 2, generator=gen)
    cur = torch.randn(a, 1, 2, generator=gen)
    ref_out, ref_rec = DO.decoder_forward(sd, pf, z, batch, agents, past, cur, s, past_len=tp, future_len=tf,
                                          num_decompose=blocks, mode="inference")
    m = m.to(DEV).set_precision("bf16")
    out, rec = m(pf.to(DEV), z.to(DEV), batch, agents, past.to(DEV), cur.to(DEV), s, mode="inference")
    assert tuple(out.shape) == (a, s, tf, 2)
    assert_close(out, ref_out, BF16_REL, "out_seq (bf16)")
    assert_close(rec, ref_rec, BF16_REL, "recover_pre_seq (bf16)")
    out2, rec2 = m(pf.to(DEV), z.to(DEV), batch, agents, past.to(DEV), cur.to(DEV), s, mode="inference")
    assert torch.equal(out, out2) and torch.equal(rec, rec2)


_ROWTILE_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from test_decoder_oracle import _load, _schema
g = _load("nba_inference")
s = g["sample_num"]
with torch.no_grad():
    m = _schema(g).to("cuda:0").set_precision("bf16")
    pf = torch.from_numpy(g["past_feature_per_agent"]).repeat_interleave(s, dim=0).cuda()
    out, rec = m(pf, torch.from_numpy(g["z"]).cuda(), g["batch"], g["agents"], torch.from_numpy(g["past_traj"]).cuda(),
                 torch.from_numpy(g["cur_location"]).cuda(), s, mode=g["mode"])
np.savez(sys.argv[2], out=out.cpu().numpy(), rec=rec.cpu().numpy())
"""


def test_decoder_bf16_fused_and_row_tile_mlp_paths_agree(tmp_path):
    """The NBA shape runs the fused MLP kernel (decoder_mlp_fused_kernel); GN_DECODER_MLP=rowtile (read once per process,
    hence the child process) runs the same module on the row-tile GEMMs, the path of feature widths above 384 / not a
    multiple of 64.  Both must meet the bf16 bar against the reference fixture and agree with each other inside it."""
    import os
    import subprocess
    import sys
    import numpy as np
    g = _load("nba_inference")
    s = g["sample_num"]
    pf = torch.from_numpy(g["past_feature_per_agent"]).repeat_interleave(s, dim=0).to(DEV)
    m = _schema(g).to(DEV).set_precision("bf16")
    out_f, rec_f = m(pf, torch.from_numpy(g["z"]).to(DEV), g["batch"], g["agents"], torch.from_numpy(g["past_traj"]).to(DEV),
                     torch.from_numpy(g["cur_location"]).to(DEV), s, mode=g["mode"])
    assert m._packed[0][0]["mlp_stream"].numel() > 0                      # the fused kernel's stream exists at this shape
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = str(tmp_path / "rowtile.npz")
    r = subprocess.run([sys.executable, "-c", _ROWTILE_CHILD, root, dst], env=dict(os.environ, GN_DECODER_MLP="rowtile"),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    child = np.load(dst)
    out_r, rec_r = torch.from_numpy(child["out"]), torch.from_numpy(child["rec"])
    for out, rec, tag in ((out_f, rec_f, "fused"), (out_r, rec_r, "row-tile")):
        assert_close(out, g["out_seq"], BF16_REL, f"out_seq ({tag})")
        assert_close(rec, g["recover_pre_seq"], BF16_REL, f"recover_pre_seq ({tag})")
    assert_close(out_f, out_r, 5e-3, "out_seq fused vs row-tile")
    assert_close(rec_f, rec_r, 5e-3, "recover_pre_seq fused vs row-tile")


def test_decoder_error_paths():
    g = _load("single_block_single_row")
    m = _schema(g).to(DEV)
    ok = (torch.zeros(1, 256, device=DEV), torch.zeros(1, 32, device=DEV), 1, 1, torch.zeros(1, 5, 2, device=DEV),
          torch.zeros(1, 1, 2, device=DEV), 1)
    m(*ok)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 252, device=DEV), *ok[1:])                    # feature width does not match the weights
    with pytest.raises(RuntimeError):
        m(ok[0].double(), *ok[1:])
    with torch.enable_grad(), pytest.raises(NotImplementedError):
        m(*ok)
    out, rec = m(torch.zeros(0, 256, device=DEV), torch.zeros(0, 32, device=DEV), 0, 1,
                 torch.zeros(0, 5, 2, device=DEV), torch.zeros(0, 1, 2, device=DEV), 1)
    assert out.shape == (0, 10, 2) and rec.shape == (0, 5, 2)
