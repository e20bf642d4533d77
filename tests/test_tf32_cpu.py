"""Host-side checks of the 3xTF32 path (no GPU): operand split numerics and the weight-stream layout the
chain engine consumes (csrc/gn_chain_tf32.cu, include/groupnet_b200.h)."""
import torch

import groupnet_b200 as gb
from groupnet_b200 import packing


def _tf32_trunc(x: torch.Tensor) -> torch.Tensor:
    """What the tensor core reads of an fp32 word: sign, exponent, 10 mantissa bits."""
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def test_split_is_exact_and_hi_is_a_tf32_number():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, generator=g) * torch.logspace(-6, 6, 4096)
    hi, lo = packing.tf32_split(x)
    assert torch.equal(_tf32_trunc(hi), hi) and torch.equal(_tf32_trunc(lo), lo)
    assert ((x - hi).abs() <= x.abs() * 2.0 ** -11).all()
    assert ((x.double() - hi.double() - lo.double()).abs() <= x.abs().double() * 2.0 ** -21).all()


def test_three_term_product_reaches_fp32_grade():
    """(A_lo B_hi + A_hi B_lo + A_hi B_hi) with tf32-truncated operands and fp32 accumulation vs float64:
    the plan of csrc/gn_tf32.cuh stays ~2 orders of magnitude inside the 1e-5 bound, a plain tf32 product does not."""
    g = torch.Generator().manual_seed(1)
    a = torch.randn(128, 256, generator=g)
    w = torch.randn(64, 256, generator=g) / 16
    ref = a.double() @ w.double().t()
    # device-side split of the activation: hi rounded, lo = exact remainder read through tf32 truncation
    a_hi, _ = packing.tf32_split(a)
    a_lo = _tf32_trunc(a - a_hi)
    w_hi, w_lo = packing.tf32_split(w)
    x3 = (a_lo @ w_hi.t()) + (a_hi @ w_lo.t()) + (a_hi @ w_hi.t())
    one = a_hi @ w_hi.t()
    scale = ref.abs().max()
    assert (x3.double() - ref).abs().max() / scale < 2e-6
    assert (one.double() - ref).abs().max() / scale > 1e-5


def _decode(stream: torch.Tensor, shapes, stage_bytes, tail=0):
    """Inverse of packing.tf_stream: rebuild every (N, K) matrix (hi + lo) from the chunked canonical stream."""
    out, off = [], 0
    for n, k in shapes:
        kc = packing.tf_chunk_k(n, k, stage_bytes)
        assert n * kc * 8 <= stage_bytes and k % kc == 0 and kc % 8 == 0
        w = torch.zeros(n, k, dtype=torch.float64)
        for c in range(k // kc):
            for part in range(2):                                   # hi, then lo
                blk = stream[off:off + n * kc].reshape(kc // 4, n, 4)   # [k-group][row][4]: byte(n,k) = (k/4)*N*16 + n*16 + (k%4)*4
                w[:, c * kc:(c + 1) * kc] += blk.permute(1, 0, 2).reshape(n, kc).double()
                off += n * kc
        out.append(w)
    assert off + tail == stream.numel()
    return out


def test_weight_streams_follow_the_documented_order():
    torch.manual_seed(3)
    pair = gb.MS_HGNN_oridinary(16, 64, 64, 64, batch_norm=0, nmp_layers=1)
    t = packing.pack_stage(pair, 0, torch.device("cpu"))
    ms = pair.nmp_mlp_start
    att = pair.attention_mlp[0].layers[0].weight
    agg = pair.edge_aggregation_list[0].agg_mlp
    node = pair.node2edge_start_mlp[0].layers

    def close(a, b):
        return (a - b.detach().double()).abs().max() <= b.abs().max() * 2.0 ** -21

    s64, s32 = 65536, 32768
    assert packing.tf_stage_bytes(0, 1, True) == s64 and packing.tf_stage_bytes(64, 1, False) == s64
    assert packing.tf_stage_bytes(64, 2, False) == s32 and packing.tf_stage_bytes(128, 1, False) == s32
    w = _decode(t["tf_chain_w"], [(128, 64), (64, 128), (128, 64), (128, 64)], s64, tail=128 + 128 * 6)
    tail = t["tf_chain_w"][-(128 + 128 * 6):]
    assert torch.equal(tail[:128], ms.MLP_factor.layers[1].weight.detach().reshape(-1))
    w4 = tail[128:].reshape(128, 6)                               # the distribution head, k-major rows of T = 6 logits
    assert torch.equal(w4[:, :6], ms.MLP_distribution.layers[1].weight.detach().t()) and w4.shape[1] == 6
    assert close(w[0], ms.init_MLP.layers[0].weight) and close(w[1], ms.init_MLP.layers[1].weight)
    assert close(w[2], ms.MLP_factor.layers[0].weight) and close(w[3], ms.MLP_distribution.layers[0].weight)
    w = _decode(t["tf_pre_w"], [(128, 64), (64, 128), (128, 64), (64, 128), (64, 64), (128, 64)], s64)
    assert close(w[5], ms.init_MLP.layers[0].weight)              # pairwise layers: Y = x' W_init0^T in the prologue
    assert close(w[0], node[0].weight[:128]) and close(w[2], node[0].weight[128:])
    assert close(w[1], node[1].weight[:, :128]) and close(w[3], node[1].weight[:, 128:])
    assert close(w[4], torch.cat((att[:, :64], att[:, 64:]), dim=0))
    w = _decode(t["tf_aggin_w"], [(128, 64)] * 6, s64)
    assert all(close(w[i], agg[i].layers[0].weight) for i in range(6))
    w = _decode(t["tf_aggout_w"], [(64, 64)] * 12, s32)
    for i in range(6):
        assert close(torch.cat((w[2 * i], w[2 * i + 1]), dim=1), agg[i].layers[1].weight)
    w = _decode(t["tf_post_w"], [(128, 128), (64, 128)], s32)
    assert close(w[0], pair.nmp_mlp_end.layers[0].weight) and close(w[1], pair.nmp_mlp_end.layers[1].weight)
    assert "tf_hagg_w" not in t
    # fused pairwise aggregation: chunks in the issue order of the kernel's MMA warp (csrc/gn_pair_agg_tf32.cu)
    w = _decode(t["tf_pagg_w"], [(64, 64)] * 24, s32)
    order = [k for k, _ in packing.pair_agg_tf32_order(12)]
    assert len(order) == 24 and sorted(order) == sorted([("a", u) for u in range(12)] + [("b", u) for u in range(12)])
    assert order[:5] == [("a", 0), ("a", 1), ("a", 2), ("a", 3), ("b", 0)] and order[-3:] == [("b", 9), ("b", 10), ("b", 11)]
    for u in range(12):                        # every GEMM 1 precedes the GEMM 2 of its unit step
        assert order.index(("a", u)) < order.index(("b", u))
    for got, (kind, u) in zip(w, order):
        tt, half = divmod(u, 2)
        want = (agg[tt].layers[0].weight[64 * half:64 * half + 64, :] if kind == "a"
                else agg[tt].layers[1].weight[:, 64 * half:64 * half + 64])
        assert close(got, want), (kind, u)

    hyp = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=1, scale=5)
    t = packing.pack_stage(hyp, 0, torch.device("cpu"))
    agg = hyp.edge_aggregation_list[0].agg_mlp
    w = _decode(t["tf_hagg_w"], [(128, 64), (64, 128)] * 10, 65536)
    for i in range(10):
        assert close(w[2 * i], agg[i].layers[0].weight) and close(w[2 * i + 1], agg[i].layers[1].weight)
    assert "tf_aggin_w" not in t and "tf_aggout_w" not in t


def test_chunk_rule_matches_the_ring_stage():
    for n, k, want in [(128, 64, 16), (64, 128, 32), (16, 128, 128), (64, 64, 32), (128, 128, 16), (256, 64, 8),
                       (48, 128, 32), (96, 128, 16), (32, 128, 64)]:
        assert packing.tf_chunk_k(n, k, 16384) == want
    for n, k, want in [(128, 64, 64), (64, 128, 128), (16, 128, 128), (128, 128, 64)]:
        assert packing.tf_chunk_k(n, k, 65536) == want
