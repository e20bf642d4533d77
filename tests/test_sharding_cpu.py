"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the shard
arithmetic, the rank configuration of the layers and the gather helper."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from groupnet_b200 import sharding


def test_shard_ranges_cover_batch():
    for total in (0, 1, 7, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            r = sharding.shard_ranges(total, world)
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import groupnet_b200 as gb
        a, b = sharding.shard_range(total, rank, world)
        # every rank builds the same weights (replicated), and is pointed at its slice
        torch.manual_seed(1234)
        m = gb.MultiScaleInteraction(64, (5, 11))
        sharding.configure_layer_for_rank(m, total, rank, world, seed=7)
        assert all(l.rng == "philox" and l.scene_offset == a for l in m.layers())
        assert len({l.philox_seed for l in m.layers()}) == 3          # distinct streams per layer
        sd = torch.cat([p.detach().flatten() for p in m.parameters()])
        ref = sd.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(sd, ref), "replicated weights differ across ranks"
        # scenes come back in global order
        local = torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 3)
        full = sharding.gather_scenes(local, total)
        assert torch.equal(full[:, 0], torch.arange(total, dtype=torch.float32))
        # weak-scaling throughput reduction used by bench.py: MAX over ranks of the step time
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 1001), nprocs=2, join=True)


def _ddp_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import groupnet_b200 as gb
        from groupnet_b200.ddp import FlatGradBucket
        torch.manual_seed(5)
        m = gb.MS_HGNN_hyper(64, 64, 64, 64, batch_norm=0, nmp_layers=1, scale=3)
        used = [p for n, p in m.named_parameters() if not n.startswith(("spatial_", "edge_aggregation_list.0.mlp"))]
        for i, p in enumerate(used):                      # fake per-rank grads; unused params keep grad None
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        bucket = FlatGradBucket(m.parameters())
        assert bucket.numel == sum(p.numel() for p in m.parameters()) == 283340
        bucket.allreduce_mean()
        mean = sum(range(1, world + 1)) / world
        for i, p in enumerate(used):
            assert torch.allclose(p.grad, torch.full_like(p, mean * (i + 1)))
        for n, p in m.named_parameters():
            if n.startswith("spatial_"):
                assert p.grad is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_flat_bucket_allreduce():
    port = _free_port()
    mp.spawn(_ddp_worker, args=(2, port), nprocs=2, join=True)


def _decoder_worker(rank, world, port, scenes):
    """The decoder shards like the layers: contiguous scene ranges, rows = scene x agent x sample, no collective."""
    import types
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import groupnet_b200 as gb
        from oracle import decoder_oracle as DO
        n, s, tp, tf = 11, 4, 5, 10
        torch.manual_seed(77)
        dec = gb.Decoder(types.SimpleNamespace(hidden_dim=64, hyper_scales=[5, 11], zdim=32, past_length=tp,
                                               future_length=tf, num_decompose=2))
        sd = {k: v.detach() for k, v in dec.state_dict().items()}
        gen = torch.Generator().manual_seed(5)
        a_all = scenes * n
        pf = torch.randn(a_all, 256, generator=gen).repeat_interleave(s, dim=0)
        z = torch.randn(a_all * s, 32, generator=gen)
        past, cur = torch.randn(a_all, tp, 2, generator=gen), torch.randn(a_all, 1, 2, generator=gen)
        a, b = sharding.shard_range(scenes, rank, world)
        with torch.no_grad():
            out, rec = DO.decoder_forward(sd, pf[a * n * s:b * n * s], z[a * n * s:b * n * s], b - a, n,
                                          past[a * n:b * n], cur[a * n:b * n], s, past_len=tp, future_len=tf,
                                          num_decompose=2, mode="inference")
            full_out = sharding.gather_scenes(out.reshape(b - a, n * s * tf * 2), scenes)
            full_rec = sharding.gather_scenes(rec.reshape(b - a, n * s * tp * 2), scenes)
            ref_out, ref_rec = DO.decoder_forward(sd, pf, z, scenes, n, past, cur, s, past_len=tp, future_len=tf,
                                                  num_decompose=2, mode="inference")
        assert torch.allclose(full_out, ref_out.reshape(scenes, -1), rtol=0, atol=2e-6)
        assert torch.allclose(full_rec, ref_rec.reshape(scenes, -1), rtol=0, atol=2e-6)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_decoder_sharding():
    port = _free_port()
    mp.spawn(_decoder_worker, args=(2, port, 7), nprocs=2, join=True)
