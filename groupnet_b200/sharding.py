"""Batch sharding of the forward path over the GPUs of one box.

Scenes are independent (every op in model/MS_HGNN_batch.py is batched over dim
0 and there is no cross-scene reduction), so the path shards by contiguous
ranges of scenes with NO data-path collective: rank r of W owns scenes
[offset(r), offset(r+1)).  Device Philox noise is keyed by the GLOBAL scene
index (`set_rng(..., scene_offset=offset(r))`), so per-scene results do not
depend on W.  The helpers below are pure host logic (testable with gloo).
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first `total % world` ranks get one extra scene."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_ranges(total: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(total, r, world) for r in range(world)]


def configure_layer_for_rank(layer, total: int, rank: int, world: int, seed: int = 0):
    """Point a layer (or MultiScaleInteraction) at this rank's slice of the global batch."""
    start, _ = shard_range(total, rank, world)
    layer.set_rng("philox", seed=seed, scene_offset=start)
    return layer


def gather_scenes(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather per-rank scene slices back into the global batch order (evaluation /
    tests only: the forward path itself needs no collective)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ranges = shard_ranges(total, world)
    assert local.shape[0] == ranges[rank][1] - ranges[rank][0]
    pad = max(b - a for a, b in ranges)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:b - a] for o, (a, b) in zip(out, ranges)], dim=0)
