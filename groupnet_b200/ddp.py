"""Data-parallel training step for the layers (BASELINE config 5).

Replicas + ONE all-reduce(sum) of a single flat fp32 gradient bucket per step over NCCL
(NVLink 5 / NVSwitch), then scale by 1/world.  The bucket is small (0.78 M floats = 3.1 MB for the
three PastEncoder layers at the NBA shape), so the collective is latency-bound: a single flat bucket
launched on the compute stream right after the last wgrad kernel, no overlap machinery.
Never-used parameters (`edge_aggregation.mlp`, `spatial_embedding`, `spatial_transform`: grad is
None) are zero-filled identically on every rank so the bucket layout is static.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch


class FlatGradBucket:
    """Static flat view over the gradients of `params` (registration order)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat: Optional[torch.Tensor] = None

    def _ensure(self, device):
        if self.flat is None or self.flat.device != device:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=device)

    def pack(self) -> torch.Tensor:
        dev = self.params[0].device
        self._ensure(dev)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()          # never-used parameter: identical zeros on all ranks
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        return self.flat

    def unpack(self) -> None:
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is not None:
                p.grad.copy_(self.flat[off:off + n].view_as(p.grad))
            off += n

    def zero_grad(self) -> None:
        """Arena mode: ONE memset, and every parameter's .grad becomes a view into the flat bucket.  The layers'
        backward then accumulates its weight gradients straight into the bucket (autograd.StageFunction) and
        `allreduce_mean` reduces it in place: no per-parameter copies around the collective (the pack / unpack pair
        was ~400 launches per step for one 3.1 MB all-reduce).  Never-used parameters hold zeros instead of None."""
        dev = self.params[0].device
        self._ensure(dev)
        self.flat.zero_()
        if not self._views_installed():
            off = 0
            for p in self.params:
                n = p.numel()
                p.grad = self.flat[off:off + n].view(p.shape)
                off += n

    def _views_installed(self) -> bool:
        if self.flat is None:
            return False
        off = 0
        for p in self.params:
            g = p.grad
            if g is None or g.data_ptr() != self.flat.data_ptr() + 4 * off or g.shape != p.shape:
                return False
            off += p.numel()
        return True

    def allreduce_mean(self, group=None) -> None:
        """grads <- mean over ranks of grads (one collective)."""
        import torch.distributed as dist
        arena = self._views_installed()
        flat = self.flat if arena else self.pack()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(dist.get_world_size(group))
        if not arena:
            self.unpack()
