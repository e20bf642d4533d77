"""Data-parallel equivalence of the training step (run under torchrun on >= 2 GPUs):
after the flat-bucket all-reduce every rank holds the gradient of the full global batch.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
      --master-port 29533 tests/dp_equivalence_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import groupnet_b200 as gb  # noqa: E402
from groupnet_b200.ddp import FlatGradBucket  # noqa: E402
from groupnet_b200.sharding import shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    total, n, d = 64 * world, 11, 64
    gen = torch.Generator().manual_seed(0)
    x_all = torch.randn(total, n, d, generator=gen)
    w_all = torch.randn(total, n, 256, generator=gen)
    noise_all = [torch.rand(total, 121, 6, generator=gen), torch.rand(total, 11, 10, generator=gen),
                 torch.rand(total, 1, 10, generator=gen)]

    def grads_of(lo, hi, scale):
        torch.manual_seed(1234)
        m = gb.MultiScaleInteraction(d, (5, 11)).to(dev).train()
        feat, _ = m(x_all[lo:hi].to(dev), noise=[u[lo:hi].to(dev) for u in noise_all])
        ((feat * w_all[lo:hi].to(dev)).sum() / scale).backward()
        return m

    lo, hi = shard_range(total, rank, world)
    m = grads_of(lo, hi, hi - lo)
    bucket = FlatGradBucket(m.parameters())
    bucket.allreduce_mean()
    ref = grads_of(0, total, total)                      # every rank also computes the global-batch gradient
    worst = 0.0
    for (name, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert (p.grad is None) == (q.grad is None), name
        if p.grad is not None:
            den = q.grad.abs().max().item() + 1e-12
            worst = max(worst, (p.grad - q.grad).abs().max().item() / den)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dp_equivalence world={world} worst_rel_err={t.item():.3e}")
    assert t.item() < 1e-4, t.item()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
