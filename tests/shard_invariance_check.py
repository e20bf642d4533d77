"""Sharding invariance of the forward (run under torchrun on >= 2 GPUs): rank r runs scenes [lo, hi) of a global
batch with scene_offset = lo; the gathered result must equal rank 0's single-GPU run of the whole batch bit for bit
(Philox noise keyed by the global scene index; no collective on the forward path)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import groupnet_b200 as gb  # noqa: E402
from groupnet_b200.sharding import shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    total, n, d = 999, 11, 64
    x_all = torch.randn(total, n, d, generator=torch.Generator().manual_seed(0))
    ok = True
    for precision in ("fp32", "tf32", "bf16"):
        torch.manual_seed(1234)
        m = gb.MultiScaleInteraction(d, (5, 11)).to(dev).eval().set_precision(precision)
        lo, hi = shard_range(total, rank, world)
        with torch.no_grad():
            m.set_rng("philox", seed=3, scene_offset=lo)
            feat, hcat = m(x_all[lo:hi].to(dev))
            parts_f = [torch.empty(shard_range(total, r, world)[1] - shard_range(total, r, world)[0], n, feat.shape[2],
                                   device=dev) for r in range(world)]
            parts_h = [torch.empty(p.shape[0], hcat.shape[1], n, device=dev) for p in parts_f]
            dist.all_gather(parts_f, feat.contiguous())       # test plumbing only: the product path has no collective
            dist.all_gather(parts_h, hcat.contiguous())
            if rank == 0:
                m.set_rng("philox", seed=3, scene_offset=0)
                f1, h1 = m(x_all.to(dev))
                ok = ok and torch.equal(torch.cat(parts_f), f1) and torch.equal(torch.cat(parts_h), h1)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0:
        print(f"shard_invariance world={world} {'ok' if ok else 'MISMATCH'}")
    dist.destroy_process_group()
    assert flag.item() == 1


if __name__ == "__main__":
    main()
