"""
GPU check (torchrun, NCCL): data-parallel classifier training over N ranks vs the same global batches on one device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/dp_train_check.py

Every rank trains a replica on its shard through heybuddy_b200.dp.distributed_train_step; rank 0 also trains a
single-device model on the whole batch.  Prints the largest parameter difference, the cross-rank replica checksum and
the step time of both.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from heybuddy_b200.dp import distributed_train_step, shard_batch
from heybuddy_b200.wakeword import WakeWordMLPModel


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    steps, batch = int(os.environ.get("DP_STEPS", "50")), int(os.environ.get("DP_BATCH", "4096"))
    dp = WakeWordMLPModel(device_id=local, seed=5)
    single = WakeWordMLPModel(device_id=local, seed=5) if rank == 0 else None
    g = torch.Generator(device="cpu").manual_seed(4001)
    direction = torch.randn(1, 1, 96, generator=g)
    t_dp = t_one = 0.0
    for step in range(steps):
        y = (torch.rand(batch, generator=g) < 0.09).to(torch.int64)           # ~ the 50:50:1000 mix of the reference (positives rare)
        x = torch.randn(batch, 16, 96, generator=g) + 0.5 * direction * y[:, None, None]
        xd, yd = x.to(dev), y.to(dev)
        xs, ys = shard_batch(xd, yd, rank, world)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        _, stats = distributed_train_step(dp, xs.contiguous(), ys.contiguous(), 1e-3, 0.7, 1e-4, min_selected=128)
        torch.cuda.synchronize()
        t_dp += time.perf_counter() - t0
        if rank == 0:
            t0 = time.perf_counter()
            _, s1 = single.train_step(xd, yd, 1e-3, 0.7, 1e-4, min_selected=128)
            torch.cuda.synchronize()
            t_one += time.perf_counter() - t0
            assert float(stats[1]) == float(s1[1]), (step, stats, s1)
    flat = torch.cat([v.reshape(-1) for v in dp.state_dict().values()]).to(dev)
    checksum = torch.stack([flat.double().sum(), flat.double().abs().sum()])
    gathered = [torch.zeros_like(checksum) for _ in range(world)]
    dist.all_gather(gathered, checksum)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        ref = torch.cat([v.reshape(-1) for v in single.state_dict().values()]).to(dev)
        d = (flat - ref).abs()
        diff, close = float(d.max()), float((d <= 2e-5).float().mean())
        print(f"DP_CHECK world={world} steps={steps} batch={batch} replicas_identical={same} max|param - single|={diff:.3e} within_2e-5={close:.5f} "
              f"(max|param|={float(ref.abs().max()):.3f}) ms/step dp={1e3 * t_dp / steps:.3f} single={1e3 * t_one / steps:.3f}", flush=True)
        assert same and close >= 0.999 and diff <= 2 * steps * 1e-3   # Adam-normalised updates: see tests/test_dp_training.py
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
