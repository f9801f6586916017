"""GPU box aid (torchrun, one rank per GPU): pinned host -> device bandwidth with ALL ranks copying at the same time."""
import os, time, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d.copy_(h, non_blocking=True); torch.cuda.synchronize()
res = []
for together in (False, True):
    for r in range(world if not together else 1):
        dist.barrier(); torch.cuda.synchronize()
        if together or r == rank:
            t0 = time.perf_counter()
            for _ in range(8): d.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            res.append(n * 8 / (time.perf_counter() - t0) / 1e9)
        dist.barrier()
t = torch.tensor(res, device="cuda", dtype=torch.float64)
out = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(out, t)
if rank == 0:
    print("H2D GB/s alone   :", [round(float(o[0]), 1) for o in out])
    print("H2D GB/s together:", [round(float(o[1]), 1) for o in out], "sum", round(sum(float(o[1]) for o in out), 1), "cores", os.cpu_count())
dist.destroy_process_group()
