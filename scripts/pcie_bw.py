"""GPU box aid: pinned host <-> device copy bandwidth (what bounds bench.py's e2e leg)."""
import torch
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
print("H2D GB/s", round(t(lambda: d.copy_(h, non_blocking=True)), 1), "D2H GB/s", round(t(lambda: h.copy_(d, non_blocking=True)), 1))
s2 = torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
def both():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5): both()
torch.cuda.synchronize()
print("H2D + D2H concurrently, GB/s each direction", round(n * 5 / (time.perf_counter() - t0) / 1e9, 1))
