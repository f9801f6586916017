"""GPU timing aid: fused classifier training step and forward (CUDA events), batch 4096 (BASELINE config 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from heybuddy_b200.wakeword import WakeWordMLPModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = WakeWordMLPModel(device_id=0, seed=5)
g = torch.Generator().manual_seed(1)
y = (torch.rand(B, generator=g) < 0.09).to(torch.int64).cuda()
x = (torch.randn(B, 16, 96, generator=g)).cuda() + 0.5 * y[:, None, None]
for _ in range(5):
    m.train_step(x, y, 1e-3, 0.7, 1e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    m.train_step(x, y, 1e-3, 0.7, 1e-4)
e1.record(); torch.cuda.synchronize()
print("train step ms:", round(e0.elapsed_time(e1) / 50, 3))
e0.record()
for _ in range(50):
    m(x)
e1.record(); torch.cuda.synchronize()
print("forward ms:", round(e0.elapsed_time(e1) / 50, 3))
