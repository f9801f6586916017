"""GPU timing aid: hb_linear_tf32x3 (tcgen05, 3 x TF32) on the classifier's product shapes, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from heybuddy_b200 import _native
lib = _native.load()
dev = torch.device("cuda", 0)
st = _native.stream_ptr(dev)
for m, n, k in ((4096, 128, 1536), (4096, 8192, 1536), (30000, 8192, 1536), (4096, 128, 96), (4096, 96, 64)):
    x, w, b = torch.randn((m, k), device=dev), torch.randn((n, k), device=dev), torch.randn(n, device=dev)
    y = torch.empty((m, n), device=dev)
    run = lambda: _native.check(lib.hb_linear_tf32x3(x.data_ptr(), k, w.data_ptr(), k, b.data_ptr(), y.data_ptr(), n, m, n, k, st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"[{m} x {k}] x [{n} x {k}]^T: {ms * 1e3:8.1f} us  {2 * m * n * k / ms / 1e9:7.1f} TFLOP/s (fp32-equivalent)")
