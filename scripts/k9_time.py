"""GPU timing aid: the K9 kernels alone (hb_k9_eq_f32 / hb_k9_tanh_f32) on 2048 of 8192 clips, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heybuddy_b200 import _native
from heybuddy_b200.dataset.draws import AugmentConfig, DrawTable
lib = _native.load()
n, T = 8192, 23040
cfg = AugmentConfig(batch_size=128, seven_band_prob=0.25, tanh_distortion_prob=0.25)
table = DrawTable.build(np.full(n, 20000), cfg, 3)
eq_idx, sos, th_idx, amt = table.k9.pack()
x = (torch.randn((n, T), device="cuda") * 0.1)
d = [torch.from_numpy(a).cuda() for a in (eq_idx, sos, th_idx, amt)]
st = _native.stream_ptr(x.device)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("eq   clips", eq_idx.size, "ms", round(t(lambda: _native.check(lib.hb_k9_eq_f32(x.data_ptr(), d[0].data_ptr(), d[1].data_ptr(), int(eq_idx.size), T, st))), 3))
print("tanh clips", th_idx.size, "ms", round(t(lambda: _native.check(lib.hb_k9_tanh_f32(x.data_ptr(), d[2].data_ptr(), d[3].data_ptr(), int(th_idx.size), T, st))), 3))
