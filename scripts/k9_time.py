"""GPU timing aid: the four K9 transforms alone at the reference's default probabilities (0.25 each) on 8192 clips, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heybuddy_b200 import _native
from heybuddy_b200.dataset import k9
from heybuddy_b200.dataset.draws import AugmentConfig, DrawTable
n, T = 8192, 23040
x = (torch.randn((n, T), device="cuda") * 0.1)
if len(sys.argv) > 1:      # zero padding like length-fixed clips: the first and the last `pad` samples of every clip
    pad = int(sys.argv[1])
    x[:, :pad] = 0
    x[:, T - pad:] = 0
bufs = {}
def scratch(name, numel, dtype):
    if name not in bufs or bufs[name].numel() < numel or bufs[name].dtype != dtype:
        bufs[name] = torch.empty(int(numel), dtype=dtype, device="cuda")
    return bufs[name]
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for name, kw in (("eq", dict(seven_band_prob=0.25)), ("tanh", dict(tanh_distortion_prob=0.25)), ("pitch", dict(pitch_shift_prob=0.25)),
                 ("bandstop", dict(band_stop_prob=0.25)), ("all four", dict(seven_band_prob=0.25, tanh_distortion_prob=0.25, pitch_shift_prob=0.25, band_stop_prob=0.25))):
    table = DrawTable.build(np.full(n, 20000), AugmentConfig(batch_size=128, **kw), 3)
    host = table.k9.pack()
    pk = {k_: (v if k_ in k9.HOST_ONLY else torch.from_numpy(np.ascontiguousarray(v)).cuda()) for k_, v in host.items()}
    counts = {k_: int(v.shape[0]) for k_, v in host.items() if k_.endswith("_idx")}
    print(f"{name:9s} {t(lambda: k9.apply_packed(x, pk, scratch)):8.3f} ms   clips {counts}  band-stop rows {host['bs_taps'].shape[0]}")
