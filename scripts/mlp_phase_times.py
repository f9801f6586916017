"""Phase times of the classifier's forward tail kernel (clock64 stamps of CTA 0 / thread 0; hb_debug_mlp_stamps)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from heybuddy_b200 import _native
from heybuddy_b200.wakeword import WakeWordMLPModel
B = 4096
m = WakeWordMLPModel(device_id=0, seed=5)
g = torch.Generator().manual_seed(1)
y = (torch.rand(B, generator=g) < 0.09).to(torch.int64).cuda()
x = torch.randn(B, 16, 96, generator=g).cuda()
for _ in range(5):
    m.train_step(x, y, 1e-3, 0.7, 1e-4)
lib = _native.load()
st = torch.zeros(32, dtype=torch.int64, device="cuda")
lib.hb_debug_mlp_stamps.argtypes = [ctypes.c_void_p]
lib.hb_debug_mlp_stamps(ctypes.c_void_p(st.data_ptr()))
m.train_step(x, y, 1e-3, 0.7, 1e-4)
torch.cuda.synchronize()
lib.hb_debug_mlp_stamps(ctypes.c_void_p(0))
v = st.cpu().tolist()
names = ["start"]
for s in range(4):
    if s > 0:
        names += [f"s{s} acquire Whg", f"s{s} hg product"]
    else:
        names += ["s0 slice sums"]
    if s < 3:
        names += [f"s{s} gate+stores", f"s{s} acquire Wo", f"s{s} o product", f"s{s} LN"]
names += ["s3 gate + logit"]
prev = v[0]
for n, t in zip(names[1:], v[1:]):
    if t == 0:
        break
    print(f"{n:20s} {t - prev:7d} clk")
    prev = t
print("total", prev - v[0])
