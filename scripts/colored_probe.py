import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from heybuddy_b200 import _native
from heybuddy_b200.dataset.draws import colored_noise_base, gaussian_pattern
lib = _native.load()
for seed in (2004, 1 << 32, (1 << 32) + 2004, 0x123456789ABCDEF0):
    for g in (0, 7, (1 << 33) + 5):
        for fd in (0.0, 1.0):
            ids = torch.tensor([g], dtype=torch.int64).cuda(); f = torch.tensor([fd], dtype=torch.float32).cuda()
            out = torch.empty((1, 16000), dtype=torch.float32, device="cuda")
            _native.check(lib.hb_colored_bases(seed, ids.data_ptr(), f.data_ptr(), 1, out.data_ptr(), _native.stream_ptr(out.device)))
            got = out.cpu().numpy()[0]
            want = colored_noise_base(gaussian_pattern(seed, g), fd)
            print(hex(seed), g, fd, "maxdiff", float(np.abs(got - want).max()), got[:4], want[:4])
