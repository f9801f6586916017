"""GPU profiling aid: per-phase cycle counts of the tcgen05 block kernels (first CTAs of each launch)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heybuddy_b200 import _native, spec
from heybuddy_b200.embeddings import SpeechEmbeddingModel
lib = _native.load()
lib.hb_debug_tc_times.argtypes = [ctypes.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
model = SpeechEmbeddingModel(device_id=0, precision="f16", load=True)
mel = torch.randn((B, 141, 32), device="cuda") * 0.5 + 11
ws_b = lib.hb_embed_clips_workspace_bytes(B, 141, 1)
ws = torch.empty(ws_b, dtype=torch.uint8, device="cuda")
out = torch.empty((B, 141, 32), device="cuda")
names = ["setup", "stage", "L0", "L0post", "L1", "L1post", "L2", "L2post", "L3", "L3post", "store", "dealloc"]
for layer, label in ((3, "block1"), (7, "block2"), (11, "block3"), (15, "block4")):
    for rep in range(2):
        n = lib.hb_embed_activation(model._handle, 1, mel.data_ptr(), B, 141, layer, out.data_ptr(), out.numel(), ws.data_ptr(), ws.numel(), None)
    torch.cuda.synchronize()
    t = np.zeros((8, 16), dtype=np.int64)
    lib.hb_debug_tc_times(t.ctypes.data)
    d = np.diff(t[:, :13], axis=1)
    nl = 3 if label == "block1" else 4
    print(label, "total", (t[:, 12] - t[:, 0]).mean())
    cols = [0, 1] + list(range(2, 2 + 2 * nl))
    print("   ", {names[c]: int(d[:, c].mean()) for c in cols}, "store", int((t[:, 11] - t[:, 2 + 2 * nl]).mean()), "dealloc", int(d[:, 11].mean()))
