"""GPU profiling aid: per-phase cycle counts of the grouped tcgen05 block kernels (first CTAs of each launch)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heybuddy_b200 import _native
from heybuddy_b200.embeddings import SpeechEmbeddingModel
lib = _native.load()
lib.hb_debug_tcg_times.argtypes = [ctypes.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
model = SpeechEmbeddingModel(device_id=0, precision="f16", load=True)
mel = torch.randn((B, 141, 32), device="cuda") * 0.5 + 11
ws = torch.empty(lib.hb_embed_clips_workspace_bytes(B, 141, 1), dtype=torch.uint8, device="cuda")
out = torch.empty((B, 141, 32), device="cuda")
for layer, label, nl in ((3, "block1", 4), (7, "block2", 4), (11, "block3", 4), (15, "block4", 4)):
    for rep in range(2):
        lib.hb_embed_activation(model._handle, 1, mel.data_ptr(), B, 141, layer, out.data_ptr(), out.numel(), ws.data_ptr(), ws.numel(), None)
    torch.cuda.synchronize()
    t = np.zeros((8, 16), dtype=np.int64)
    lib.hb_debug_tcg_times(t.ctypes.data)
    d = np.diff(t[:, :9], axis=1).mean(axis=0).astype(int)
    names = ["(first tile)", "stage"] + [f"L{i}" for i in range(nl)]
    print(label, "kernel total", int((t[:, 8] - t[:, 0]).mean()), "| second tile:", dict(zip(names[1:], d[1:2 + nl])), "store", int((t[:, 7] - t[:, 2 + nl]).mean()),
          "tile total", int((t[:, 7] - t[:, 1]).mean()))
    f = (t[:, 9:15] - t[:, 9:10]).mean(axis=0).astype(int)
    print("    layer 1 timeline: enter 0, weights ready", f[1], "MMAs issued", f[2], "| epilogue w2: full seen", f[3], "done", f[4], "| w9 done", f[5])
