"""GPU profiling aid (build with HB_NVCC_EXTRA=-DHB_TC_FINE): MMA-issue / epilogue timeline inside layer 1 of each block."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heybuddy_b200 import _native
from heybuddy_b200.embeddings import SpeechEmbeddingModel
lib = _native.load()
lib.hb_debug_tc_fine.argtypes = [ctypes.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
model = SpeechEmbeddingModel(device_id=0, precision="f16", load=True)
mel = torch.randn((B, 141, 32), device="cuda") * 0.5 + 11
ws = torch.empty(lib.hb_embed_clips_workspace_bytes(B, 141, 1), dtype=torch.uint8, device="cuda")
out = torch.empty((B, 141, 32), device="cuda")
for layer, label in ((3, "block1"), (7, "block2"), (11, "block3"), (15, "block4")):
    for rep in range(2):
        lib.hb_embed_activation(model._handle, 1, mel.data_ptr(), B, 141, layer, out.data_ptr(), out.numel(), ws.data_ptr(), ws.numel(), None)
    torch.cuda.synchronize()
    t = np.zeros((8, 48), dtype=np.int64)
    lib.hb_debug_tc_fine(t.ctypes.data)
    print(label)
    for cta in (2, 5):
        r = t[cta]; t0 = r[0]
        mma = [int(x - t0) for x in r[0:8] if x]
        e2 = [int(x - t0) for x in r[16:24] if x]
        e9 = [int(x - t0) for x in r[32:40] if x]
        print("  cta", cta, "mma[enter,wready,(slot_ok,committed)*]", mma, "| epi w2 [(enter,full,done)*]", e2, "| epi w9", e9)
