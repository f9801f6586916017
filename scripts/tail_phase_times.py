"""GPU profiling aid: phase cycle counts of the last tail layer kernel (conv2d_19) of an 8192-clip embed call, first 8 CTAs."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from heybuddy_b200 import _native, spec
from heybuddy_b200.embeddings import SpeechEmbeddingModel
lib = _native.load()
lib.hb_debug_tail_times.argtypes = [ctypes.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
model = SpeechEmbeddingModel(device_id=0, precision="f16", load=True)
mel = torch.randn((B, 141, 32), device="cuda") * 0.5 + 11
offs = np.asarray(spec.embedding_frame_offsets(spec.CLIP_SAMPLES), dtype=np.int32)
for _ in range(3):
    out = model.run_clips_device(mel, offs)
torch.cuda.synchronize()
t = np.zeros((8, 16), dtype=np.int64)
lib.hb_debug_tail_times(t.ctypes.data)
rel = t[:, :10] - t[:, :1]
names = ["start", "setup", "t2 staged", "t2 issued", "t2 ready", "t2 drained", "all tiles", "dealloc", "(tiles)", "t1 done"]
for i, n in enumerate(names):
    print(f"{n:12s}", (t[:, 8] if i == 8 else rel[:, i]).tolist())
