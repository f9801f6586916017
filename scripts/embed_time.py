"""GPU timing aid: hb_embed_clips alone (CUDA events), B clips of 141 frames."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from heybuddy_b200 import spec
from heybuddy_b200.embeddings import SpeechEmbeddingModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
model = SpeechEmbeddingModel(device_id=0, precision="f16", load=True)
OFFS = spec.embedding_frame_offsets(spec.CLIP_SAMPLES)
mels = [torch.randn((B, 141, 32), device="cuda") * 0.5 + 11 for _ in range(3)]
for i in range(3):
    model.run_clips_device(mels[i % 3], OFFS)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    model.run_clips_device(mels[i % 3], OFFS)
e1.record()
torch.cuda.synchronize()
print("embed ms per", B, "clips:", round(e0.elapsed_time(e1) / 10, 3))
