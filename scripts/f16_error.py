"""GPU: error of the f16 (tcgen05) embedding mode against the fp32 parity mode on the same mel input."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from heybuddy_b200 import spec
from heybuddy_b200.embeddings import SpeechEmbeddingModel
g = torch.Generator().manual_seed(3)
mel = (torch.randn((256, 141, 32), generator=g) * 0.6 + 1.0).cuda()
offs = spec.embedding_frame_offsets(spec.CLIP_SAMPLES)
a = SpeechEmbeddingModel(device_id=0, precision="f16", load=True).run_clips_device(mel, offs).double()
b = SpeechEmbeddingModel(device_id=0, precision="fp32", load=True).run_clips_device(mel, offs).double()
print("rel L2 %.3e  max|err|/max|ref| %.3e  per-clip worst rel L2 %.3e" % (
    float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max()),
    float(((a - b).flatten(1).norm(dim=1) / b.flatten(1).norm(dim=1)).max())))
