"""GPU profiling aid: cycle stamps of the fused augment + mel kernel's phases (8 CTAs of the ninth wave) on the bench's chunk."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from heybuddy_b200 import _native
from heybuddy_b200.dataset.augmented import AugmentedAudioGenerator
from heybuddy_b200.embeddings import SpeechEmbeddings
from heybuddy_b200.pipeline import FeaturizePipeline
dev = torch.device("cuda", 0)
lib = _native.load()
lib.hb_debug_aug_times.argtypes = [ctypes.c_void_p]
noise, rir = bench.make_banks(dev)
aug = AugmentedAudioGenerator([], device_id=0, augmentation_dataset=noise, impulse_response_dataset=rir, batch_size=128,
                              colored_noise_min_f_decay=0.0, colored_noise_max_f_decay=0.0, seed=2004)
pipe = FeaturizePipeline(aug, SpeechEmbeddings(device_id=0, precision="f16"), device_id=0)
clips = bench.make_sources(8192, 2001, dev)
table = aug.next_table(clips.lengths)
chunk = pipe.upload(clips, table)
for _ in range(3):
    pipe.run_fused_front(chunk)
torch.cuda.synchronize()
t = np.zeros((8, 8), dtype=np.int64)
lib.hb_debug_aug_times(t.ctypes.data)
params = aug.clip_params(table)
names = ["start", "loaded", "K1-K3 done", "abs sum", "FFT done", "scaled+sync", "mel done"]
for k in range(8):
    i = 1184 + k
    rel = (t[k, :7] - t[k, 0]).tolist()
    print(f"clip {i}: reverb {int(params['rir_index'][i]) >= 0} noise {int(params['noise_offset'][i]) >= 0} coloured {int(params['colored_index'][i]) >= 0}", dict(zip(names, rel)))
