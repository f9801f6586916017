"""GPU: where does the host-path (e2e) step time go?  cProfile over FeaturizePipeline.featurize_host."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from heybuddy_b200.dataset.augmented import AugmentedAudioGenerator
from heybuddy_b200.dataset.draws import DrawTable
from heybuddy_b200.embeddings import SpeechEmbeddings
from heybuddy_b200.pipeline import FeaturizePipeline
dev = torch.device("cuda:0")
noise, rir = bench.make_banks(dev)
speech = SpeechEmbeddings(device_id=0, precision="f16")
aug = AugmentedAudioGenerator([], device_id=0, augmentation_dataset=noise, impulse_response_dataset=rir, batch_size=128,
                              colored_noise_min_f_decay=0.0, colored_noise_max_f_decay=0.0, seed=2004)
pipe = FeaturizePipeline(aug, speech, device_id=0)
clips = bench.make_sources(8192, 2001, dev).pin()
table = aug.next_table(clips.lengths)
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
bps = sub // 128
parts = []
for lo in range(0, len(table.batches), bps):
    d = DrawTable(cfg=table.cfg, seed=table.seed)
    d.batches, d.noise_clip_cursor, d.rir_index = table.batches[lo:lo + bps], table.noise_clip_cursor[lo:lo + bps], table.rir_index[lo:lo + bps]
    parts.append(d)
out = torch.empty((8192, 16, 96), dtype=torch.float32).pin_memory()
for _ in range(2):
    pipe.featurize_host(clips, parts, sub, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    pipe.featurize_host(clips, parts, sub, out=out)
torch.cuda.synchronize()
print("ms per step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    pipe.featurize_host(clips, parts, sub, out=out)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
