"""e2e leg probe: generate_sharded with different sinks / thread counts, host-side time breakdown."""
import os, sys, time, tempfile, shutil
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("HEYBUDDY_B200_ALLOW_RANDOM_INIT", "1")
import bench
from heybuddy_b200.dataset.features import RaggedClipSource, TrainingFeaturesGenerator
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
noise, rir = bench.make_banks(dev)
pool = bench.make_sources(3 * 8192, 2001, dev).pin()
rows = 8192 * 6
for outdir, nopin in ((tempfile.gettempdir(), "0"), ("/dev/shm", "0"), ("/dev/shm", "1")):
    os.environ["HEYBUDDY_B200_NO_PINNED_FILE"] = nopin
    for sub in (4096,):
        for threads in (2, 4, 8):
            d = tempfile.mkdtemp(prefix="hb_probe_", dir=outdir)
            gen = TrainingFeaturesGenerator(device_id=0, use_autoconfigure=False, augment_batch_size=128, augment_background_dataset=noise,
                                            augment_impulse_dataset=rir, augment_colored_noise_min_f_decay=0.0, augment_colored_noise_max_f_decay=0.0,
                                            chunk_clips=sub, source=RaggedClipSource(pool), sample_batch_size=3 * 8192, seed=2004)
            gen.generate_sharded(8192 * 2, os.path.join(d, "w.npy"), writer_threads=threads)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            gen.generate_sharded(rows, os.path.join(d, "x.npy"), writer_threads=threads)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            pipe = gen._pipe[1]
            print(f"{outdir} nopin={nopin} sink={gen.last_sink[:12]} sub={sub} threads={threads}: {dt / 6 * 1e3:.2f} ms/step  stats={ {k: round(v, 4) for k, v in pipe.last_stats.items()} }", flush=True)
            shutil.rmtree(d)
# pinned-memory sink (no file) for comparison
out = torch.empty((rows, 16, 96), dtype=torch.float32).pin_memory()
gen._run(0, rows, lambda lo, hi: out[lo:hi], False, False)
torch.cuda.synchronize(); t0 = time.perf_counter()
gen._run(0, rows, lambda lo, hi: out[lo:hi], False, False)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pinned sink: {dt / 6 * 1e3:.2f} ms/step stats={ {k: round(v, 4) for k, v in gen._pipe[1].last_stats.items()} }")
