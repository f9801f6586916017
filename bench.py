#!/usr/bin/env python
"""
bench.py -- clip-seconds featurized per second (augmentation + log-mel + speech embeddings).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): synthetic int16 source clips of ragged
length U[6400, 22400] (band-limited noise under a raised-cosine envelope, peak 32767), noise bank
2048 x 160000 f32 consumed as one contiguous stream, RIR bank 271 x L~U[3200, 24000], augmentation
batch 128 with probabilities coloured 0.25 (f_decay 0 = white, SNR U[10,30]) / gain 1.0 /
background 0.75 (SNR U[-10,15]) / reverb 0.75, random-init embedding weights (seed 3001).
One *step* = one pass of the hot path over one chunk of CHUNK clips per GPU (64 augmentation
batches): length fix + augmentation (one kernel) -> mel -> embedding conv stack -> [CHUNK,16,96] f32, one
hb_featurize_i16 call per step.
K steps x CHUNK clips ~ the 100k clips of the config at the default K.

  value   whole-job clip-seconds per second with the step's inputs already resident in HBM.
  e2e     the same through the public host API (FeaturizePipeline.featurize_stream, the streaming form of
          featurize_host): pinned host int16 clips -> H2D -> pipeline -> D2H of the embeddings, every step's copies
          inside the timed region.
  roofline  the dominant stage (embedding conv stack): algorithmic FLOPs / CUDA-event time vs the measured
          tensor peak in MEASURED_PEAKS.json.
  cpu_baseline  the oracle's restatement of the reference pipeline (reference control flow: 4 overlapping mel
          windows and 16 embedding windows per clip, batch 32) timed on the host cores on a bounded sample.

Each rank uses a pool of distinct chunks (> L2) and cycles through them, so no step re-reads inputs
that are still in the 126 MB L2.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = int(os.environ.get("HB_BENCH_CHUNK", "8192"))       # clips per GPU per step
E2E_SUB = int(os.environ.get("HB_BENCH_E2E_SUB", "4096"))   # e2e: sub-chunk pipelined H2D / compute / D2H inside a step
POOL = int(os.environ.get("HB_BENCH_POOL", "3"))            # distinct chunks cycled through
AUG_BATCH = 128
NOISE_CLIPS, NOISE_LEN = int(os.environ.get("HB_BENCH_NOISE_CLIPS", "2048")), 160000
N_RIRS = 271
CLIP_SECONDS = 1.44
METRIC = "clip-sec featurized/sec (aug+mel+embed)"
UNIT = "clip-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("HEYBUDDY_B200_EMBED_PRECISION", None))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(precision):
    return {
        "workload": "feature generator with augmentation (BASELINE configs[1]): ragged int16 clips U[6400,22400] -> 1.44 s, "
                    "coloured(white) 0.25 / gain 1.0 / background 0.75 / reverb 0.75, aug batch 128, random-init embedding weights",
        "clips_per_step_per_gpu": CHUNK, "aug_batch": AUG_BATCH, "noise_bank": [NOISE_CLIPS, NOISE_LEN], "rirs": N_RIRS,
        "embed_precision": precision, "l2_policy": f"{POOL} distinct input chunks cycled (> L2), noise bank 1.3 GB streamed",
    }


# --------------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------------
def make_sources(n, seed, device):
    """Ragged int16 clips generated on the GPU (input synthesis, outside any timed region)."""
    import torch

    from heybuddy_b200.pipeline import RaggedClips

    g = torch.Generator(device=device).manual_seed(seed)
    lengths = torch.randint(6400, 22400, (n,), generator=g, device=device)
    parts = []
    for lo in range(0, n, 2048):
        ln = lengths[lo:lo + 2048]
        x = torch.randn((ln.numel(), 1, 22400 + 7), generator=g, device=device)
        x = torch.nn.functional.avg_pool1d(x, 8, 1)[:, 0, :22400]                     # band limit
        i = torch.arange(22400, device=device)[None, :]
        env = 0.5 - 0.5 * torch.cos(2 * np.pi * i / ln[:, None])                          # raised cosine over the clip
        x = x * env * (i < ln[:, None])
        x = x / x.abs().amax(dim=1, keepdim=True) * 32767.0
        parts.append((x.to(torch.int16), i < ln[:, None]))
    samples = torch.cat([x[m] for x, m in parts]).cpu().numpy()
    lengths = lengths.cpu().numpy().astype(np.int64)
    return RaggedClips(samples, np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64))


def make_banks(device, seed_noise=2002, seed_rir=2003):
    import torch

    from heybuddy_b200.dataset.augmented import NoiseBank, RirBank

    g = torch.Generator(device=device).manual_seed(seed_noise)
    margin = AUG_BATCH * 23040
    noise = NoiseBank.__new__(NoiseBank)
    stream = torch.empty(NOISE_CLIPS * NOISE_LEN + margin, dtype=torch.float32, device=device)
    stream[:NOISE_CLIPS * NOISE_LEN].normal_(0.0, 0.1, generator=g)
    stream[NOISE_CLIPS * NOISE_LEN:] = stream[:margin]
    noise.stream = stream
    noise.clip_lengths = np.full(NOISE_CLIPS, NOISE_LEN, dtype=np.int64)
    noise.clip_starts = np.arange(NOISE_CLIPS + 1, dtype=np.int64) * NOISE_LEN
    noise.num_samples = NOISE_CLIPS * NOISE_LEN
    rng = np.random.Generator(np.random.PCG64(seed_rir))
    rirs = []
    for _ in range(N_RIRS):
        ln = int(rng.integers(3200, 24000))
        r = np.exp(-np.arange(ln) / rng.uniform(300.0, 3000.0)) * rng.standard_normal(ln)
        r[int(rng.integers(0, 200))] = 4.0
        rirs.append(r.astype(np.float32))
    return noise, RirBank(rirs, device)


# --------------------------------------------------------------------------------------------------
# clocks sampler (nvml; the recipe's clocks line)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self._stop = index, [], set(), threading.Event()
        self.max_mhz = None
        self.thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # pragma: no cover
            self.nv = None
            self.err = str(exc)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# CPU reference arm / baseline (oracle restatement of the reference pipeline)
# --------------------------------------------------------------------------------------------------
def cpu_reference_clips_per_s(n_clips, threads, seed=77, steps=1, warmup=0):
    """
    The reference's control flow with the oracle's CPU kernels behind it (ORT / torch_audiomentations /
    speechbrain are not installable offline): augment batch of 8 (constants.py:118) -> 4 x
    17280-sample mel windows -> 16 x 76-frame embedding windows per clip, spectrogram/embedding batch 32.
    Returns (clip_seconds_per_second, seconds_per_step).
    """
    import torch

    from heybuddy_b200 import spec
    from heybuddy_b200.dataset.draws import AugmentConfig, draw_batch
    from oracle import augment as oaug, embed as oembed, mel as omel, pipeline as opipe

    torch.set_num_threads(threads)
    rng = np.random.Generator(np.random.PCG64(seed))
    weights = spec.init_embedding_weights()
    cfg = AugmentConfig(batch_size=8, colored_noise_min_f_decay=0.0, colored_noise_max_f_decay=0.0)
    noise = (0.1 * rng.standard_normal((8 * spec.CLIP_SAMPLES * 2,))).astype(np.float32)
    rirs = []
    for _ in range(8):
        ln = int(rng.integers(3200, 24000))
        r = np.exp(-np.arange(ln) / rng.uniform(300.0, 3000.0)) * rng.standard_normal(ln)
        r[int(rng.integers(0, 200))] = 4.0
        rirs.append(r.astype(np.float32))
    sources = []
    for _ in range(n_clips):
        ln = int(rng.integers(6400, 22400))
        x = np.convolve(rng.standard_normal(ln), np.ones(8) / 8, mode="same") * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(ln) / ln))
        sources.append((x / np.abs(x).max() * 32767).astype(np.int16))

    def one_pass():
        out = []
        for g, lo in enumerate(range(0, n_clips, 8)):
            batch = sources[lo:lo + 8]
            d = draw_batch(2004, g, [c.shape[0] for c in batch], cfg)
            fixed = np.stack([oaug.to_target_length(c, int(p)) for c, p in zip(batch, d.pad_before)])
            b = fixed.shape[0]
            aug = oaug.augment_batch(
                fixed, colored_base=d.colored_base if d.colored_apply else None, colored_snr_db=d.colored_snr_db,
                gain_db=d.gain_db if d.gain_apply else None,
                noise=noise[:b * spec.CLIP_SAMPLES].reshape(b, -1) if d.background_apply else None, noise_snr_db=d.noise_snr_db,
                rir=rirs[g % len(rirs)] if d.reverb_apply else None, dtype=np.float32)
            out.append(aug)
        audio = np.concatenate(out)
        return opipe.speech_embeddings(
            [a for a in audio], mel_fn=lambda a: omel.mel_spectrogram(a, dtype=np.float32),
            embed_fn=lambda w: oembed.speech_embedding_model(w, weights), spectrogram_batch_size=32, embedding_batch_size=32)

    for _ in range(warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(steps):
        emb = one_pass()
    dt = (time.perf_counter() - t0) / steps
    assert emb.shape == (n_clips, 16, 96)
    return n_clips * CLIP_SECONDS / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = int(os.environ.get("HB_BENCH_REF_CLIPS", "256"))
    value, dt = cpu_reference_clips_per_s(sample, threads, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config("cpu-f32"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} clips per step: reference control flow (aug batch 8, 4 mel windows + 16 embedding windows per clip, "
                                   "batch 32) with the oracle's numpy/torch-CPU kernels; ORT/torch_audiomentations/speechbrain unavailable offline"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from heybuddy_b200 import _native, spec
    from heybuddy_b200.dataset.augmented import AugmentedAudioGenerator
    from heybuddy_b200.embeddings import DEFAULT_EMBED_PRECISION, SpeechEmbeddings
    from heybuddy_b200.pipeline import FeaturizePipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    precision = args.precision or DEFAULT_EMBED_PRECISION
    lib = _native.load()

    # ---- inputs: every rank owns its own clips (global batch ids rank*... so draws are world-size independent) ----
    noise_bank, rir_bank = make_banks(device)
    speech = SpeechEmbeddings(device_id=local_rank, precision=precision)
    aug = AugmentedAudioGenerator(
        [], device_id=local_rank, augmentation_dataset=noise_bank, impulse_response_dataset=rir_bank, batch_size=AUG_BATCH,
        colored_noise_min_f_decay=0.0, colored_noise_max_f_decay=0.0, seed=2004,
        first_batch=rank * POOL * (CHUNK // AUG_BATCH))
    aug._noise_cursor = (rank * 211) % NOISE_CLIPS
    pipe = FeaturizePipeline(aug, speech, device_id=local_rank)
    pool_clips = [make_sources(CHUNK, 2001 + 97 * rank + i, device).pin() for i in range(POOL)]
    pool_tables = [aug.next_table(c.lengths) for c in pool_clips]
    pool_dev = [pipe.upload(c, t) for c, t in zip(pool_clips, pool_tables)]
    # the same draws cut into sub-chunk tables for the pipelined host path (tables are per chunk of whole batches)
    sub = min(E2E_SUB, CHUNK)
    assert sub % AUG_BATCH == 0 and CHUNK % sub == 0
    bps = sub // AUG_BATCH
    from heybuddy_b200.dataset.draws import DrawTable
    pool_subtables = []
    for t in pool_tables:
        parts = []
        for lo in range(0, len(t.batches), bps):
            d = DrawTable(cfg=t.cfg, seed=t.seed)
            d.batches, d.noise_clip_cursor, d.rir_index = t.batches[lo:lo + bps], t.noise_clip_cursor[lo:lo + bps], t.rir_index[lo:lo + bps]
            parts.append(d)
        pool_subtables.append(parts)
    out_dev = torch.empty((CHUNK, 16, spec.EMB_DIM), dtype=torch.float32, device=device)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs -------------------------------------------------------------------------
    for i in range(args.warmup):
        pipe.run_device(pool_dev[i % POOL], out=out_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.hb_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        pipe.run_device(pool_dev[(args.warmup + i) % POOL], out=out_dev)     # the product path: one hb_featurize_i16 call per step
    stop.record()
    barrier()
    launches = lib.hb_launch_count() - launches0
    elapsed_ms = start.elapsed_time(stop)
    # per-stage CUDA-event times: a second, untimed-for-`value` pass with stage marks (the staged calls launch the same kernels)
    pipe.profile = True
    pipe.run_device(pool_dev[0], out=out_dev)
    barrier()
    pipe.collect_stage_times()
    for i in range(args.steps):
        pipe.run_device(pool_dev[(args.warmup + i) % POOL], out=out_dev)
    barrier()
    pipe.profile = False
    stage_ms = pipe.collect_stage_times()
    checksum = float(out_dev.float().abs().mean().item())

    # ---- e2e: host buffers through the public API, copies inside the timed region -----------------------------
    host_out = torch.empty((CHUNK, 16, spec.EMB_DIM), dtype=torch.float32).pin_memory()   # D2H lands here directly
    # ONE streaming call over the K steps' host chunks: every step's inputs cross PCIe inside the timed region and its
    # [CHUNK,16,96] result is read back; uploads of step i+1 overlap the compute of step i (pipeline fill / drain paid once)
    items = [(pool_clips[(args.warmup + i) % POOL], pool_subtables[(args.warmup + i) % POOL], host_out) for i in range(args.steps)]
    # warm-up = the same streaming call over max(W, 3) steps, untimed: on a freshly booted box the first process that moves this
    # much pinned memory pays one-off host / IOMMU first-touch costs (the first e2e pass was up to 3x slower than the second)
    pipe.featurize_stream([items[i % len(items)] for i in range(max(args.warmup, 3))], sub)
    barrier()
    t0 = time.perf_counter()
    h2d, d2h = pipe.featurize_stream(items, sub)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_host_busy = max(0.0, 1.0 - pipe.last_stream_wait_s / max(e2e_s, 1e-9))   # share of the e2e time the host thread was NOT waiting on the GPU
    clocks = sampler.stop()

    times = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(times[0]), float(times[1])
    total_clips = CHUNK * args.steps * world
    value = total_clips * CLIP_SECONDS / (elapsed_ms * 1e-3)
    e2e_value = total_clips * CLIP_SECONDS / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = {}
        peaks_src = "fallback (B200_PROFILING.md): 6650 GB/s, 1590 TFLOP/s"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
            peaks_src = "MEASURED_PEAKS.json"
        except Exception:
            pass
        # dominant stage = the embedding conv stack (tensor-core bound by design)
        flops_per_clip = 2.0 * spec.embedding_macs_per_clip(spec.mel_frames(spec.CLIP_SAMPLES))
        embed_ms = stage_ms.get("embed", 0.0)
        embed_launches = max(1, args.steps)
        achieved_tflops = flops_per_clip * CHUNK * args.steps / (embed_ms * 1e-3) / 1e12 if embed_ms > 0 else None
        tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0 if not peaks else peaks.get("bf16_tflops", 1590.0)))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # augment = length fix + K1-K4 in one kernel: int16 source (28,800 B on average) + noise row + f32 result
        stage_bytes = {"augment": 28800 + 92160 + 92160, "mel": 110208}
        stages = {}
        for name, ms in stage_ms.items():
            entry = {"ms_per_step": ms / args.steps, "share": ms / max(sum(stage_ms.values()), 1e-9)}
            if name in ("augment", "mel") and ms > 0:
                gbs = stage_bytes[name] * CHUNK * args.steps / (ms * 1e-3) / 1e9
                entry.update({"achieved_gbs": gbs, "hbm_frac": gbs / hbm_peak})
            stages[name] = entry
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (embed), f32 (augment, mel)" if precision == "f16" else "f32",
            "data": "synthetic", "config": workload_config(precision),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "host_busy_frac": e2e_host_busy,
                    "api": f"FeaturizePipeline.featurize_stream over the K steps' host chunks (pinned int16 clips in, pinned f32 [n,16,96] out, "
                           f"{sub}-clip sub-chunks, H2D / compute / D2H on three streams)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "kernel": "embedding conv stack (hb_embed_clips)", "bound": "tensor", "achieved": achieved_tflops,
                "peak": tensor_peak, "unit": "TFLOP/s", "frac": (achieved_tflops / tensor_peak) if achieved_tflops else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of the stage's 7 launches, one ncu --set full capture of a steady
                # 8192-clip step (profiles/r1b_ncu_summary.txt: 1.49 GB read + 1.18 GB written), scaled to this run's chunk size
                "traffic": 2.67e9 * CHUNK / 8192, "traffic_unit": "bytes per step (embed stage, ncu)", "peak_source": f"{peaks_src} (bf16 sustained; f16 runs at the bf16 rate)",
                "algorithmic_flops_per_clip": flops_per_clip,
                "note": "fully-convolutional evaluation: one 141-frame strip per clip instead of 16 windows (14 unique)",
            },
            "stages": stages,
            "output_checksum": checksum,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            sample = int(os.environ.get("HB_BENCH_CPU_CLIPS", "1536"))   # ~10 s of host work
            v, dt = cpu_reference_clips_per_s(sample, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{sample} clips, reference control flow (aug batch 8, 4 mel windows + 16 embedding windows per clip, batch 32) "
                          f"with the oracle's numpy/torch-CPU kernels, {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
