"""
Seeded augmentation draw table.

The reference's augmentation RNG is unseeded (np.random / torch.rand / dataset.shuffle(),
SURVEY.md 0.6, A.3 item 7), so "identical inputs" is defined by a *draw table*: for every
augmentation batch g (clips [g*B, (g+1)*B)) the random choices the reference would make, generated
by a seeded host RNG in the reference's call order (augmented.py:314-392):

    per-clip pad offsets (augmented.py:222)            -> pad_before[n]
    [per transform: apply coin, then parameters] in Compose order
        pitch shift, band stop                         (K9: not on the north-star path, probs forced to 0)
        coloured noise: coin, snr, f_decay, N(0,1)[16000] pattern (augmented.py:107-115)
        gain: coin, dB ~ U(-18, 6)                     (augmented.py:116-120)
    background coin (augmented.py:383)                 -> noise stream advance -> rand(B) SNRs (:269-270)
    reverb coin (augmented.py:387)                     -> RIR advance

Batch g's draws come from ``Generator(PCG64([seed, g, stream]))`` so the table does not depend on how
batches are sharded over ranks; the stateful cursors (noise stream position, RIR index) are prefix
sums over the table.  The oracle (tests) and the CUDA kernel consume the same table.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from heybuddy_b200 import spec
from heybuddy_b200.constants import *  # noqa: F401,F403


@dataclass
class AugmentConfig:
    """Probabilities / ranges of the north-star transforms (defaults: reference constants.py)."""
    batch_size: int = 128
    target_samples: int = spec.CLIP_SAMPLES
    colored_noise_prob: float = DEFAULT_AUGMENT_COLORED_NOISE_PROB
    colored_noise_min_snr_db: float = DEFAULT_AUGMENT_COLORED_NOISE_MIN_SNR_DB
    colored_noise_max_snr_db: float = DEFAULT_AUGMENT_COLORED_NOISE_MAX_SNR_DB
    colored_noise_min_f_decay: float = DEFAULT_AUGMENT_COLORED_NOISE_MIN_F_DECAY
    colored_noise_max_f_decay: float = DEFAULT_AUGMENT_COLORED_NOISE_MAX_F_DECAY
    gain_prob: float = DEFAULT_AUGMENT_GAIN_PROB
    gain_min_db: float = spec.GAIN_MIN_DB
    gain_max_db: float = spec.GAIN_MAX_DB
    background_noise_prob: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_PROB
    background_noise_min_snr_db: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_MIN_SNR_DB
    background_noise_max_snr_db: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_MAX_SNR_DB
    reverb_prob: float = DEFAULT_AUGMENT_REVERB_PROB


def colored_noise_base(gauss: np.ndarray, f_decay: float) -> np.ndarray:
    """
    torch_audiomentations ``_gen_noise``: rfft of the 1 s N(0,1) pattern,
    ``1/linspace(1, sqrt(sr/2), bins)**f_decay`` mask, irfft, unit RMS.  float64 math, f32 result.
    Host-side table preparation (one 64 KB pattern per coloured batch), not on the per-clip path.
    """
    g = np.asarray(gauss, dtype=np.float64)
    sr = g.shape[0]
    s = np.fft.rfft(g)
    mask = 1.0 / (np.linspace(1.0, (sr / 2) ** 0.5, s.shape[0], dtype=np.float64) ** float(f_decay))
    c = np.fft.irfft(s * mask, n=sr)
    return (c / np.sqrt(np.mean(c * c))).astype(np.float32)


def pad_before_for(length: int, target: int, rng: np.random.Generator) -> int:
    """augmented.py:216-226: randint(int(s/4), int(3s/4)) (high exclusive); s == 1 pads right; s <= 0 none."""
    s = target - length
    if s <= 1:
        return 0
    lo, hi = int(s / 4), int(3 * s / 4)
    return int(rng.integers(lo, hi)) if hi > lo else lo


@dataclass
class BatchDraw:
    """Every random choice of one augmentation batch."""
    index: int
    pad_before: np.ndarray                    # i32 [b]
    colored_apply: bool = False
    colored_snr_db: float = 0.0
    colored_f_decay: float = 0.0
    colored_base: Optional[np.ndarray] = None  # f32 [16000] when applied
    gain_apply: bool = False
    gain_db: float = 0.0
    background_apply: bool = False
    noise_snr_db: Optional[np.ndarray] = None  # f32 [b] when applied
    reverb_apply: bool = False

    @property
    def gain_linear(self) -> float:
        return float(10.0 ** (self.gain_db / 20.0)) if self.gain_apply else 1.0


def draw_batch(seed: int, index: int, lengths: Sequence[int], cfg: AugmentConfig,
               have_background: bool = True, have_reverb: bool = True, light: bool = False) -> BatchDraw:
    """
    The draws of augmentation batch ``index`` (clip lengths ``lengths``), in the reference's order.
    ``light`` skips building the coloured pattern (only the coins are needed to advance the stream cursors).
    """
    # two sub-streams per batch: the pad offsets consume a length-dependent number of draws, the batch-level
    # coins must not depend on them (any rank can then replay the coins of earlier batches without their clips)
    rng_pad = np.random.Generator(np.random.PCG64([int(seed), int(index), 1]))
    rng = np.random.Generator(np.random.PCG64([int(seed), int(index), 0]))
    b = len(lengths)
    d = BatchDraw(index=index, pad_before=np.array(
        [pad_before_for(int(n), cfg.target_samples, rng_pad) for n in lengths], dtype=np.int32))
    # coloured noise
    if rng.random() < cfg.colored_noise_prob:
        d.colored_apply = True
        d.colored_snr_db = float(rng.uniform(cfg.colored_noise_min_snr_db, cfg.colored_noise_max_snr_db))
        d.colored_f_decay = float(rng.uniform(cfg.colored_noise_min_f_decay, cfg.colored_noise_max_f_decay))
        gauss = rng.standard_normal(spec.COLORED_BASE_SAMPLES)
        d.colored_base = None if light else colored_noise_base(gauss, d.colored_f_decay)
    # gain
    if rng.random() < cfg.gain_prob:
        d.gain_apply = True
        d.gain_db = float(rng.uniform(cfg.gain_min_db, cfg.gain_max_db))
    # background noise: one coin per batch, one SNR per clip
    if rng.random() < cfg.background_noise_prob and have_background:
        d.background_apply = True
        span = cfg.background_noise_max_snr_db - cfg.background_noise_min_snr_db
        d.noise_snr_db = (rng.random(b, dtype=np.float32) * np.float32(span)
                          + np.float32(cfg.background_noise_min_snr_db)).astype(np.float32)
    # reverb
    if rng.random() < cfg.reverb_prob and have_reverb:
        d.reverb_apply = True
    return d


@dataclass
class DrawTable:
    """Draws of consecutive batches plus the stateful cursors (noise stream clip, RIR index)."""
    cfg: AugmentConfig
    seed: int
    batches: List[BatchDraw] = field(default_factory=list)
    noise_clip_cursor: List[int] = field(default_factory=list)   # first noise-bank clip of the batch (-1: none)
    rir_index: List[int] = field(default_factory=list)           # RIR of the batch (-1: none)

    @classmethod
    def build(cls, lengths: Sequence[int], cfg: AugmentConfig, seed: int, noise_clip_lengths: Optional[np.ndarray] = None,
              num_rirs: int = 0, first_batch: int = 0, noise_cursor: int = 0, rir_cursor: int = 0) -> "DrawTable":
        """
        ``lengths``: source clip lengths in order.  ``noise_clip_lengths``: lengths of the bank's clips in
        stream order (the reference pulls whole clips until >= B*T samples, augmented.py:246-251, and drops the
        unused tail); ``num_rirs``: size of the RIR bank (one RIR per applied batch, wrapping).
        """
        t = cls(cfg=cfg, seed=seed)
        n = len(lengths)
        have_bg = noise_clip_lengths is not None and len(noise_clip_lengths) > 0
        have_rev = num_rirs > 0
        g = first_batch
        for start in range(0, n, cfg.batch_size):
            d = draw_batch(seed, g, lengths[start:start + cfg.batch_size], cfg, have_bg, have_rev)
            t.batches.append(d)
            if d.background_apply:
                need = len(d.pad_before) * cfg.target_samples
                t.noise_clip_cursor.append(noise_cursor)
                got = 0
                while got < need:
                    got += int(noise_clip_lengths[noise_cursor % len(noise_clip_lengths)])
                    noise_cursor += 1
                noise_cursor %= len(noise_clip_lengths)
            else:
                t.noise_clip_cursor.append(-1)
            if d.reverb_apply:
                t.rir_index.append(rir_cursor % num_rirs)
                rir_cursor += 1
            else:
                t.rir_index.append(-1)
            g += 1
        t.final_noise_cursor = noise_cursor
        t.final_rir_cursor = rir_cursor
        return t
