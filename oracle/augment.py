"""
Oracle: batched augmentation (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, per augmentation batch, what
``AugmentedAudioGenerator.execute_augment_batch``
(/root/reference/src/python/heybuddy/dataset/augmented.py:297-394) computes for
the north-star subset of transforms, in the reference's order:

  a1  to_target_length                      augmented.py:200-232
  K1  AddColoredNoise (per_batch)           augmented.py:107-115  [torch_audiomentations>=0.11, absent: UNPINNED]
  K2  Gain (per_batch)                      augmented.py:116-120  [torch_audiomentations, absent: UNPINNED]
  K3  background noise at per-clip SNR      augmented.py:234-276  [torchaudio.functional.add_noise: PINNED, present]
  K4  reverb, one RIR per batch             augmented.py:387-392  [speechbrain>=1.0 reverberate, absent: UNPINNED]

The reference's RNG is unseeded; both this oracle and the CUDA
kernel consume one *draw table* row per batch (``heybuddy_b200.draws``), generated
on the host in the reference's call order.

All arithmetic is numpy; ``dtype`` selects float64 (exact answer) or float32.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from heybuddy_b200 import spec


def to_target_length(audio: np.ndarray, pad_before: int, target: int = spec.CLIP_SAMPLES) -> np.ndarray:
    """
    augmented.py:200-232 with the random left pad supplied by the draw table:
    int16 -> f32 / 32768; >= target samples -> front-truncate; else zero-pad with
    ``pad_before`` zeros on the left (the reference draws it from
    randint(total//4, 3*total//4); total == 1 pads right).
    """
    audio = np.asarray(audio)
    if audio.dtype == np.int16:
        audio = audio.astype(np.float32) / 32768.0
    n = audio.shape[0]
    if n >= target:
        return audio[:target].astype(np.float32)
    out = np.zeros(target, dtype=np.float32)
    out[pad_before:pad_before + n] = audio
    return out


def colored_noise_base(gauss: np.ndarray, f_decay: float, dtype=np.float64) -> np.ndarray:
    """
    torch_audiomentations ``_gen_noise``: 1 s N(0,1) pattern -> rfft -> multiply by
    ``1/linspace(1, sqrt(sr/2), n_bins)**f_decay`` -> irfft -> unit RMS.
    ``gauss`` is the f32[16000] N(0,1) draw.  Returns the 1 s unit-RMS pattern.
    """
    g = np.asarray(gauss, dtype=dtype)
    sr = g.shape[0]
    s = np.fft.rfft(g)
    mask = 1.0 / (np.linspace(1.0, (sr / 2) ** 0.5, s.shape[0], dtype=dtype) ** dtype(f_decay))
    c = np.fft.irfft(s * mask, n=sr).astype(dtype)
    c = c / np.sqrt(np.mean(c * c))
    return c.astype(np.float32)


def add_colored_noise(x: np.ndarray, base: np.ndarray, snr_db: float, dtype=np.float64) -> np.ndarray:
    """``x[b] += rms(x[b]) / 10**(snr/20) * tile(base)[:T]`` -- same pattern and snr for the whole batch."""
    x = np.asarray(x, dtype=dtype)
    t = x.shape[-1]
    reps = int(np.ceil(t / base.shape[0]))
    noise = np.tile(base.astype(dtype), reps)[:t]
    rms = np.sqrt(np.mean(x * x, axis=-1, keepdims=True))
    return x + (rms / dtype(10.0 ** (snr_db / 20.0))) * noise[None, :]


def gain(x: np.ndarray, gain_db: float, dtype=np.float64) -> np.ndarray:
    """``x * 10**(g/20)``."""
    return np.asarray(x, dtype=dtype) * dtype(10.0 ** (gain_db / 20.0))


def add_noise(x: np.ndarray, noise: np.ndarray, snr_db: np.ndarray, dtype=np.float64) -> np.ndarray:
    """
    ``torchaudio.functional.add_noise`` (call site augmented.py:272-276):
    scale = 10**((10*(log10||x||^2 - log10||n||^2) - snr)/20); y = x + scale*n.
    All-zero x or n -> log10(0) -> inf/nan exactly like the dependency.
    """
    x = np.asarray(x, dtype=dtype)
    n = np.asarray(noise, dtype=dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        e_s = np.sum(x * x, axis=-1)
        e_n = np.sum(n * n, axis=-1)
        orig = 10.0 * (np.log10(e_s) - np.log10(e_n))
        scale = 10.0 ** ((orig - np.asarray(snr_db, dtype=dtype)) / 20.0)
        return x + scale[:, None] * n


def rotate_rir(rir: np.ndarray, t: int) -> np.ndarray:
    """
    speechbrain ``convolve1d(use_fft=True, rotation_index=argmax|rir|)``: truncate to
    ``t`` samples, then ``[rir[d:], zeros(t-L), rir[:d]]`` so the direct path sits at lag 0.
    """
    rir = np.asarray(rir, dtype=np.float32)
    d = int(np.argmax(np.abs(rir)))
    if rir.shape[0] > t:
        rir = rir[:t]
    d = min(d, rir.shape[0])
    k = np.zeros(t, dtype=np.float32)
    after, before = rir[d:], rir[:d]
    k[:after.shape[0]] = after
    if before.shape[0]:
        k[t - before.shape[0]:] = before
    return k


def reverberate(x: np.ndarray, rir: np.ndarray, dtype=np.float64) -> np.ndarray:
    """
    speechbrain ``reverberate(waveforms, rir, rescale_amp="avg")`` (call site
    augmented.py:388-392): circular length-T convolution with the rotated RIR,
    then rescale so mean|y| == mean|x|:  y * mean|x| / (mean|y| + 1e-14).
    """
    x = np.asarray(x, dtype=dtype)
    t = x.shape[-1]
    k = rotate_rir(rir, t).astype(dtype)
    y = np.fft.irfft(np.fft.rfft(x, axis=-1) * np.fft.rfft(k)[None, :], n=t, axis=-1).astype(dtype)
    amp_x = np.mean(np.abs(x), axis=-1, keepdims=True)
    amp_y = np.mean(np.abs(y), axis=-1, keepdims=True)
    return y / (amp_y + dtype(spec.REVERB_EPS)) * amp_x


def augment_batch(
    clips: np.ndarray,
    *,
    colored_base: Optional[np.ndarray] = None,
    colored_snr_db: float = 0.0,
    gain_db: Optional[float] = None,
    noise: Optional[np.ndarray] = None,
    noise_snr_db: Optional[Sequence[float]] = None,
    rir: Optional[np.ndarray] = None,
    dtype=np.float64,
) -> np.ndarray:
    """
    One augmentation batch in the reference's order (augmented.py:363-392):
    coloured noise -> gain -> background noise -> reverb.  ``clips`` is the
    already length-fixed ``f32[B, T]`` stack (augmented.py:363-366).  A stage whose
    argument is ``None`` was not drawn for this batch.
    """
    x = np.asarray(clips, dtype=dtype)
    if colored_base is not None:
        x = add_colored_noise(x, colored_base, colored_snr_db, dtype=dtype)
    if gain_db is not None:
        x = gain(x, gain_db, dtype=dtype)
    if noise is not None:
        x = add_noise(x, noise, np.asarray(noise_snr_db), dtype=dtype)
    if rir is not None:
        x = reverberate(x, rir, dtype=dtype)
    return x.astype(np.float32)
