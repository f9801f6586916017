"""
Oracle: the reference's per-clip numpy augmentations (TEST INFRASTRUCTURE, see oracle/__init__.py).

``audiomentations.Compose([SevenBandParametricEQ, TanhDistortion])`` (reference dataset/augmented.py:79-90, applied at
:325-328).  ``audiomentations>=0.37`` (environment.yml:11) is absent from the image and from /root/reference:
**PARITY UNPINNED** -- restated from the library's published behaviour, with scipy's ``sosfilt`` / numpy's ``percentile`` doing
the arithmetic the library delegates to them.  The draws (coins, filter parameters, distortion amount) come from the draw table.

The two batch transforms ahead of AddColoredNoise / Gain (``torch_audiomentations.Compose``, augmented.py:93-106, per_batch):
``BandStopFilter`` follows ``julius.BandPassFilter`` / ``LowPassFilters`` line by line in torch float32 (julius absent: unpinned);
``PitchShift`` is ``torch_pitch_shift.pitch_shift`` with the REAL ``torch.stft`` / ``torchaudio`` TimeStretch (phase vocoder) /
``torch.istft`` / ``torchaudio`` Resample calls the library makes (those four are pinned by being the library's own code; the glue --
n_fft = sr // 64, hop = n_fft // 32, rate = 1 / shift, crop / pad -- is restated, unpinned).
"""
from __future__ import annotations

import numpy as np


def seven_band_eq(x: np.ndarray, sos5: np.ndarray) -> np.ndarray:
    """``x`` f32 [T]; ``sos5`` [7, 5] = (b0, b1, b2, a1, a2) / a0 per section -> scipy.signal.sosfilt(sos, x) as float32."""
    from scipy.signal import sosfilt

    sos = np.concatenate([sos5[:, :3], np.ones((sos5.shape[0], 1)), sos5[:, 3:]], axis=1)
    return sosfilt(sos, np.asarray(x, dtype=np.float64)).astype(np.float32)


def tanh_distortion(x: np.ndarray, amount: float) -> np.ndarray:
    """audiomentations TanhDistortion.apply: percentile threshold -> tanh -> loudness (RMS) match."""
    x = np.asarray(x, dtype=np.float64)
    threshold = np.percentile(np.abs(x), 100.0 - 99.0 * float(amount))
    y = np.tanh(x * (0.5 / (threshold + 1e-6)))
    rms_before = np.sqrt(np.mean(x * x))
    if rms_before > 1e-9:
        y = y * (rms_before / np.sqrt(np.mean(y * y)))
    return y.astype(np.float32)


def band_stop(x: np.ndarray, low: float, high: float, zeros: int = 8) -> np.ndarray:
    """
    ``x`` f32 [b, T] -> ``x - julius.bandpass_filter(x, low, high)`` (cut-offs as fractions of the sample rate), following
    julius.lowpass.LowPassFilters: one half size for both filters from the LOWER cut-off, Hann window, sinc, unit-sum
    normalisation, replicate padding, ``fft_conv1d`` above 32 taps (here: float64 FFT convolution) and ``conv1d`` below.
    """
    import math

    import torch
    import torch.nn.functional as F

    inp = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    half = int(zeros / min(c for c in (low, high) if c > 0) / 2)
    window = torch.hann_window(2 * half + 1, periodic=False)
    time = torch.arange(-half, half + 1)
    filters = []
    for cutoff in (low, high):
        arg = 2 * cutoff * math.pi * time
        sinc = torch.where(arg == 0, torch.ones_like(arg), torch.sin(arg) / arg)
        f = 2 * cutoff * window * sinc
        filters.append(f / f.sum())
    filt = torch.stack(filters)[:, None]                       # [2, 1, 2 half + 1]
    padded = F.pad(inp[:, None], (half, half), mode="replicate")
    if half > 32:
        n = padded.shape[-1] + filt.shape[-1] - 1
        spec_x = torch.fft.rfft(padded.double(), n=n)
        spec_f = torch.fft.rfft(filt.double().flip(-1), n=n)     # conv1d is a cross-correlation
        full = torch.fft.irfft(spec_x * spec_f.permute(1, 0, 2), n=n)
        lows = full[..., filt.shape[-1] - 1: filt.shape[-1] - 1 + inp.shape[-1]].float()
    else:
        lows = F.conv1d(padded, filt)
    return (inp - (lows[:, 1] - lows[:, 0])).numpy()


def pitch_shift(x: np.ndarray, shift, sample_rate: int = 16000) -> np.ndarray:
    """``x`` f32 [b, T], ``shift`` a ``fractions.Fraction`` -> ``torch_pitch_shift.pitch_shift`` (the library's calls, its glue restated)."""
    import warnings

    import torch
    import torchaudio.transforms as T

    inp = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    n_fft = sample_rate // 64
    hop = n_fft // 32
    resampler = T.Resample(sample_rate, int(sample_rate / shift))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                        # "A window was not provided": the library passes none
        out = torch.stft(inp, n_fft, hop, return_complex=True)[None, ...]
        stretcher = T.TimeStretch(fixed_rate=float(1 / shift), n_freq=out.shape[2], hop_length=hop)
        out = stretcher(out)
        out = torch.istft(out[0], n_fft, hop)
    out = resampler(out)
    if out.shape[1] >= inp.shape[1]:
        out = out[:, :inp.shape[1]]
    else:
        out = torch.nn.functional.pad(out, (0, inp.shape[1] - out.shape[1], 0, 0))
    return out.numpy()


def apply_table(fixed: np.ndarray, table) -> np.ndarray:
    """
    Length-fixed clips f32 [n, T] -> the same after the table's K9 draws: per clip EQ then distortion (audiomentations Compose
    order, augmented.py:325-328), then per batch PitchShift and BandStopFilter (the head of the batch Compose, :369-372).
    """
    out = np.array(fixed, dtype=np.float32, copy=True)
    k9 = table.k9
    if k9 is None:
        return out
    e = 0
    for i in range(out.shape[0]):
        if k9.eq_apply[i]:
            out[i] = seven_band_eq(out[i], k9.eq_sos[e])
            e += 1
        if k9.tanh_apply[i]:
            out[i] = tanh_distortion(out[i], float(k9.tanh_amount[i]))
    starts = np.concatenate(([0], np.cumsum(k9.sizes))).astype(int)
    for b in range(len(k9.sizes)):
        r0, r1 = starts[b], starts[b + 1]
        if k9.ps_apply[b]:
            out[r0:r1] = pitch_shift(out[r0:r1], k9.ps_shift[b], k9.sample_rate)
        if k9.bs_apply[b]:
            out[r0:r1] = band_stop(out[r0:r1], float(k9.bs_low[b]), float(k9.bs_high[b]))
    return out
