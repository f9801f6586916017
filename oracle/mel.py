"""
Oracle: 32-bin log-mel front end (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows the reference call site ``MelSpectrogramModel.__call__``
(/root/reference/src/python/heybuddy/spectrogram.py:23-32): input f32 ``[b, t]``
(already scaled by 32767, embeddings.py:182), model output ``[b, 1, F, 32]`` dB,
return ``squeeze(out)/10 + 2``.

The model itself (``mel-spectrogram.onnx``, sha256 ba2b0e0f...176f,
spectrogram.py:20-21) is not on disk -> **values are parity-unpinned**; shapes are
pinned (17280->105, 23040->141 per clip / 4x105 as executed, 12640->76 frames;
tests/test_embeddings.py:10,14, src/ts/src/models/mel-spectrogram.ts:38-42).
The restated algorithm is torchaudio's MelSpectrogram(n_fft=512, win_length=400,
hop_length=160, center=False, n_mels=32, f_min=60, f_max=3800, HTK, norm=None,
power=2) followed by 10*log10(max(P, 1e-10)) without top_db.
"""
from __future__ import annotations

import numpy as np

from heybuddy_b200 import spec


def frames_view(audio: np.ndarray) -> np.ndarray:
    """``[b, t] -> [b, F, 512]`` strided frames, frame f = samples [160 f, 160 f + 512)."""
    b, t = audio.shape
    n = spec.mel_frames(t)
    if n <= 0:
        return np.zeros((b, 0, spec.N_FFT), dtype=audio.dtype)
    idx = np.arange(n)[:, None] * spec.HOP + np.arange(spec.N_FFT)[None, :]
    return audio[:, idx]


def mel_db(audio: np.ndarray, dtype=np.float64) -> np.ndarray:
    """
    ``[b, t]`` (int16-range floats) -> dB mel ``[b, F, 32]`` (the ONNX model's output,
    squeezed).  ``dtype`` is the arithmetic type of the restatement (float64 =
    exact answer; float32 = what an fp32 implementation sees).
    """
    audio = np.asarray(audio)
    if audio.ndim == 1:
        audio = audio[None, :]
    x = frames_view(audio.astype(dtype))
    w = spec.hann_window_padded().astype(dtype)
    spectrum = np.fft.rfft(x * w, axis=-1)
    power = (spectrum.real.astype(dtype) ** 2 + spectrum.imag.astype(dtype) ** 2)
    fb = spec.mel_filterbank().astype(dtype)
    mel = power @ fb
    return (10.0 * np.log10(np.maximum(mel, dtype(spec.MEL_FLOOR)))).astype(dtype)


def mel_spectrogram(audio: np.ndarray, dtype=np.float64) -> np.ndarray:
    """The reference-visible value: ``squeeze(model(audio))/10 + 2`` as float32 ``[b, F, 32]``."""
    db = mel_db(audio, dtype=dtype)
    return (db / spec.MEL_POST_DIV + spec.MEL_POST_ADD).astype(np.float32)
