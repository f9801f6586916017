"""
Oracle: the reference's per-clip numpy augmentations (TEST INFRASTRUCTURE, see oracle/__init__.py).

``audiomentations.Compose([SevenBandParametricEQ, TanhDistortion])`` (reference dataset/augmented.py:79-90, applied at
:325-328).  ``audiomentations>=0.37`` (environment.yml:11) is absent from the image and from /root/reference:
**PARITY UNPINNED** -- restated from the library's published behaviour, with scipy's ``sosfilt`` / numpy's ``percentile`` doing
the arithmetic the library delegates to them.  The draws (coins, filter parameters, distortion amount) come from the draw table.
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


def apply_table(fixed: np.ndarray, table) -> np.ndarray:
    """Length-fixed clips f32 [n, T] -> the same after the table's K9 draws (EQ first, then distortion: Compose order)."""
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
    return out
