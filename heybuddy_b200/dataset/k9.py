"""
K9 -- the per-clip numpy augmentations of the reference (``audiomentations.Compose([SevenBandParametricEQ, TanhDistortion])``,
reference ``dataset/augmented.py:79-90``, applied to every length-fixed clip at ``:325-328`` before the batch transforms).

``audiomentations`` is absent from the image and from ``/root/reference``, so both transforms are RESTATED from the
library's published behaviour -- **parity unpinned** -- with every constant kept as data below so that it can be corrected:

* ``SevenBandParametricEQ(min_gain_db=-g, max_gain_db=g)``: one low-shelf, five peaking and one high-shelf biquad (RBJ cookbook
  forms), each with its centre frequency drawn uniformly on the mel scale inside its band, its gain uniform in ``[-g, g]`` dB and
  its Q uniform in the filter's range; the cascade is applied causally (``scipy.signal.sosfilt``, zero initial state).  A centre
  frequency above Nyquist is pulled to ``0.9999 * sr / 2`` (the library's guard against an unstable shelf at 16 kHz).
* ``TanhDistortion(min_distortion, max_distortion)``: ``amount ~ U(min, max)``; ``threshold = percentile(|x|, 100 - 99 * amount)``
  (linear interpolation); ``y = tanh(x * 0.5 / (threshold + 1e-6))``; ``y *= rms(x) / rms(y)`` when ``rms(x) > 1e-9``.

Draws come from the draw table's Philox generator (streams 4-6 of ``dataset/draws.py``): per clip a coin and the parameters of
each transform.  The biquad coefficients are computed on the host in float64 (vectorised) and shipped per selected clip; the
filters and the distortion run on the device (``hb_k9_eq_f32`` / ``hb_k9_tanh_f32``, ``csrc/k9.cu``).

The two remaining K9 transforms -- ``torch_audiomentations.PitchShift`` (phase-vocoder + resampler of ``torch_pitch_shift``) and
``BandStopFilter`` (``julius`` windowed-sinc low-pass pair with replicate padding, kernels up to several clip lengths long) -- are
not built: a non-zero probability raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from heybuddy_b200 import spec
from heybuddy_b200.dataset.draws import (STREAM_EQ_COIN, STREAM_EQ_GAIN, STREAM_TANH, AugmentConfig, philox4x32, uniform53)

__all__ = ["K9Draws", "EQ_BANDS", "biquad_sos", "apply_device"]

# (kind, min centre Hz, max centre Hz, min Q, max Q): SevenBandParametricEQ's filter bank (audiomentations defaults, restated)
EQ_BANDS = (
    ("low_shelf", 42.0, 95.0, 0.1, 0.999),
    ("peaking", 91.0, 204.0, 0.5, 5.0),
    ("peaking", 196.0, 441.0, 0.5, 5.0),
    ("peaking", 421.0, 948.0, 0.5, 5.0),
    ("peaking", 909.0, 2045.0, 0.5, 5.0),
    ("peaking", 1957.0, 4404.0, 0.5, 5.0),
    ("high_shelf", 4216.0, 9486.0, 0.1, 0.999),
)
N_BANDS = len(EQ_BANDS)
_S32, _MASK = np.uint64(32), np.uint64(0xFFFFFFFF)


def _hz_to_mel(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def _mel_to_hz(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def biquad_sos(kind: str, center_hz: np.ndarray, gain_db: np.ndarray, q: np.ndarray, sample_rate: int = spec.SAMPLE_RATE) -> np.ndarray:
    """RBJ cookbook biquads, vectorised over clips -> ``[n, 5]`` = (b0, b1, b2, a1, a2) / a0 in float64."""
    f = np.minimum(np.asarray(center_hz, dtype=np.float64), 0.9999 * sample_rate / 2.0)
    w0 = 2.0 * np.pi * f / sample_rate
    a_lin = 10.0 ** (np.asarray(gain_db, dtype=np.float64) / 40.0)
    alpha = np.sin(w0) / 2.0 / np.asarray(q, dtype=np.float64)
    c = np.cos(w0)
    if kind == "peaking":
        b0, b1, b2 = 1.0 + alpha * a_lin, -2.0 * c, 1.0 - alpha * a_lin
        a0, a1, a2 = 1.0 + alpha / a_lin, -2.0 * c, 1.0 - alpha / a_lin
    else:
        s = 2.0 * np.sqrt(a_lin) * alpha
        sign = 1.0 if kind == "low_shelf" else -1.0          # the high shelf mirrors the (A - 1) cos terms
        b0 = a_lin * ((a_lin + 1.0) - sign * (a_lin - 1.0) * c + s)
        b1 = sign * 2.0 * a_lin * ((a_lin - 1.0) - sign * (a_lin + 1.0) * c)
        b2 = a_lin * ((a_lin + 1.0) - sign * (a_lin - 1.0) * c - s)
        a0 = (a_lin + 1.0) + sign * (a_lin - 1.0) * c + s
        a1 = -sign * 2.0 * ((a_lin - 1.0) + sign * (a_lin + 1.0) * c)
        a2 = (a_lin + 1.0) + sign * (a_lin - 1.0) * c - s
    return np.stack([b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0], axis=-1)


class K9Draws:
    """Per-clip K9 draws of a :class:`~heybuddy_b200.dataset.draws.DrawTable` (struct of arrays over the table's clips)."""

    def __init__(self) -> None:
        self.eq_apply = np.zeros(0, bool)
        self.eq_sos = np.zeros((0, N_BANDS, 5), np.float64)      # rows of the clips with eq_apply, in clip order
        self.tanh_apply = np.zeros(0, bool)
        self.tanh_amount = np.zeros(0, np.float64)               # per clip (0 where not applied)

    @classmethod
    def build(cls, cfg: AugmentConfig, seed: int, gids: np.ndarray, sizes: np.ndarray, within: np.ndarray, g_of: np.ndarray,
              batch_u: np.ndarray) -> "K9Draws":
        if cfg.pitch_shift_prob or cfg.band_stop_prob:
            raise NotImplementedError("PitchShift / BandStopFilter (torch_audiomentations, augmented.py:93-106) are not built; set their "
                                      "probabilities to 0 (heybuddy_b200/dataset/k9.py)")
        k = cls()
        n = int(within.shape[0])
        lo, hi = g_of & _MASK, g_of >> _S32
        # SevenBandParametricEQ: coin, then per band (centre on the mel scale, gain dB, Q)
        x0, x1, _, _ = philox4x32(within, STREAM_EQ_COIN, lo, hi, seed)
        k.eq_apply = uniform53(x0, x1) < cfg.seven_band_prob
        sel = np.nonzero(k.eq_apply)[0]
        sos = np.zeros((sel.shape[0], N_BANDS, 5), dtype=np.float64)
        if sel.shape[0]:
            g = float(cfg.seven_band_gain_db)
            for b, (kind, f_lo, f_hi, q_lo, q_hi) in enumerate(EQ_BANDS):
                u = []
                for j in range(3):
                    a0, a1, _, _ = philox4x32(within[sel] * np.uint64(32) + np.uint64(3 * b + j), STREAM_EQ_GAIN, lo[sel], hi[sel], seed)
                    u.append(uniform53(a0, a1))
                center = _mel_to_hz(_hz_to_mel(f_lo) + u[0] * (_hz_to_mel(f_hi) - _hz_to_mel(f_lo)))
                sos[:, b] = biquad_sos(kind, center, -g + u[1] * 2.0 * g, q_lo + u[2] * (q_hi - q_lo))
        k.eq_sos = sos
        # TanhDistortion: coin (words 0, 1) and amount (words 2, 3) of one counter
        t0, t1, t2, t3 = philox4x32(within, STREAM_TANH, lo, hi, seed)
        k.tanh_apply = uniform53(t0, t1) < cfg.tanh_distortion_prob
        amount = cfg.tanh_min_distortion + uniform53(t2, t3) * (cfg.tanh_max_distortion - cfg.tanh_min_distortion)
        k.tanh_amount = np.where(k.tanh_apply, amount, 0.0)
        return k

    def slice(self, b0: int, b1: int, r0: int, r1: int) -> "K9Draws":
        k = K9Draws()
        k.eq_apply, k.tanh_apply, k.tanh_amount = self.eq_apply[r0:r1], self.tanh_apply[r0:r1], self.tanh_amount[r0:r1]
        e0, e1 = int(np.count_nonzero(self.eq_apply[:r0])), int(np.count_nonzero(self.eq_apply[:r1]))
        k.eq_sos = self.eq_sos[e0:e1]
        return k

    def pack(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """(eq clip indices i32[k], sos f64[k,7,5], tanh clip indices i32[j], tanh amounts f32[j])."""
        eq_idx = np.nonzero(self.eq_apply)[0].astype(np.int32)
        th_idx = np.nonzero(self.tanh_apply)[0].astype(np.int32)
        return eq_idx, np.ascontiguousarray(self.eq_sos, dtype=np.float64), th_idx, self.tanh_amount[th_idx].astype(np.float32)


def apply_device(fixed, table, gen=None):
    """Runs the table's K9 transforms in place on cuda f32 ``[n, T]`` length-fixed clips (EQ first, then distortion: Compose order)."""
    import torch

    from heybuddy_b200 import _native

    k9: Optional[K9Draws] = table.k9
    if k9 is None:
        return fixed
    lib = _native.load()
    n, t = fixed.shape
    eq_idx, sos, th_idx, amount = k9.pack()
    dev = fixed.device
    with torch.cuda.device(dev):
        st = _native.stream_ptr(dev)
        if eq_idx.size:
            idx_d, sos_d = torch.from_numpy(eq_idx).to(dev), torch.from_numpy(sos).to(dev)
            _native.check(lib.hb_k9_eq_f32(fixed.data_ptr(), idx_d.data_ptr(), sos_d.data_ptr(), int(eq_idx.size), t, st), "hb_k9_eq_f32")
        if th_idx.size:
            idx_d, amt_d = torch.from_numpy(th_idx).to(dev), torch.from_numpy(amount).to(dev)
            _native.check(lib.hb_k9_tanh_f32(fixed.data_ptr(), idx_d.data_ptr(), amt_d.data_ptr(), int(th_idx.size), t, st), "hb_k9_tanh_f32")
    return fixed
