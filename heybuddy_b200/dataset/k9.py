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

The two batch transforms that precede AddColoredNoise / Gain in the reference's ``torch_audiomentations.Compose``
(``augmented.py:93-106``; ``mode="per_batch"``: one coin and one parameter set per augmentation batch) are restated the same way --
``torch_audiomentations``, ``julius`` and ``torch_pitch_shift`` are all absent, **parity unpinned**:

* ``BandStopFilter()`` (library defaults: centre 200-4000 Hz uniform on the mel scale, bandwidth fraction 0.5-1.99):
  ``x - julius.bandpass_filter(x, low, high)`` = ``x - (lowpass_high(x) - lowpass_low(x))``; both low-passes are Hann-windowed sincs of
  ``2 h + 1`` taps with ``h = int(zeros / low / 2)``, ``zeros = 8``, each normalised to unit sum, applied to the clip replicate-padded by
  ``h``.  The device runs it as a uniformly partitioned overlap-save convolution on the clip's exact-length FFT (``hb_k9_bandstop_f32``).
* ``PitchShift(-s, +s semitones, sample_rate)``: the shift is one of ``torch_pitch_shift.get_fast_shifts`` -- ratios of products of the
  sample rate's prime factors inside the semitone range (16 kHz, +-3 semitones: 125/128 and 128/125) -- and ``pitch_shift`` is
  ``torch.stft(n_fft = sr // 64, hop = n_fft // 32, rectangular window)`` -> ``torchaudio`` phase vocoder at rate ``1 / shift`` ->
  ``torch.istft`` -> ``torchaudio`` sinc resampling ``sr -> int(sr / shift)`` -> crop / zero-pad to the clip length (``hb_k9_pitch_f32``).
"""
from __future__ import annotations

from fractions import Fraction
from typing import Dict, List, Optional, Tuple

import numpy as np

from heybuddy_b200 import spec
from heybuddy_b200.dataset.draws import (STREAM_EQ_COIN, STREAM_EQ_GAIN, STREAM_TANH, U_BANDSTOP_CENTER, U_BANDSTOP_COIN, U_BANDSTOP_WIDTH,
                                         U_PITCH_COIN, U_PITCH_SEMITONES, AugmentConfig, philox4x32, uniform53)

__all__ = ["K9Draws", "EQ_BANDS", "BANDSTOP", "HOST_ONLY", "biquad_sos", "bandstop_cutoffs", "bandstop_fir", "fast_shifts", "sinc_resample_kernel",
           "pitch_tables", "pitch_plan", "apply_packed", "apply_device"]

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
# torch_audiomentations.BandStopFilter defaults (the reference passes none, augmented.py:102-106) + julius' ``zeros``
BANDSTOP = {"min_center_hz": 200.0, "max_center_hz": 4000.0, "min_bandwidth_fraction": 0.5, "max_bandwidth_fraction": 1.99, "zeros": 8}
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


def bandstop_cutoffs(u_center: np.ndarray, u_width: np.ndarray, sample_rate: int = spec.SAMPLE_RATE) -> Tuple[np.ndarray, np.ndarray]:
    """Uniforms -> (low, high) cut-offs as fractions of the sample rate: centre uniform on the mel scale, bandwidth = centre * fraction."""
    lo_m, hi_m = _hz_to_mel(BANDSTOP["min_center_hz"]), _hz_to_mel(BANDSTOP["max_center_hz"])
    center = _mel_to_hz(lo_m + np.asarray(u_center, dtype=np.float64) * (hi_m - lo_m))
    frac = BANDSTOP["min_bandwidth_fraction"] + np.asarray(u_width, dtype=np.float64) * (BANDSTOP["max_bandwidth_fraction"] - BANDSTOP["min_bandwidth_fraction"])
    return center * (1.0 - 0.5 * frac) / sample_rate, center * (1.0 + 0.5 * frac) / sample_rate


def bandstop_fir(low: float, high: float) -> Tuple[np.ndarray, int]:
    """
    The band-PASS FIR ``lowpass_high - lowpass_low`` of ``julius.BandPassFilter(low, high)`` (f32 ``[2 h + 1]``) and its half size
    ``h = int(zeros / low / 2)``: ``2 c * hann(2 h + 1) * sinc(2 c pi t)``, ``t = -h .. h``, each divided by its sum.
    """
    h = int(BANDSTOP["zeros"] / float(low) / 2)
    t = np.arange(-h, h + 1, dtype=np.float64)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(2 * h + 1) / (2 * h)) if h > 0 else np.ones(1)

    def lowpass(c: float) -> np.ndarray:
        f = 2.0 * c * window * np.sinc(2.0 * c * t)        # np.sinc(x) = sin(pi x) / (pi x)
        return f / f.sum()

    return (lowpass(float(high)) - lowpass(float(low))).astype(np.float32), h


def _prime_factors(n: int) -> List[int]:
    out, d = [], 2
    while n > 1:
        while n % d == 0:
            out.append(d)
            n //= d
        d += 1
    return out


def fast_shifts(sample_rate: int, semitones: float) -> List[Fraction]:
    """
    ``torch_pitch_shift.get_fast_shifts`` with torch_audiomentations' condition: ratios i / j of products of subsets of the sample
    rate's prime factors with ``2 ** (-s / 12) <= ratio <= 2 ** (s / 12)`` and ``ratio != 1``, sorted ascending (the library holds
    them in a set; ``random.choices`` picks one per batch).
    """
    factors = _prime_factors(int(sample_rate))
    products = {1}
    for f in factors:
        products |= {p * f for p in products}
    products.discard(1)
    lo, hi = 2.0 ** (-float(semitones) / 12.0), 2.0 ** (float(semitones) / 12.0)
    out = {Fraction(i, j) for i in products for j in products}
    return sorted(r for r in out if lo <= float(r) <= hi and r != 1)


class K9Draws:
    """
    K9 draws of a :class:`~heybuddy_b200.dataset.draws.DrawTable`: per-clip arrays for the numpy transforms (struct of arrays over
    the table's clips) and per-batch arrays for PitchShift / BandStopFilter.
    """

    def __init__(self) -> None:
        self.eq_apply = np.zeros(0, bool)
        self.eq_sos = np.zeros((0, N_BANDS, 5), np.float64)      # rows of the clips with eq_apply, in clip order
        self.tanh_apply = np.zeros(0, bool)
        self.tanh_amount = np.zeros(0, np.float64)               # per clip (0 where not applied)
        self.sizes = np.zeros(0, np.int64)                       # clips per batch
        self.ps_apply = np.zeros(0, bool)                        # per batch
        self.ps_shift: List[Optional[Fraction]] = []             # per batch: the pitch ratio, or None
        self.bs_apply = np.zeros(0, bool)                        # per batch
        self.bs_low = np.zeros(0, np.float64)                    # per batch: cut-offs as fractions of the sample rate
        self.bs_high = np.zeros(0, np.float64)
        self.target_samples = spec.CLIP_SAMPLES
        self.sample_rate = spec.SAMPLE_RATE

    @classmethod
    def build(cls, cfg: AugmentConfig, seed: int, gids: np.ndarray, sizes: np.ndarray, within: np.ndarray, g_of: np.ndarray,
              batch_u: np.ndarray) -> "K9Draws":
        k = cls()
        k.target_samples = int(cfg.target_samples)
        k.sizes = np.asarray(sizes, dtype=np.int64)
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
        # per batch: PitchShift (coin, choice among the fast shifts) and BandStopFilter (coin, centre, bandwidth fraction)
        nb = int(batch_u.shape[0])
        k.ps_apply = batch_u[:, U_PITCH_COIN] < cfg.pitch_shift_prob
        k.ps_shift = [None] * nb
        if k.ps_apply.any():
            shifts = fast_shifts(k.sample_rate, cfg.pitch_shift_semitones)
            if not shifts:
                raise ValueError(f"PitchShift: no fast shift ratio within +-{cfg.pitch_shift_semitones} semitones at {k.sample_rate} Hz "
                                 "(torch_audiomentations raises for the same configuration)")
            for b in np.nonzero(k.ps_apply)[0]:
                k.ps_shift[b] = shifts[min(int(batch_u[b, U_PITCH_SEMITONES] * len(shifts)), len(shifts) - 1)]
        k.bs_apply = batch_u[:, U_BANDSTOP_COIN] < cfg.band_stop_prob
        low, high = bandstop_cutoffs(batch_u[:, U_BANDSTOP_CENTER], batch_u[:, U_BANDSTOP_WIDTH], k.sample_rate)
        k.bs_low, k.bs_high = np.where(k.bs_apply, low, 0.0), np.where(k.bs_apply, high, 0.0)
        return k

    def slice(self, b0: int, b1: int, r0: int, r1: int) -> "K9Draws":
        k = K9Draws()
        k.target_samples, k.sample_rate = self.target_samples, self.sample_rate
        k.eq_apply, k.tanh_apply, k.tanh_amount = self.eq_apply[r0:r1], self.tanh_apply[r0:r1], self.tanh_amount[r0:r1]
        e0, e1 = int(np.count_nonzero(self.eq_apply[:r0])), int(np.count_nonzero(self.eq_apply[:r1]))
        k.eq_sos = self.eq_sos[e0:e1]
        k.sizes, k.ps_apply, k.ps_shift = self.sizes[b0:b1], self.ps_apply[b0:b1], self.ps_shift[b0:b1]
        k.bs_apply, k.bs_low, k.bs_high = self.bs_apply[b0:b1], self.bs_low[b0:b1], self.bs_high[b0:b1]
        return k

    def any(self) -> bool:
        return bool(self.eq_apply.any() or self.tanh_apply.any() or self.ps_apply.any() or self.bs_apply.any())

    def _batch_rows(self, which: np.ndarray) -> np.ndarray:
        """Clip rows (table order) of the batches flagged in ``which``."""
        starts = np.concatenate(([0], np.cumsum(self.sizes)))
        rows = [np.arange(starts[b], starts[b + 1]) for b in np.nonzero(which)[0]]
        return np.concatenate(rows).astype(np.int32) if rows else np.zeros(0, np.int32)

    def pack(self) -> Dict[str, np.ndarray]:
        """
        Device-side form: ``eq_idx`` i32[ke], ``eq_sos`` f64[ke,7,5], ``th_idx`` i32[kt], ``th_amt`` f32[kt];
        ``bs_idx`` i32[kb] clips of the band-stop batches, ``bs_meta`` i32[kb,3] = (first tap row, partitions, half size h) and
        ``bs_taps`` f32[R,T]: every band-stop batch's FIR cut into partitions of T / 2 taps, one zero-padded row each;
        ``ps_idx`` i32[kp] clips of the pitch-shift batches grouped by ratio; ``ps_ratios`` i64[V,2] (numerator, denominator of the
        distinct ratios drawn) and ``ps_counts`` i64[V] (clips per ratio) stay on the host.
        """
        t = self.target_samples
        eq_idx = np.nonzero(self.eq_apply)[0].astype(np.int32)
        th_idx = np.nonzero(self.tanh_apply)[0].astype(np.int32)
        out = {"eq_idx": eq_idx, "eq_sos": np.ascontiguousarray(self.eq_sos, dtype=np.float64), "th_idx": th_idx,
               "th_amt": self.tanh_amount[th_idx].astype(np.float32)}
        part = t // 2
        rows: List[np.ndarray] = []
        meta: List[Tuple[int, int, int]] = []
        for b in np.nonzero(self.bs_apply)[0]:
            taps, h = bandstop_fir(self.bs_low[b], self.bs_high[b])
            n_part = -(-taps.shape[0] // part)
            block = np.zeros((n_part, t), dtype=np.float32)
            for p in range(n_part):
                seg = taps[p * part:(p + 1) * part]
                block[p, :seg.shape[0]] = seg
            meta += [(len(rows), n_part, h)] * int(self.sizes[b])
            rows += list(block)
        out["bs_idx"] = self._batch_rows(self.bs_apply)
        out["bs_meta"] = np.asarray(meta, dtype=np.int32).reshape(-1, 3)
        out["bs_taps"] = np.stack(rows) if rows else np.zeros((0, t), np.float32)
        ratios = sorted({r for r in self.ps_shift if r is not None})
        per_ratio = [self._batch_rows(np.asarray([s == r for s in self.ps_shift], dtype=bool)) for r in ratios]
        out["ps_idx"] = np.concatenate(per_ratio).astype(np.int32) if per_ratio else np.zeros(0, np.int32)
        out["ps_ratios"] = np.asarray([(r.numerator, r.denominator) for r in ratios], dtype=np.int64).reshape(-1, 2)
        out["ps_counts"] = np.asarray([len(rows) for rows in per_ratio], dtype=np.int64)
        return out


PITCH_N_FFT_DIV, PITCH_HOP_DIV = 64, 32          # torch_pitch_shift: n_fft = sample_rate // 64, hop_length = n_fft // 32
RESAMPLE_LOWPASS_WIDTH, RESAMPLE_ROLLOFF = 6, 0.99  # torchaudio.transforms.Resample defaults (sinc_interp_hann)


def sinc_resample_kernel(orig: int, new: int) -> Tuple[np.ndarray, int]:
    """
    ``torchaudio.functional.functional._get_sinc_resample_kernel(orig, new, gcd=1)`` restated (sinc_interp_hann, width 6, roll-off
    0.99): f32 ``[new, 2 width + orig]`` and ``width``.  The phase term ``-i / new`` is an int64 tensor divided by an int, i.e.
    float32, before it meets the float64 tap index -- kept, it moves the kernel by 1e-8 (pinned by tests/test_k9.py).
    """
    base_freq = min(orig, new) * RESAMPLE_ROLLOFF
    width = int(np.ceil(RESAMPLE_LOWPASS_WIDTH * orig / base_freq))
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)[:, None]
    t = (phase + idx) * base_freq
    t = np.clip(t, -RESAMPLE_LOWPASS_WIDTH, RESAMPLE_LOWPASS_WIDTH)
    window = np.cos(t * np.pi / RESAMPLE_LOWPASS_WIDTH / 2.0) ** 2
    t = t * np.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        kernels = np.where(t == 0, 1.0, np.sin(t) / t) * window * (base_freq / orig)
    return kernels.astype(np.float32), width


def pitch_tables(target_samples: int, shift: Fraction, sample_rate: int = spec.SAMPLE_RATE) -> Dict[str, object]:
    """
    The host-computed tables of one pitch ratio (``hb_pitch_plan_create``).  The vocoder's time steps are taken from
    ``torch.arange(0, frames, rate, dtype=float32)`` itself, as ``torchaudio.functional.phase_vocoder`` does: its float32 values
    decide which frame pair feeds an output frame, and they are not the correctly rounded ``i * rate`` (vectorised evaluation).
    """
    import math

    import torch

    n_fft = sample_rate // PITCH_N_FFT_DIV
    hop = n_fft // PITCH_HOP_DIV
    frames = 1 + target_samples // hop
    rate = float(1 / shift)
    steps = torch.arange(0, frames, rate, dtype=torch.float32)
    new_rate = int(sample_rate / shift)
    g = math.gcd(sample_rate, new_rate)
    kernel, width = sinc_resample_kernel(sample_rate // g, new_rate // g)
    return {"n_fft": n_fft, "hop": hop, "frames_in": frames, "frames_out": int(steps.shape[0]),
            "idx0": steps.long().numpy().astype(np.int32), "idx1": (steps + 1).long().numpy().astype(np.int32),
            "alpha": (steps % 1.0).numpy().astype(np.float32),
            "phase_advance": torch.linspace(0, math.pi * hop, n_fft // 2 + 1).numpy().astype(np.float32),
            "orig": sample_rate // g, "up": new_rate // g, "width": width, "kernel": np.ascontiguousarray(kernel)}


class _PitchPlan:
    """Device-side plan of one (clip length, ratio, device): owns the ``hb_pitch_plan`` handle."""

    def __init__(self, target_samples: int, shift: Fraction, device) -> None:
        import ctypes

        import torch

        from heybuddy_b200 import _native

        self.lib = _native.load()
        self.tables = pitch_tables(target_samples, shift)
        t = self.tables
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _native.check(self.lib.hb_pitch_plan_create(
                ctypes.byref(handle), int(target_samples), t["n_fft"], t["hop"], t["frames_out"], t["idx0"].ctypes.data, t["idx1"].ctypes.data,
                t["alpha"].ctypes.data, t["phase_advance"].ctypes.data, t["orig"], t["up"], t["width"], t["kernel"].ctypes.data),
                "hb_pitch_plan_create")
        self.handle = handle

    def workspace_bytes(self, k: int) -> int:
        from heybuddy_b200 import _native

        n = self.lib.hb_k9_pitch_workspace_bytes(self.handle, int(k))
        _native.check(n, "hb_k9_pitch_workspace_bytes")
        return int(n)

    def __del__(self):
        try:
            self.lib.hb_pitch_plan_destroy(self.handle)
        except Exception:
            pass


_PITCH_PLANS: Dict[Tuple[int, Fraction, int], _PitchPlan] = {}


def pitch_plan(target_samples: int, shift: Fraction, device) -> _PitchPlan:
    import torch

    key = (int(target_samples), Fraction(shift), torch.device(device).index or 0)
    if key not in _PITCH_PLANS:
        _PITCH_PLANS[key] = _PitchPlan(target_samples, shift, device)
    return _PITCH_PLANS[key]


def apply_packed(fixed, pk: Dict[str, object], scratch) -> None:
    """
    Runs the packed K9 draws (``K9Draws.pack()`` with the arrays on ``fixed``'s device; ``ps_ratios`` stays a host array) in place
    on cuda f32 ``[n, T]`` length-fixed clips, in the reference's order: the per-clip numpy transforms (EQ, then distortion:
    augmented.py:325-328), then PitchShift and BandStopFilter (the head of the batch Compose, :369-372).
    ``scratch(name, numel, dtype)`` returns a device buffer of at least ``numel`` elements.
    """
    import torch

    from heybuddy_b200 import _native

    lib = _native.load()
    n, t = fixed.shape
    dev = fixed.device
    with torch.cuda.device(dev):
        st = _native.stream_ptr(dev)
        if pk["eq_idx"].numel():
            _native.check(lib.hb_k9_eq_f32(fixed.data_ptr(), pk["eq_idx"].data_ptr(), pk["eq_sos"].data_ptr(), int(pk["eq_idx"].numel()), t, st), "hb_k9_eq_f32")
        if pk["th_idx"].numel():
            _native.check(lib.hb_k9_tanh_f32(fixed.data_ptr(), pk["th_idx"].data_ptr(), pk["th_amt"].data_ptr(), int(pk["th_idx"].numel()), t, st), "hb_k9_tanh_f32")
        if pk["ps_idx"].numel():
            ratios, counts = pk["ps_ratios"], pk["ps_counts"]
            at = 0
            for (num, den), cnt in zip(ratios, counts):
                plan = pitch_plan(t, Fraction(int(num), int(den)), dev)
                nbytes = plan.workspace_bytes(int(cnt))
                ws = scratch("ps_ws", nbytes, torch.uint8)
                idx = pk["ps_idx"][at:at + int(cnt)]
                _native.check(lib.hb_k9_pitch_f32(plan.handle, fixed.data_ptr(), idx.data_ptr(), int(cnt), ws.data_ptr(), int(ws.numel()), st), "hb_k9_pitch_f32")
                at += int(cnt)
        if pk["bs_idx"].numel():
            kb, rows = int(pk["bs_idx"].numel()), int(pk["bs_taps"].shape[0])
            specs = scratch("bs_specs", rows * (t // 2 + 1) * 2, torch.float32)
            _native.check(lib.hb_rir_spectrum(pk["bs_taps"].data_ptr(), specs.data_ptr(), rows, t, st), "hb_rir_spectrum")
            copy = scratch("bs_copy", kb * t, torch.float32)
            _native.check(lib.hb_k9_bandstop_f32(fixed.data_ptr(), pk["bs_idx"].data_ptr(), pk["bs_meta"].data_ptr(), specs.data_ptr(), copy.data_ptr(),
                                                 kb, t, st), "hb_k9_bandstop_f32")


HOST_ONLY = ("ps_ratios", "ps_counts")        # entries of K9Draws.pack() that are not uploaded


def apply_device(fixed, table, gen=None):
    """Runs the table's K9 transforms in place on cuda f32 ``[n, T]`` length-fixed clips (synchronous uploads: tests, AugmentedAudioGenerator)."""
    import torch

    k9: Optional[K9Draws] = table.k9
    if k9 is None:
        return fixed
    dev = fixed.device
    pk = {name: (arr if name in HOST_ONLY else torch.from_numpy(np.ascontiguousarray(arr)).to(dev)) for name, arr in k9.pack().items()}
    bufs: Dict[str, object] = {}

    def scratch(name, numel, dtype):
        bufs[name] = torch.empty(max(int(numel), 1), dtype=dtype, device=dev)
        return bufs[name]

    apply_packed(fixed, pk, scratch)
    torch.cuda.current_stream(dev).synchronize()      # the scratch buffers die with this frame
    return fixed
