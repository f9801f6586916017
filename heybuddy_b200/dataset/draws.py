"""
Seeded augmentation draw table (counter-based, vectorised).

The reference's augmentation RNG is unseeded (np.random / torch.rand / dataset.shuffle(),
SURVEY.md 0.6, A.3 item 7), so "identical inputs" is defined by a *draw table*: for every
augmentation batch g (clips [g*B, (g+1)*B)) the random choices the reference would make, in the
reference's call order (augmented.py:314-392):

    per-clip pad offsets (augmented.py:222)            -> pad_before[n]
    [per transform: apply coin, then parameters] in Compose order
        pitch shift, band stop                         (K9 rows, see dataset/augmented.py)
        coloured noise: coin, snr, f_decay, N(0,1)[16000] pattern (augmented.py:107-115)
        gain: coin, dB ~ U(-18, 6)                     (augmented.py:116-120)
    background coin (augmented.py:383)                 -> noise stream advance -> rand(B) SNRs (:269-270)
    reverb coin (augmented.py:387)                     -> RIR advance

Every draw is one output of a counter-based generator -- Philox4x32-10 with key = seed and counter =
(index, stream, batch g) -- so the table of ANY range of batches is a handful of numpy vector operations
(no per-clip or per-batch Python), does not depend on how batches are sharded over ranks, and the bulky
draws (the 16000-sample N(0,1) pattern of a coloured-noise batch) are regenerated on the device from the
same counters by ``hb_colored_bases`` instead of being shipped over PCIe.  The stateful cursors (noise
stream position, RIR index) are prefix sums over the coins.  The oracle (tests) and the CUDA kernels consume
the same table.

Streams (counter word 1): 0 batch-level scalars, 1 pad offsets, 2 per-clip noise SNRs, 3 coloured pattern,
4 tanh distortion (per clip), 5 seven-band EQ coin (per clip), 6 seven-band EQ gains (per clip x band).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from heybuddy_b200 import spec
from heybuddy_b200.constants import *  # noqa: F401,F403

__all__ = ["AugmentConfig", "BatchDraw", "DrawTable", "philox4x32", "uniform53", "uniform24", "gaussian_pattern",
           "colored_noise_base", "pad_before_from_uniform", "advance_noise_cursor", "batch_coins",
           "STREAM_BATCH", "STREAM_PAD", "STREAM_SNR", "STREAM_PATTERN", "STREAM_TANH", "STREAM_EQ_COIN", "STREAM_EQ_GAIN"]

STREAM_BATCH, STREAM_PAD, STREAM_SNR, STREAM_PATTERN, STREAM_TANH, STREAM_EQ_COIN, STREAM_EQ_GAIN = range(7)
# indices of the batch-level scalars inside stream 0
(U_COLORED_COIN, U_COLORED_SNR, U_COLORED_FDECAY, U_GAIN_COIN, U_GAIN_DB, U_BG_COIN, U_REVERB_COIN,
 U_PITCH_COIN, U_PITCH_SEMITONES, U_BANDSTOP_COIN, U_BANDSTOP_CENTER, U_BANDSTOP_WIDTH) = range(12)
N_BATCH_SCALARS = 12

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32(c0, c1, c2, c3, seed: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """
    Philox4x32-10 (Salmon et al., SC'11; the generator behind cuRAND's Philox): counter words ``c0..c3`` (broadcastable
    integer arrays < 2**32), key = the low / high 32 bits of ``seed``.  Returns the four 32-bit output words as uint64 arrays.
    ``csrc/draws.cu`` evaluates the same function on the device.
    """
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)))
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = (p1 >> _S32) ^ c1 ^ np.uint64(k0), p1 & _MASK, (p0 >> _S32) ^ c3 ^ np.uint64(k1), p0 & _MASK
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniform53(x0: np.ndarray, x1: np.ndarray) -> np.ndarray:
    """Two 32-bit words -> float64 in [0, 1) with 53 random bits (the numpy / C++ ``generate_canonical`` construction)."""
    return ((x0 >> np.uint64(5)).astype(np.float64) * 67108864.0 + (x1 >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


def uniform24(x: np.ndarray) -> np.ndarray:
    """One 32-bit word -> float32 in [0, 1) with 24 random bits (what ``torch.rand`` produces for float32)."""
    return ((x >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def gaussian_pattern(seed: int, batch: int, n: int = spec.COLORED_BASE_SAMPLES) -> np.ndarray:
    """
    The N(0,1) pattern of coloured-noise batch ``batch`` (float64): counter j of stream 3 gives samples 4j..4j+3 through two
    Box-Muller pairs, u1 = ((x >> 9) + 0.5) / 2**23 in (0, 1) and u2 = (x >> 8) / 2**24 in [0, 1) -- both exactly representable
    in float32, so the device (fp32 logf / sincospif) and this float64 restatement start from identical uniforms.
    """
    assert n % 4 == 0
    j = np.arange(n // 4, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32(j, STREAM_PATTERN, int(batch) & 0xFFFFFFFF, int(batch) >> 32, seed)
    out = np.empty((n // 4, 4), dtype=np.float64)
    for col, (a, b) in enumerate(((x0, x1), (x2, x3))):
        u1 = ((a >> np.uint64(9)).astype(np.float64) + 0.5) / 8388608.0
        u2 = (b >> np.uint64(8)).astype(np.float64) / 16777216.0
        r = np.sqrt(-2.0 * np.log(u1))
        out[:, 2 * col] = r * np.cos(2.0 * np.pi * u2)
        out[:, 2 * col + 1] = r * np.sin(2.0 * np.pi * u2)
    return out.reshape(-1)


def colored_noise_base(gauss: np.ndarray, f_decay: float) -> np.ndarray:
    """
    torch_audiomentations ``_gen_noise``: rfft of the 1 s N(0,1) pattern,
    ``1/linspace(1, sqrt(sr/2), bins)**f_decay`` mask, irfft, unit RMS.  float64 math, f32 result.
    Host-side restatement of what ``hb_colored_bases`` computes on the device (used by ``BatchDraw.colored_base``,
    i.e. by tests and by the per-batch reference surface; the fused path never ships a pattern over PCIe).
    """
    g = np.asarray(gauss, dtype=np.float64)
    sr = g.shape[0]
    s = np.fft.rfft(g)
    mask = 1.0 / (np.linspace(1.0, (sr / 2) ** 0.5, s.shape[0], dtype=np.float64) ** float(f_decay))
    c = np.fft.irfft(s * mask, n=sr)
    return (c / np.sqrt(np.mean(c * c))).astype(np.float32)


def pad_before_from_uniform(lengths: np.ndarray, target: int, u: np.ndarray) -> np.ndarray:
    """
    augmented.py:216-226 with the draw supplied as a uniform: ``randint(int(s/4), int(3s/4))`` (high exclusive) for s = missing
    samples; s == 1 pads right (0 before); s <= 0 no pad.  ``lo + floor(u * (hi - lo))``, vectorised.
    """
    s = np.maximum(target - np.asarray(lengths, dtype=np.int64), 0)
    lo, hi = s // 4, (3 * s) // 4
    span = np.maximum(hi - lo, 0)
    pad = lo + np.minimum(np.floor(u * span).astype(np.int64), np.maximum(span - 1, 0))
    return np.where(s <= 1, 0, pad).astype(np.int32)


@dataclass
class AugmentConfig:
    """Probabilities / ranges of the batch transforms (defaults: reference constants.py)."""
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
    # K9: per-clip numpy transforms (augmented.py:79-90) and the two remaining batch transforms (:93-106)
    seven_band_prob: float = 0.0
    seven_band_gain_db: float = DEFAULT_AUGMENT_SEVEN_BAND_GAIN_DB
    tanh_distortion_prob: float = 0.0
    tanh_min_distortion: float = DEFAULT_AUGMENT_TANH_MIN_DISTORTION
    tanh_max_distortion: float = DEFAULT_AUGMENT_TANH_MAX_DISTORTION
    pitch_shift_prob: float = 0.0
    pitch_shift_semitones: int = DEFAULT_AUGMENT_PITCH_SHIFT_SEMITONES
    band_stop_prob: float = 0.0


def batch_scalars(seed: int, batch_ids: np.ndarray) -> np.ndarray:
    """The N_BATCH_SCALARS float64 uniforms of every batch in ``batch_ids`` -> ``[n_batches, N_BATCH_SCALARS]``."""
    g = np.asarray(batch_ids, dtype=np.uint64)[:, None]
    j = np.arange(N_BATCH_SCALARS, dtype=np.uint64)[None, :]
    x0, x1, _, _ = philox4x32(j, STREAM_BATCH, g & _MASK, g >> _S32, seed)
    return uniform53(x0, x1)


def batch_coins(seed: int, batch_ids: np.ndarray, cfg: AugmentConfig, have_background: bool, have_reverb: bool) -> Tuple[np.ndarray, np.ndarray]:
    """(background applied, reverb applied) for every batch in ``batch_ids`` -- all the cursor prefix sums need."""
    u = batch_scalars(seed, batch_ids)
    bg = (u[:, U_BG_COIN] < cfg.background_noise_prob) & bool(have_background)
    rev = (u[:, U_REVERB_COIN] < cfg.reverb_prob) & bool(have_reverb)
    return bg, rev


def advance_noise_cursor(cursor: int, need: int, starts: np.ndarray) -> int:
    """
    augmented.py:246-251: whole noise clips are pulled, starting at clip ``cursor``, until at least ``need`` samples are
    there.  ``starts`` = cumulative clip starts ``[n_clips + 1]``.  Returns the cursor after the pull (mod n_clips).
    """
    n = starts.shape[0] - 1
    total = int(starts[n])
    cycles, rem = divmod(need - 1, total)
    rem += 1                                    # 1 <= rem <= total samples still missing after `cycles` whole passes
    # partial sums from `cursor`, wrapping: S_m = samples of clips cursor .. cursor + m - 1
    head = int(starts[n] - starts[cursor])      # samples up to the end of the bank
    if rem <= head:
        m = int(np.searchsorted(starts[cursor + 1:], starts[cursor] + rem, side="left")) + 1
    else:
        m = (n - cursor) + int(np.searchsorted(starts[1:], rem - head, side="left")) + 1
    return (cursor + cycles * n + m) % n


class BatchDraw:
    """Every random choice of one augmentation batch (a view into a :class:`DrawTable`)."""
    __slots__ = ("table", "k")

    def __init__(self, table: "DrawTable", k: int) -> None:
        self.table, self.k = table, k

    def _rows(self) -> slice:
        t = self.table
        return slice(int(t.starts[self.k]), int(t.starts[self.k + 1]))

    index = property(lambda s: int(s.table.first_batch + s.k))
    pad_before = property(lambda s: s.table.pad_before[s._rows()])
    colored_apply = property(lambda s: bool(s.table.colored_apply[s.k]))
    colored_snr_db = property(lambda s: float(s.table.colored_snr_db[s.k]))
    colored_f_decay = property(lambda s: float(s.table.colored_f_decay[s.k]))
    gain_apply = property(lambda s: bool(s.table.gain_apply[s.k]))
    gain_db = property(lambda s: float(s.table.gain_db[s.k]))
    background_apply = property(lambda s: bool(s.table.background_apply[s.k]))
    reverb_apply = property(lambda s: bool(s.table.reverb_apply[s.k]))

    @property
    def noise_snr_db(self) -> Optional[np.ndarray]:
        return self.table.noise_snr_db[self._rows()] if self.background_apply else None

    @property
    def gain_linear(self) -> float:
        return float(self.table.gain_linear[self.k])

    @property
    def colored_base(self) -> Optional[np.ndarray]:
        """f32 [16000] unit-RMS pattern (host float64 restatement of ``hb_colored_bases``); None when not applied."""
        if not self.colored_apply:
            return None
        return colored_noise_base(gaussian_pattern(self.table.seed, self.index), self.colored_f_decay)


class DrawTable:
    """
    Draws of consecutive batches ``first_batch .. first_batch + n_batches - 1`` as struct-of-arrays, plus the stateful
    cursors (first noise-bank clip of the batch, RIR of the batch; -1 = not applied).
    """

    def __init__(self, cfg: AugmentConfig, seed: int) -> None:
        self.cfg, self.seed = cfg, int(seed)
        self.first_batch = 0
        self.sizes = np.zeros(0, np.int64)
        self.starts = np.zeros(1, np.int64)
        self.pad_before = np.zeros(0, np.int32)
        self.noise_snr_db = np.zeros(0, np.float32)
        for name in ("colored_apply", "gain_apply", "background_apply", "reverb_apply"):
            setattr(self, name, np.zeros(0, bool))
        for name in ("colored_snr_db", "colored_f_decay", "gain_db", "gain_linear"):
            setattr(self, name, np.zeros(0, np.float64))
        self.noise_clip_cursor: List[int] = []
        self.rir_index: List[int] = []
        self.final_noise_cursor = 0
        self.final_rir_cursor = 0
        self.k9 = None      # optional K9 draws (dataset/k9.py)

    # -- views ---------------------------------------------------------------------------------------------
    @property
    def n_batches(self) -> int:
        return int(self.sizes.shape[0])

    @property
    def n_clips(self) -> int:
        return int(self.starts[-1])

    @property
    def batches(self) -> List[BatchDraw]:
        return [BatchDraw(self, k) for k in range(self.n_batches)]

    def slice(self, b0: int, b1: int) -> "DrawTable":
        """Batches [b0, b1) of this table (views, no copies)."""
        t = DrawTable(self.cfg, self.seed)
        t.first_batch = self.first_batch + b0
        t.sizes = self.sizes[b0:b1]
        t.starts = self.starts[b0:b1 + 1] - self.starts[b0]
        r0, r1 = int(self.starts[b0]), int(self.starts[b1])
        t.pad_before, t.noise_snr_db = self.pad_before[r0:r1], self.noise_snr_db[r0:r1]
        for name in ("colored_apply", "gain_apply", "background_apply", "reverb_apply", "colored_snr_db", "colored_f_decay",
                     "gain_db", "gain_linear"):
            setattr(t, name, getattr(self, name)[b0:b1])
        t.noise_clip_cursor, t.rir_index = self.noise_clip_cursor[b0:b1], self.rir_index[b0:b1]
        t.k9 = self.k9.slice(b0, b1, r0, r1) if self.k9 is not None else None
        return t

    # -- construction ----------------------------------------------------------------------------------------
    @classmethod
    def build(cls, lengths: Sequence[int], cfg: AugmentConfig, seed: int, noise_clip_lengths: Optional[np.ndarray] = None,
              num_rirs: int = 0, first_batch: int = 0, noise_cursor: int = 0, rir_cursor: int = 0) -> "DrawTable":
        """
        ``lengths``: source clip lengths in order.  ``noise_clip_lengths``: lengths of the bank's clips in
        stream order (the reference pulls whole clips until >= B*T samples, augmented.py:246-251, and drops the
        unused tail); ``num_rirs``: size of the RIR bank (one RIR per applied batch, wrapping).
        """
        t = cls(cfg, seed)
        lengths = np.asarray(lengths, dtype=np.int64).reshape(-1)
        n, b = int(lengths.shape[0]), int(cfg.batch_size)
        nb = (n + b - 1) // b
        t.first_batch = int(first_batch)
        t.sizes = np.full(nb, b, dtype=np.int64)
        if nb:
            t.sizes[-1] = n - (nb - 1) * b
        t.starts = np.concatenate(([0], np.cumsum(t.sizes))).astype(np.int64)
        have_bg = noise_clip_lengths is not None and len(noise_clip_lengths) > 0
        have_rev = num_rirs > 0
        gids = np.arange(first_batch, first_batch + nb, dtype=np.uint64)
        u = batch_scalars(seed, gids)
        lin = lambda col, lo, hi: lo + u[:, col] * (hi - lo)
        t.colored_apply = u[:, U_COLORED_COIN] < cfg.colored_noise_prob
        t.colored_snr_db = np.where(t.colored_apply, lin(U_COLORED_SNR, cfg.colored_noise_min_snr_db, cfg.colored_noise_max_snr_db), 0.0)
        t.colored_f_decay = np.where(t.colored_apply, lin(U_COLORED_FDECAY, cfg.colored_noise_min_f_decay, cfg.colored_noise_max_f_decay), 0.0)
        t.gain_apply = u[:, U_GAIN_COIN] < cfg.gain_prob
        t.gain_db = np.where(t.gain_apply, lin(U_GAIN_DB, cfg.gain_min_db, cfg.gain_max_db), 0.0)
        t.gain_linear = np.where(t.gain_apply, 10.0 ** (t.gain_db / 20.0), 1.0)
        t.background_apply = (u[:, U_BG_COIN] < cfg.background_noise_prob) & have_bg
        t.reverb_apply = (u[:, U_REVERB_COIN] < cfg.reverb_prob) & have_rev
        # per-clip draws: clip i of batch g is counter (i, stream, g)
        within = np.arange(n, dtype=np.uint64) - np.repeat(t.starts[:-1], t.sizes).astype(np.uint64)
        g_of = np.repeat(gids, t.sizes)
        x0, x1, _, _ = philox4x32(within, STREAM_PAD, g_of & _MASK, g_of >> _S32, seed)
        t.pad_before = pad_before_from_uniform(lengths, cfg.target_samples, uniform53(x0, x1))
        s0, _, _, _ = philox4x32(within, STREAM_SNR, g_of & _MASK, g_of >> _S32, seed)
        span = np.float32(cfg.background_noise_max_snr_db - cfg.background_noise_min_snr_db)
        snr = (uniform24(s0) * span + np.float32(cfg.background_noise_min_snr_db)).astype(np.float32)   # torch.rand(B) * span + min, f32 (:269-270)
        t.noise_snr_db = np.where(np.repeat(t.background_apply, t.sizes), snr, np.float32(0.0)).astype(np.float32)
        # stateful cursors: prefix over the coins
        t.noise_clip_cursor = [-1] * nb
        if have_bg and t.background_apply.any():
            ncl = np.asarray(noise_clip_lengths, dtype=np.int64)
            starts = np.concatenate(([0], np.cumsum(ncl))).astype(np.int64)
            noise_cursor %= len(ncl)
            for k in np.nonzero(t.background_apply)[0]:
                t.noise_clip_cursor[k] = int(noise_cursor)
                noise_cursor = advance_noise_cursor(int(noise_cursor), int(t.sizes[k]) * cfg.target_samples, starts)
        rev_rank = np.cumsum(t.reverb_apply) - 1
        t.rir_index = [int((rir_cursor + r) % num_rirs) if a else -1 for a, r in zip(t.reverb_apply, rev_rank)] if have_rev else [-1] * nb
        t.final_noise_cursor = int(noise_cursor)
        t.final_rir_cursor = int(rir_cursor + int(t.reverb_apply.sum()))
        if cfg.seven_band_prob or cfg.tanh_distortion_prob or cfg.pitch_shift_prob or cfg.band_stop_prob:
            from heybuddy_b200.dataset.k9 import K9Draws

            t.k9 = K9Draws.build(cfg, seed, gids, t.sizes, within, g_of, u)
        return t

    # -- packing for the device ------------------------------------------------------------------------------
    def colored_slots(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(slot of every batch or -1, global batch ids of the coloured batches, their f_decay as f32)."""
        slot = np.where(self.colored_apply, np.cumsum(self.colored_apply) - 1, -1).astype(np.int32)
        which = np.nonzero(self.colored_apply)[0]
        return slot, (self.first_batch + which).astype(np.int64), self.colored_f_decay[which].astype(np.float32)

    def clip_records(self, record_dtype: np.dtype, noise_clip_starts: Optional[np.ndarray], noise_stream_len: int) -> np.ndarray:
        """Per-clip ``hb_clip_aug`` records of the whole table in one vectorised pass."""
        n, t = self.n_clips, self.cfg.target_samples
        r = np.zeros(n, dtype=record_dtype)
        if n == 0:
            return r
        rep = lambda v, dt: np.repeat(np.asarray(v, dtype=dt), self.sizes)
        slot, _, _ = self.colored_slots()
        r["gain"] = rep(self.gain_linear, np.float32)
        r["colored_index"] = rep(slot, np.int32)
        r["colored_snr_db"] = rep(self.colored_snr_db, np.float32)
        r["rir_index"] = rep(np.asarray(self.rir_index, dtype=np.int64), np.int32)
        cursors = np.asarray(self.noise_clip_cursor, dtype=np.int64)
        has_bg = self.background_apply & (cursors >= 0)
        if has_bg.any():
            base = np.where(has_bg, np.asarray(noise_clip_starts, dtype=np.int64)[np.maximum(cursors, 0)], 0)
            if np.any(has_bg & (base + self.sizes * t > noise_stream_len)):
                raise ValueError("noise bank wrap margin too small for this batch size")
            within = np.arange(n, dtype=np.int64) - np.repeat(self.starts[:-1], self.sizes)
            r["noise_offset"] = np.where(np.repeat(has_bg, self.sizes), np.repeat(base, self.sizes) + within * t, -1)
            r["noise_snr_db"] = self.noise_snr_db
        else:
            r["noise_offset"] = -1
        return r
