"""
``AugmentedAudioGenerator`` -- B200 replacement of reference ``heybuddy/dataset/augmented.py``.

Same constructor / ``execute_augment_batch`` / ``__call__`` surface (augmented.py:25-52, 297, 396).
The north-star transforms -- coloured/white noise, gain, background noise at a target SNR, RIR
reverb -- run in ONE fused CUDA kernel per clip (``hb_augment_clips_f32``) with the noise bank and
the RIR spectra resident in HBM; the length fix runs in ``hb_fix_length_i16``.  The K9 transforms
(augmented.py:79-106) run on the device between the length fix and that kernel, in the
reference's order: the per-clip numpy ones (SevenBandParametricEQ, TanhDistortion), then the head of the batch
Compose (PitchShift, BandStopFilter) -- ``heybuddy_b200/dataset/k9.py``.  Their probabilities default to 0 here
(BASELINE configs[1] has none of them; the reference's constants are 0.25 each) and all four libraries are absent
offline, so their arithmetic is restated from published behaviour (parity unpinned).

Randomness comes from the seeded draw table (``heybuddy_b200.dataset.draws``), generated in the
reference's call order; with ``seed=None`` a fresh seed is drawn (the reference is unseeded).
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Optional, Sequence, Union

import numpy as np

from heybuddy_b200 import _native, spec
from heybuddy_b200.constants import *  # noqa: F401,F403
from heybuddy_b200.dataset.draws import AugmentConfig, BatchDraw, DrawTable
from heybuddy_b200.util import logger

__all__ = ["AugmentedAudioGenerator", "NoiseBank", "RirBank", "rotate_rir"]


def _audio_array(item: Any) -> np.ndarray:
    """Dataset row -> waveform (accepts {"audio": {"array"}}, {"array"}, arrays and python lists, augmented.py:278-295)."""
    if isinstance(item, dict):
        if "audio" in item:
            item = item["audio"]
        if isinstance(item, dict):
            sr = int(item.get("sampling_rate", spec.SAMPLE_RATE))
            if sr != spec.SAMPLE_RATE:
                raise ValueError(f"resampling ({sr} Hz) is outside the B200 hot path; provide 16 kHz audio")
            item = item["array"]
    if isinstance(item, list):
        if not item:
            raise ValueError("Audio list is empty")
        first = item[0][0] if isinstance(item[0], list) else item[0]
        return np.array(item, dtype=np.float32 if isinstance(first, float) else np.int16)
    return np.asarray(item)


def rotate_rir(rir: np.ndarray, t: int) -> np.ndarray:
    """
    speechbrain ``convolve1d(use_fft=True, rotation_index=argmax|rir|)``: the RIR truncated to ``t``
    samples and rotated so its direct path sits at lag 0: ``[rir[d:], zeros, rir[:d]]``.
    """
    rir = np.asarray(rir, dtype=np.float32).reshape(-1)
    d = int(np.argmax(np.abs(rir)))
    if rir.shape[0] > t:
        rir = rir[:t]
    d = min(d, rir.shape[0])
    k = np.zeros(t, dtype=np.float32)
    k[:rir.shape[0] - d] = rir[d:]
    if d:
        k[t - d:] = rir[:d]
    return k


class NoiseBank:
    """
    Background-noise clips resident in HBM as one contiguous stream (the reference concatenates
    whole clips until it has B*T samples, augmented.py:246-257, so a batch's rows are consecutive
    slices of the stream).  A wrap margin (copy of the head) makes a batch that runs past the end
    addressable without a gather.
    """

    def __init__(self, clips: Union[np.ndarray, Sequence[Any]], device, margin_samples: int) -> None:
        import torch

        if isinstance(clips, np.ndarray) and clips.ndim == 2:
            arrays = None
            self.clip_lengths = np.full(clips.shape[0], clips.shape[1], dtype=np.int64)
            flat = np.ascontiguousarray(clips, dtype=np.float32).reshape(-1)
        else:
            arrays = [_audio_array(c).astype(np.float32).reshape(-1) for c in clips]
            self.clip_lengths = np.array([a.shape[0] for a in arrays], dtype=np.int64)
            flat = np.concatenate(arrays) if arrays else np.zeros(0, np.float32)
        if flat.size == 0:
            raise ValueError("empty noise bank")
        self.clip_starts = np.concatenate([[0], np.cumsum(self.clip_lengths)]).astype(np.int64)
        self.num_samples = int(flat.size)
        reps = int(np.ceil(margin_samples / flat.size))
        margin = np.tile(flat, reps)[:margin_samples]
        # 16-byte aligned rows need T % 4 == 0 and clip starts % 4 == 0; the kernel falls back to scalar loads otherwise
        self.stream = torch.from_numpy(np.concatenate([flat, margin])).to(device)

    def __len__(self) -> int:
        return len(self.clip_lengths)

    def offset_of_clip(self, clip_index: int) -> int:
        return int(self.clip_starts[clip_index % len(self.clip_lengths)])


class RirBank:
    """Room impulse responses: rotated time-domain kernels -> device spectra (``hb_rir_spectrum``), once."""

    def __init__(self, rirs: Sequence[Any], device, t: int = spec.CLIP_SAMPLES) -> None:
        import torch

        self.kernels_host = np.stack([rotate_rir(_audio_array(r).astype(np.float32), t) for r in rirs])
        self.t = t
        n = self.kernels_host.shape[0]
        kernels = torch.from_numpy(self.kernels_host).to(device)
        self.spec = torch.empty((n, t // 2 + 1, 2), dtype=torch.float32, device=device)
        lib = _native.load()
        with torch.cuda.device(device):
            _native.check(lib.hb_rir_spectrum(kernels.data_ptr(), self.spec.data_ptr(), n, t, _native.stream_ptr(device)),
                          "hb_rir_spectrum")
            torch.cuda.current_stream(device).synchronize()

    def __len__(self) -> int:
        return int(self.spec.shape[0])


class AugmentedAudioGenerator:
    """A generator that yields augmented audio samples (reference augmented.py:16-427)."""

    def __init__(
        self,
        source_dataset: Any,
        device_id: Optional[int] = None,
        augmentation_dataset: Optional[Any] = None,
        impulse_response_dataset: Optional[Any] = None,
        target_length: float = 1.44,
        sample_rate: int = 16000,
        batch_size: int = 128,
        seven_band_aug_prob: float = 0.0,
        seven_band_aug_gain_db: float = DEFAULT_AUGMENT_SEVEN_BAND_GAIN_DB,
        tanh_distortion_prob: float = 0.0,
        tanh_min_distortion: float = DEFAULT_AUGMENT_TANH_MIN_DISTORTION,
        tanh_max_distortion: float = DEFAULT_AUGMENT_TANH_MAX_DISTORTION,
        pitch_shift_prob: float = 0.0,
        pitch_shift_semitones: int = DEFAULT_AUGMENT_PITCH_SHIFT_SEMITONES,
        band_stop_prob: float = 0.0,
        colored_noise_prob: float = DEFAULT_AUGMENT_COLORED_NOISE_PROB,
        colored_noise_min_snr_db: float = DEFAULT_AUGMENT_COLORED_NOISE_MIN_SNR_DB,
        colored_noise_max_snr_db: float = DEFAULT_AUGMENT_COLORED_NOISE_MAX_SNR_DB,
        colored_noise_min_f_decay: float = DEFAULT_AUGMENT_COLORED_NOISE_MIN_F_DECAY,
        colored_noise_max_f_decay: float = DEFAULT_AUGMENT_COLORED_NOISE_MAX_F_DECAY,
        background_noise_prob: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_PROB,
        background_noise_min_snr_db: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_MIN_SNR_DB,
        background_noise_max_snr_db: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_MAX_SNR_DB,
        gain_prob: float = DEFAULT_AUGMENT_GAIN_PROB,
        reverb_prob: float = DEFAULT_AUGMENT_REVERB_PROB,
        seed: Optional[int] = None,
        first_batch: int = 0,
    ) -> None:
        if sample_rate != spec.SAMPLE_RATE:
            raise ValueError("the B200 hot path is 16 kHz only")
        self.device_id = device_id
        self.source_dataset = source_dataset
        self.target_length = target_length
        self.sample_rate = sample_rate
        self.batch_size = batch_size
        self.reverb_prob = reverb_prob
        self.background_noise_prob = background_noise_prob
        self.background_noise_min_snr_db = background_noise_min_snr_db
        self.background_noise_max_snr_db = background_noise_max_snr_db
        # The reference raises when a probability is set but the dataset is missing (augmented.py:73-76),
        # which makes its own bare-constructor test unrunnable; degrade to "not applied" instead.
        if background_noise_prob > 0 and augmentation_dataset is None:
            logger.warning("Background noise is enabled but no augmentation dataset is provided; it will not be applied")
        if reverb_prob > 0 and impulse_response_dataset is None:
            logger.warning("Reverb is enabled but no impulse response dataset is provided; it will not be applied")
        self.cfg = AugmentConfig(
            batch_size=batch_size, target_samples=self.target_num_samples,
            colored_noise_prob=colored_noise_prob, colored_noise_min_snr_db=colored_noise_min_snr_db,
            colored_noise_max_snr_db=colored_noise_max_snr_db, colored_noise_min_f_decay=colored_noise_min_f_decay,
            colored_noise_max_f_decay=colored_noise_max_f_decay, gain_prob=gain_prob,
            background_noise_prob=background_noise_prob, background_noise_min_snr_db=background_noise_min_snr_db,
            background_noise_max_snr_db=background_noise_max_snr_db, reverb_prob=reverb_prob,
            seven_band_prob=seven_band_aug_prob, seven_band_gain_db=seven_band_aug_gain_db, tanh_distortion_prob=tanh_distortion_prob,
            tanh_min_distortion=tanh_min_distortion, tanh_max_distortion=tanh_max_distortion,
            pitch_shift_prob=pitch_shift_prob, pitch_shift_semitones=pitch_shift_semitones, band_stop_prob=band_stop_prob)
        self.seed = int(np.random.SeedSequence().entropy % (2 ** 63)) if seed is None else int(seed)
        self._augmentation_dataset = augmentation_dataset
        self._impulse_response_dataset = impulse_response_dataset
        self._noise_bank: Optional[NoiseBank] = augmentation_dataset if isinstance(augmentation_dataset, NoiseBank) else None
        self._rir_bank: Optional[RirBank] = impulse_response_dataset if isinstance(impulse_response_dataset, RirBank) else None
        self._batch_index = first_batch
        self._noise_cursor = 0
        self._rir_cursor = 0
        self._source_iter: Optional[Iterator[Any]] = None

    # -- reference properties ------------------------------------------------------------------
    @property
    def device(self):
        return _native.require_cuda(self.device_id)

    @property
    def target_num_samples(self) -> int:
        return int(self.target_length * self.sample_rate)

    @property
    def noise_bank(self) -> Optional[NoiseBank]:
        if self._noise_bank is None and self._augmentation_dataset is not None:
            self._noise_bank = NoiseBank(self._augmentation_dataset, self.device,
                                         margin_samples=self.batch_size * self.target_num_samples)
        return self._noise_bank

    @property
    def rir_bank(self) -> Optional[RirBank]:
        if self._rir_bank is None and self._impulse_response_dataset is not None:
            self._rir_bank = RirBank(self._impulse_response_dataset, self.device, self.target_num_samples)
        return self._rir_bank

    def get_next_audio_sample_dict(self) -> Dict[str, Any]:
        """Next source row; the source is iterated in order and restarts when exhausted (augmented.py:176-186)."""
        if self._source_iter is None:
            self._source_iter = iter(self.source_dataset)
        try:
            item = next(self._source_iter)
        except StopIteration:
            self._source_iter = iter(self.source_dataset)
            item = next(self._source_iter)
        return item if isinstance(item, dict) and "audio" in item else {"audio": item}

    # -- device path ------------------------------------------------------------------------------
    def fix_length_device(self, clips: Sequence[np.ndarray], pad_before: np.ndarray):
        """Ragged int16 / float clips -> cuda f32 ``[n, T]`` (``hb_fix_length_i16`` for int16 sources)."""
        import torch

        t = self.target_num_samples
        n = len(clips)
        dev = self.device
        if all(c.dtype == np.int16 for c in clips):
            lengths = np.array([c.shape[0] for c in clips], dtype=np.int64)
            offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
            flat = torch.from_numpy(np.concatenate(clips) if n else np.zeros(0, np.int16)).to(dev)
            off_d = torch.from_numpy(offsets).to(dev)
            pad_d = torch.from_numpy(np.ascontiguousarray(pad_before, dtype=np.int32)).to(dev)
            out = torch.empty((n, t), dtype=torch.float32, device=dev)
            lib = _native.load()
            with torch.cuda.device(dev):
                _native.check(lib.hb_fix_length_i16(flat.data_ptr(), off_d.data_ptr(), pad_d.data_ptr(), out.data_ptr(), n, t,
                                                    _native.stream_ptr(dev)), "hb_fix_length_i16")
            return out
        # float sources: same rule on the host staging buffer (one copy, then H2D)
        host = np.zeros((n, t), dtype=np.float32)
        for i, c in enumerate(clips):
            c = c.astype(np.float32) / 32768.0 if c.dtype == np.int16 else c.astype(np.float32)
            if c.shape[0] >= t:
                host[i] = c[:t]
            else:
                host[i, pad_before[i]:pad_before[i] + c.shape[0]] = c
        return torch.from_numpy(host).to(dev)

    def clip_params(self, table: DrawTable) -> np.ndarray:
        """Per-clip ``hb_clip_aug`` records of a whole draw table (numpy structured array), one vectorised pass."""
        nb = self.noise_bank if bool(np.any(table.background_apply)) else None
        return table.clip_records(_native.CLIP_AUG_DTYPE, nb.clip_starts if nb is not None else None,
                                  int(nb.stream.numel()) if nb is not None else 0)

    def colored_bases_device(self, table: DrawTable, device, out=None):
        """
        The coloured-noise patterns of the table's coloured batches, generated ON the device from the table's counters
        (``hb_colored_bases``): cuda f32 ``[k, 16000]`` or None.  Only the k batch ids and f_decay values cross PCIe.
        """
        import torch

        _, ids, f_decay = table.colored_slots()
        k = int(ids.shape[0])
        if k == 0:
            return None
        ids_d = torch.from_numpy(ids).to(device)
        fd_d = torch.from_numpy(f_decay).to(device)
        if out is None:
            out = torch.empty((k, spec.COLORED_BASE_SAMPLES), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _native.check(_native.load().hb_colored_bases(table.seed & (2 ** 64 - 1), ids_d.data_ptr(), fd_d.data_ptr(), k,
                                                          out.data_ptr(), _native.stream_ptr(device)), "hb_colored_bases")
        return out[:k]

    def augment_device(self, fixed, table: DrawTable, out=None):
        """
        cuda f32 ``[n, T]`` length-fixed clips + their draw table -> cuda f32 ``[n, T]`` augmented clips;
        all batches of the table in one launch.
        """
        import torch

        dev = fixed.device
        t = self.target_num_samples
        n = fixed.shape[0]
        params = self.clip_params(table)
        assert params.shape[0] == n, (params.shape, n)
        params_d = torch.from_numpy(params.view(np.uint8).reshape(n, -1)).to(dev)
        bases_d = self.colored_bases_device(table, dev)
        if table.k9 is not None:
            from heybuddy_b200.dataset import k9

            fixed = k9.apply_device(fixed, table, self)
        if out is None:
            out = torch.empty_like(fixed)
        lib = _native.load()
        nb, rb = self.noise_bank, self.rir_bank
        with torch.cuda.device(dev):
            _native.check(
                lib.hb_augment_clips_f32(
                    fixed.data_ptr(), nb.stream.data_ptr() if nb is not None else None,
                    bases_d.data_ptr() if bases_d is not None else None,
                    rb.spec.data_ptr() if rb is not None else None,
                    params_d.data_ptr(), out.data_ptr(), n, t, _native.stream_ptr(dev)),
                "hb_augment_clips_f32")
        return out

    def next_table(self, lengths: Sequence[int]) -> DrawTable:
        """Draws for the next ``ceil(len/B)`` batches, advancing the generator's batch / noise / RIR cursors."""
        nb, rb = self.noise_bank, self.rir_bank
        table = DrawTable.build(
            lengths, self.cfg, self.seed,
            noise_clip_lengths=nb.clip_lengths if nb is not None else None,
            num_rirs=len(rb) if rb is not None else 0,
            first_batch=self._batch_index, noise_cursor=self._noise_cursor, rir_cursor=self._rir_cursor)
        self._batch_index += table.n_batches
        self._noise_cursor = table.final_noise_cursor
        self._rir_cursor = table.final_rir_cursor
        return table

    # -- reference surface ------------------------------------------------------------------------
    def execute_augment_batch(self, batch: Sequence[Any]):
        """Augments one batch of dataset rows -> cuda f32 ``[B, T]`` (augmented.py:297-394)."""
        clips = [_audio_array(a) for a in batch]
        saved = self.cfg.batch_size
        self.cfg.batch_size = max(len(clips), 1)  # the given rows are ONE augmentation batch
        try:
            table = self.next_table([c.shape[0] for c in clips])
        finally:
            self.cfg.batch_size = saved
        fixed = self.fix_length_device(clips, table.pad_before)
        return self.augment_device(fixed, table)

    def __call__(self, num_samples: int, **kwargs: Any) -> Iterator[Dict[str, Any]]:
        """Generates augmented audio samples as dataset rows (augmented.py:396-427)."""
        total_batches = int(np.ceil(num_samples / self.batch_size))
        for i in range(total_batches):
            batch_samples = min(self.batch_size, num_samples - i * self.batch_size)
            items = [self.get_next_audio_sample_dict() for _ in range(batch_samples)]
            results = self.execute_augment_batch([it["audio"] for it in items]).cpu().numpy()  # one D2H per batch
            for audio, item in zip(results, items):
                yield {"audio": {"array": audio, "sampling_rate": self.sample_rate},
                       **{k: v for k, v in item.items() if k != "audio"}}
