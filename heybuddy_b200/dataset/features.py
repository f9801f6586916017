"""
``TrainingFeaturesGenerator`` -- source clips -> augmentation -> log-mel -> speech embeddings -> ``[N, 16, 96]``
(reference ``heybuddy/dataset/features.py:30-908``), on the fused B200 pipeline.

What is kept: the constructor's keyword surface (TTS keywords are accepted and ignored with the source stage),
``autoconfigure`` / ``generate`` / ``__call__`` / ``default`` / ``get_wake_phrase_file_name`` /
``get_training_features`` / ``get_validation_features``, super-batches of ``sample_batch_size``, the cache
naming ``<safe>[_tst][_adv|_val].npy`` and the reuse/extend rule (features.py:686-760).

What differs: the Piper TTS stage is the path's *input boundary* (SURVEY.md 2 #15): clips come from ``source`` --
any callable ``n -> iterable of int16 clips`` (or a sequence) -- and default to a seeded synthetic source
(Piper voices cannot be downloaded offline).  No per-super-batch subprocess: the reference forks to let
PyTorch's host memory die with the child (features.py:516-532); here the pipeline reuses fixed device and
pinned buffers.  With ``rank`` / ``world_size`` every rank featurizes a contiguous range of augmentation
batches and writes its own row range of one shared ``.npy`` memmap (no collective on the data path).
"""
from __future__ import annotations

import math
import os
from typing import Any, Callable, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from heybuddy_b200 import spec
from heybuddy_b200.constants import *  # noqa: F401,F403
from heybuddy_b200.dataset.draws import AugmentConfig, DrawTable, draw_batch
from heybuddy_b200.dataset.precalculated import LOCAL_DIR, PrecalculatedDatasetIterator, open_shared_memmap
from heybuddy_b200.util import logger, safe_name

__all__ = ["TrainingFeaturesGenerator", "SyntheticSpeechSource", "shard_batches"]

SupplementalDatasetType = Optional[Any]


class SyntheticSpeechSource:
    """
    Stand-in for the TTS stage: seeded int16 clips of ragged length U[6400, 22400] -- band-limited noise under a
    raised-cosine envelope scaled to peak 32767, the type and range Piper emits (piper/pretrained.py:406-408).
    Clip i depends only on (seed, i), so any rank can generate any row range.
    """

    def __init__(self, seed: int = 2001, min_len: int = 6400, max_len: int = 22400) -> None:
        self.seed, self.min_len, self.max_len = seed, min_len, max_len

    def clip(self, i: int) -> np.ndarray:
        rng = np.random.Generator(np.random.PCG64([self.seed, int(i)]))
        n = int(rng.integers(self.min_len, self.max_len))
        x = np.convolve(rng.standard_normal(n + 7), np.ones(8) / 8.0, mode="valid")
        x *= 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)
        return (x / np.abs(x).max() * 32767.0).astype(np.int16)

    def __call__(self, n: int, start: int = 0) -> List[np.ndarray]:
        return [self.clip(start + i) for i in range(n)]


def shard_batches(n_batches: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of augmentation batches owned by ``rank``: [lo, hi)."""
    per, extra = divmod(n_batches, world_size)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


class TrainingFeaturesGenerator:
    """Generate a dataset of features."""

    def __init__(
        self,
        device_id: Optional[int] = None,
        use_tqdm: bool = False,
        use_autoconfigure: bool = True,
        sample_rate: int = 16000,
        sample_batch_size: int = DEFAULT_FEATURE_BATCH_SIZE,
        tts_text: str = "Hello, world!",
        tts_adversarial: bool = False,
        augment_target_length: float = 1.44,
        augment_batch_size: int = DEFAULT_AUGMENT_BATCH_SIZE,
        augment_sample_ratio: float = DEFAULT_AUGMENT_SAMPLE_RATIO,
        augment_background_dataset: SupplementalDatasetType = None,
        augment_impulse_dataset: SupplementalDatasetType = None,
        augment_seven_band_prob: float = 0.0,
        augment_tanh_distortion_prob: float = 0.0,
        augment_pitch_shift_prob: float = 0.0,
        augment_band_stop_prob: float = 0.0,
        augment_colored_noise_prob: float = DEFAULT_AUGMENT_COLORED_NOISE_PROB,
        augment_colored_noise_min_snr_db: float = DEFAULT_AUGMENT_COLORED_NOISE_MIN_SNR_DB,
        augment_colored_noise_max_snr_db: float = DEFAULT_AUGMENT_COLORED_NOISE_MAX_SNR_DB,
        augment_colored_noise_min_f_decay: float = DEFAULT_AUGMENT_COLORED_NOISE_MIN_F_DECAY,
        augment_colored_noise_max_f_decay: float = DEFAULT_AUGMENT_COLORED_NOISE_MAX_F_DECAY,
        augment_background_noise_prob: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_PROB,
        augment_background_noise_min_snr_db: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_MIN_SNR_DB,
        augment_background_noise_max_snr_db: float = DEFAULT_AUGMENT_BACKGROUND_NOISE_MAX_SNR_DB,
        augment_gain_prob: float = DEFAULT_AUGMENT_GAIN_PROB,
        augment_reverb_prob: float = DEFAULT_AUGMENT_REVERB_PROB,
        embedding_spectrogram_batch_size: int = DEFAULT_EMBEDDING_SPECTROGRAM_BATCH_SIZE,
        embedding_batch_size: int = DEFAULT_EMBEDDING_BATCH_SIZE,
        # B200 additions
        source: Optional[Union[Callable[..., Iterable[np.ndarray]], Sequence[np.ndarray]]] = None,
        seed: int = 2004,
        precision: Optional[str] = None,
        chunk_clips: int = 8192,
        rank: int = 0,
        world_size: int = 1,
        # stale spellings used by the reference's own tests (tests/test_feature_generator.py:17-24)
        device: Optional[Any] = None,
        tts_num_threads: Optional[int] = None,
        augment_num_threads: Optional[int] = None,
        **tts_kwargs: Any,
    ) -> None:
        if device is not None and device_id is None:
            import torch

            d = torch.device(device)
            device_id = d.index if d.index is not None else 0
        self.device_id = device_id
        self.use_tqdm = use_tqdm
        self.use_autoconfigure = use_autoconfigure
        self.sample_rate = sample_rate
        self.sample_batch_size = sample_batch_size
        self.tts_text = tts_text
        self.tts_adversarial = tts_adversarial
        self.tts_kwargs = tts_kwargs
        self.augment_target_length = augment_target_length
        self.augment_batch_size = augment_batch_size
        self.augment_sample_ratio = augment_sample_ratio
        self.augment_background_dataset = augment_background_dataset
        self.augment_impulse_dataset = augment_impulse_dataset
        self.augment_probs = dict(
            seven_band_aug_prob=augment_seven_band_prob, tanh_distortion_prob=augment_tanh_distortion_prob,
            pitch_shift_prob=augment_pitch_shift_prob, band_stop_prob=augment_band_stop_prob,
            colored_noise_prob=augment_colored_noise_prob, colored_noise_min_snr_db=augment_colored_noise_min_snr_db,
            colored_noise_max_snr_db=augment_colored_noise_max_snr_db, colored_noise_min_f_decay=augment_colored_noise_min_f_decay,
            colored_noise_max_f_decay=augment_colored_noise_max_f_decay, background_noise_prob=augment_background_noise_prob,
            background_noise_min_snr_db=augment_background_noise_min_snr_db,
            background_noise_max_snr_db=augment_background_noise_max_snr_db, gain_prob=augment_gain_prob, reverb_prob=augment_reverb_prob)
        self.embedding_spectrogram_batch_size = embedding_spectrogram_batch_size
        self.embedding_batch_size = embedding_batch_size
        self.source = source if source is not None else SyntheticSpeechSource(seed=2001 if not tts_adversarial else 2011)
        self.seed = seed
        self.precision = precision
        self.chunk_clips = chunk_clips
        self.rank, self.world_size = rank, world_size
        self._pipe = None
        self._generated = 0

    @property
    def device(self):
        from heybuddy_b200 import _native

        return _native.require_cuda(self.device_id)

    def autoconfigure(self) -> None:
        """Batch sizes from device memory (features.py:171-218); any B200 lands in the >= 8 GiB bucket."""
        import torch

        self.device_id = torch.cuda.current_device() if self.device_id is None else self.device_id
        total_gib = torch.cuda.get_device_properties(self.device).total_memory / (2 << 29)
        if total_gib >= 8:
            self.augment_batch_size, self.embedding_spectrogram_batch_size, self.embedding_batch_size = 128, 128, 128
        elif total_gib >= 4:
            self.augment_batch_size, self.embedding_spectrogram_batch_size, self.embedding_batch_size = 64, 64, 64
        else:
            self.augment_batch_size, self.embedding_spectrogram_batch_size, self.embedding_batch_size = 16, 32, 32

    # -- pipeline -------------------------------------------------------------------------------------------
    def _pipeline(self, augment: bool):
        from heybuddy_b200.dataset.augmented import AugmentedAudioGenerator
        from heybuddy_b200.embeddings import SpeechEmbeddings
        from heybuddy_b200.pipeline import FeaturizePipeline

        key = ("aug" if augment else "plain", self.augment_batch_size)
        if self._pipe is None or self._pipe[0] != key:
            probs = dict(self.augment_probs)
            if not augment:
                for k in ("colored_noise_prob", "background_noise_prob", "gain_prob", "reverb_prob"):
                    probs[k] = 0.0
            gen = AugmentedAudioGenerator(
                [], device_id=self.device_id, augmentation_dataset=self.augment_background_dataset if augment else None,
                impulse_response_dataset=self.augment_impulse_dataset if augment else None,
                target_length=self.augment_target_length, sample_rate=self.sample_rate, batch_size=self.augment_batch_size,
                seed=self.seed, **probs)
            speech = SpeechEmbeddings(device_id=self.device_id, precision=self.precision)
            self._pipe = (key, FeaturizePipeline(gen, speech, device_id=self.device_id), gen)
        return self._pipe[1], self._pipe[2]

    def _source_clips(self, n: int, start: int) -> List[np.ndarray]:
        if callable(self.source):
            try:
                clips = list(self.source(n, start=start))
            except TypeError:
                clips = list(self.source(n))
        else:
            clips = [self.source[(start + i) % len(self.source)] for i in range(n)]
        out = []
        for c in clips:
            c = c["audio"]["array"] if isinstance(c, dict) and "audio" in c else (c["array"] if isinstance(c, dict) else c)
            c = np.asarray(c)
            if c.dtype != np.int16:  # float clips in [-1, 1] -> the int16 the TTS stage would have produced
                c = np.clip(np.round(c * 32767.0), -32768, 32767).astype(np.int16)
            out.append(c)
        return out

    def generate(self, num_samples: int, sample_save_path: Optional[str] = None, augmented_sample_save_path: Optional[str] = None,
                 testing: bool = False, validation: bool = False, first_sample: Optional[int] = None) -> np.ndarray:
        """Generates ``num_samples`` clips and computes their embeddings -> ``f32 [num_samples, 16, 96]`` (features.py:360-490)."""
        from heybuddy_b200.pipeline import RaggedClips

        if self.use_autoconfigure:
            self.autoconfigure()
        start = self._generated if first_sample is None else first_sample
        pipe, gen = self._pipeline(augment=not validation)
        t = gen.target_num_samples
        b = self.augment_batch_size
        assert start % b == 0 or validation, "super-batches must start on an augmentation-batch boundary"
        clips = self._source_clips(num_samples, start)
        ragged = RaggedClips.from_list(clips)
        chunk = max(b, (self.chunk_clips // b) * b)
        tables = []
        gen._batch_index = start // b
        if not validation:
            # cursors at this super-batch's first augmentation batch, independent of how batches were sharded
            gen._noise_cursor, gen._rir_cursor = self._cursors_at(gen, start // b)
        for lo in range(0, num_samples, chunk):
            lengths = ragged.lengths[lo:lo + chunk]
            table = gen.next_table(lengths)
            if validation:
                # validation path: centre pad only (features.py:413-427), truncating long clips to T explicitly
                for d, l0 in zip(table.batches, range(0, len(lengths), b)):
                    ln = lengths[l0:l0 + b]
                    d.pad_before = np.maximum((t - ln) // 2, 0).astype(np.int32)
            tables.append(table)
        out, _, _ = pipe.featurize_host(ragged, tables, chunk)
        self._generated = start + num_samples
        return out

    def _cursors_at(self, gen, batch_index: int) -> Tuple[int, int]:
        """Noise-stream / RIR cursors before augmentation batch ``batch_index`` (prefix over the light draws)."""
        nb, rb = gen.noise_bank, gen.rir_bank
        noise_cursor = rir_cursor = 0
        if batch_index == 0 or (nb is None and rb is None):
            return 0, 0
        cfg = gen.cfg
        need = cfg.batch_size * cfg.target_samples
        for g in range(batch_index):
            d = draw_batch(self.seed, g, [cfg.target_samples] * cfg.batch_size, cfg, nb is not None, rb is not None, light=True)
            if d.background_apply:
                got = 0
                while got < need:
                    got += int(nb.clip_lengths[noise_cursor % len(nb)])
                    noise_cursor += 1
                noise_cursor %= len(nb)
            if d.reverb_apply:
                rir_cursor += 1
        return noise_cursor, rir_cursor

    def __call__(self, num_samples: int, sample_save_path: Optional[str] = None, augmented_sample_save_path: Optional[str] = None,
                 testing: bool = False, validation: bool = False) -> np.ndarray:
        """Super-batches of at most ``sample_batch_size`` samples (features.py:492-535), in process."""
        size = max(self.augment_batch_size, (self.sample_batch_size // max(self.augment_batch_size, 1)) * self.augment_batch_size)
        if self.use_autoconfigure:
            self.autoconfigure()
            size = max(self.augment_batch_size, (self.sample_batch_size // self.augment_batch_size) * self.augment_batch_size)
        parts, done = [], 0
        while done < num_samples:
            n = min(size, num_samples - done)
            parts.append(self.generate(n, sample_save_path, augmented_sample_save_path, testing, validation, first_sample=done))
            done += n
        return parts[0] if len(parts) == 1 else np.concatenate(parts)

    def generate_sharded(self, num_samples: int, path: str, barrier: Optional[Callable[[], None]] = None,
                         validation: bool = False) -> Tuple[int, int]:
        """
        Multi-GPU form: this rank featurizes its contiguous block of augmentation batches and writes rows
        [lo*B, hi*B) of the shared ``.npy`` at ``path`` (rank 0 creates it).  Returns the row range written.
        """
        if self.use_autoconfigure:
            self.autoconfigure()
        b = self.augment_batch_size
        n_batches = math.ceil(num_samples / b)
        lo_b, hi_b = shard_batches(n_batches, self.rank, self.world_size)
        lo, hi = lo_b * b, min(hi_b * b, num_samples)
        mm = open_shared_memmap(path, (num_samples, len(spec.embedding_frame_offsets(spec.CLIP_SAMPLES)), spec.EMB_DIM),
                                self.rank, barrier)
        size = max(b, (self.sample_batch_size // b) * b)
        row = lo
        while row < hi:
            n = min(size, hi - row)
            mm[row:row + n] = self.generate(n, validation=validation, first_sample=row)
            row += n
        mm.flush()
        if barrier is not None:
            barrier()
        return lo, hi

    # -- reference classmethods ----------------------------------------------------------------------------------
    @classmethod
    def default(cls, wake_phrase: str, adversarial: bool = False, **kwargs: Any) -> "TrainingFeaturesGenerator":
        kwargs.setdefault("use_autoconfigure", True)
        return cls(tts_text=wake_phrase, tts_adversarial=adversarial, **kwargs)

    @classmethod
    def get_wake_phrase_file_name(cls, wake_phrase: str, testing: bool = False) -> str:
        return safe_name(wake_phrase).strip("_") + ("_tst" if testing else "")

    @classmethod
    def _cached(cls, name: str, directory: str, use_cache: bool) -> Tuple[Optional[PrecalculatedDatasetIterator], int]:
        if not use_cache:
            return None, 0
        try:
            ds = PrecalculatedDatasetIterator(name, directory=directory)
            return ds, len(ds)
        except FileNotFoundError:
            return None, 0

    @classmethod
    def _features(cls, name: str, want: int, directory: str, use_cache: bool, keep_in_memory: bool, make: Callable[[], "TrainingFeaturesGenerator"],
                  **call_kwargs: Any) -> PrecalculatedDatasetIterator:
        """
        Reuse ``<name>.npy`` when it has enough rows, otherwise generate the missing rows (features.py:686-760).  The
        reference concatenates old + new in memory and rewrites the whole file; here the new rows are appended in place
        (``util/npy_append.py``: the header gets spare digits once, then only the row count changes) -- same file contents.
        """
        existing, have = cls._cached(name, directory, use_cache)
        if existing is not None and have >= want:
            return existing
        gen = make()
        if have > 0:
            from heybuddy_b200.util.npy_append import AppendableNumpyArrayFile, AppendableNumpyHeaderInfo

            gen._generated = have
            new_rows = gen(want - have, **call_kwargs)
            path = os.path.join(directory, f"{name}.npy")
            del existing                       # drop the read-only memmap before the file changes under it
            AppendableNumpyHeaderInfo.ensure_appendable(path, in_place=True)
            with AppendableNumpyArrayFile(path) as out:
                out.append(new_rows)
            return PrecalculatedDatasetIterator(name, directory=directory)
        feats = gen(want, **call_kwargs)
        return PrecalculatedDatasetIterator.from_array(feats, name=name, directory=directory, keep_in_memory=keep_in_memory)

    @classmethod
    def get_training_features(cls, wake_phrase: str, num_positive_samples: int, num_adversarial_samples: int, testing: bool = False,
                              use_cache: bool = True, save_samples: bool = False, keep_in_memory: bool = False,
                              directory: Optional[str] = None, **kwargs: Any) -> Tuple[PrecalculatedDatasetIterator, PrecalculatedDatasetIterator]:
        """(positive, adversarial) iterators over ``<name>.npy`` / ``<name>_adv.npy`` (features.py:628-838)."""
        directory = directory or LOCAL_DIR
        name = cls.get_wake_phrase_file_name(wake_phrase, testing=testing)
        positive = cls._features(name, num_positive_samples, directory, use_cache, keep_in_memory,
                                 lambda: cls.default(wake_phrase, **kwargs), testing=testing)
        adversarial = cls._features(f"{name}_adv", num_adversarial_samples, directory, use_cache, keep_in_memory,
                                    lambda: cls.default(wake_phrase, adversarial=True, **kwargs), testing=testing)
        return positive, adversarial

    @classmethod
    def get_validation_features(cls, wake_phrase: str, num_positive_samples: int, use_cache: bool = True, save_samples: bool = False,
                                keep_in_memory: bool = False, directory: Optional[str] = None, **kwargs: Any) -> PrecalculatedDatasetIterator:
        """Un-augmented, centre-padded positives in ``<name>_val.npy`` (features.py:840-908)."""
        directory = directory or LOCAL_DIR
        name = cls.get_wake_phrase_file_name(wake_phrase) + "_val"
        return cls._features(name, num_positive_samples, directory, use_cache, keep_in_memory,
                             lambda: cls.default(wake_phrase, **kwargs), validation=True)
