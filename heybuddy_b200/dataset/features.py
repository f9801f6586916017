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
from heybuddy_b200.dataset.precalculated import LOCAL_DIR, PrecalculatedDatasetIterator, open_shared_memmap
from heybuddy_b200.util import logger, safe_name

__all__ = ["TrainingFeaturesGenerator", "SyntheticSpeechSource", "RaggedClipSource", "shard_batches"]

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


class RaggedClipSource:
    """
    A source that already holds its clips the way the TTS stage hands them over -- one concatenated int16 buffer (pinned host
    memory when ``pin=True``) + offsets -- and serves row ranges as views, wrapping around like the reference's source iterator
    (augmented.py:176-186 restarts the dataset when it is exhausted).  ``ragged(n, start)`` is the fast path
    ``TrainingFeaturesGenerator`` looks for: no per-clip Python between the TTS buffer and the H2D copy.
    """

    def __init__(self, clips, pin: bool = False) -> None:
        from heybuddy_b200.pipeline import RaggedClips

        self.clips = clips if isinstance(clips, RaggedClips) else RaggedClips.from_list(list(clips))
        if pin and self.clips.pinned is None:
            self.clips = self.clips.pin()

    def __len__(self) -> int:
        return len(self.clips)

    def ragged(self, n: int, start: int = 0):
        """Yields ``(rows, RaggedClips view)`` pieces covering clips ``start .. start + n`` (mod len), each a contiguous view."""
        total = len(self.clips)
        at, left = start % total, n
        while left > 0:
            take = min(left, total - at)
            yield take, self.clips.slice(at, at + take)
            at, left = (at + take) % total, left - take

    def __call__(self, n: int, start: int = 0) -> List[np.ndarray]:
        out = []
        for _, part in self.ragged(n, start):
            out += [part.samples[part.offsets[i]:part.offsets[i + 1]] for i in range(len(part))]
        return out


def shard_batches(n_batches: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of augmentation batches owned by ``rank``: [lo, hi)."""
    per, extra = divmod(n_batches, world_size)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


# Train / test / validation splits must not share utterances or draws (the reference gets this for free: every TTS call and every
# np.random draw is fresh).  Each split reads the source at its own offset and seeds its draw table differently.
SPLIT_SOURCE_OFFSET = {"train": 0, "test": 1 << 28, "validation": 1 << 29}
SPLIT_SEED_SALT = {"train": 0, "test": 0x7E57, "validation": 0x7A11D}


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
        augment_seven_band_gain_db: float = DEFAULT_AUGMENT_SEVEN_BAND_GAIN_DB,
        augment_tanh_distortion_prob: float = 0.0,
        augment_tanh_min_distortion: float = DEFAULT_AUGMENT_TANH_MIN_DISTORTION,
        augment_tanh_max_distortion: float = DEFAULT_AUGMENT_TANH_MAX_DISTORTION,
        augment_pitch_shift_prob: float = 0.0,
        augment_pitch_shift_semitones: int = DEFAULT_AUGMENT_PITCH_SHIFT_SEMITONES,
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
        chunk_clips: int = 4096,
        rank: int = 0,
        world_size: int = 1,
        embedding_weights: Optional[Any] = None,
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
            seven_band_aug_prob=augment_seven_band_prob, seven_band_aug_gain_db=augment_seven_band_gain_db,
            tanh_distortion_prob=augment_tanh_distortion_prob, tanh_min_distortion=augment_tanh_min_distortion,
            tanh_max_distortion=augment_tanh_max_distortion, pitch_shift_prob=augment_pitch_shift_prob,
            pitch_shift_semitones=augment_pitch_shift_semitones, band_stop_prob=augment_band_stop_prob,
            colored_noise_prob=augment_colored_noise_prob, colored_noise_min_snr_db=augment_colored_noise_min_snr_db,
            colored_noise_max_snr_db=augment_colored_noise_max_snr_db, colored_noise_min_f_decay=augment_colored_noise_min_f_decay,
            colored_noise_max_f_decay=augment_colored_noise_max_f_decay, background_noise_prob=augment_background_noise_prob,
            background_noise_min_snr_db=augment_background_noise_min_snr_db,
            background_noise_max_snr_db=augment_background_noise_max_snr_db, gain_prob=augment_gain_prob, reverb_prob=augment_reverb_prob)
        self.embedding_spectrogram_batch_size = embedding_spectrogram_batch_size
        self.embedding_batch_size = embedding_batch_size
        if source is None:
            # The TTS stage (Piper voices, hub downloads) is the path's input boundary and is not available offline: without an
            # explicit source the generator featurizes SYNTHETIC band-limited noise, not speech of `tts_text`.  Say so, loudly.
            logger.warning(
                f"TrainingFeaturesGenerator(tts_text={tts_text!r}): no `source` of TTS clips was given -- featurizing SYNTHETIC "
                "noise bursts (SyntheticSpeechSource), NOT speech; the wake phrase only names the output file. Pass source=<callable "
                "or sequence of int16 16 kHz clips> for real training data.")
            source = SyntheticSpeechSource(seed=2001 if not tts_adversarial else 2011)
        self.source = source
        self.seed = seed
        self.precision = precision
        self.chunk_clips = chunk_clips
        self.rank, self.world_size = rank, world_size
        self.embedding_weights = embedding_weights
        self._pipe = None
        self._generated = 0
        self._cursor_cache = {}   # (split seed) -> (batch index, noise cursor, rir cursor) of the furthest prefix computed so far
        self.last_h2d_bytes = self.last_d2h_bytes = 0
        self.last_sink = ""

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
                for k in ("colored_noise_prob", "background_noise_prob", "gain_prob", "reverb_prob", "seven_band_aug_prob",
                          "tanh_distortion_prob", "pitch_shift_prob", "band_stop_prob"):
                    probs[k] = 0.0
            gen = AugmentedAudioGenerator(
                [], device_id=self.device_id, augmentation_dataset=self.augment_background_dataset if augment else None,
                impulse_response_dataset=self.augment_impulse_dataset if augment else None,
                target_length=self.augment_target_length, sample_rate=self.sample_rate, batch_size=self.augment_batch_size,
                seed=self.seed, **probs)
            speech = SpeechEmbeddings(device_id=self.device_id, precision=self.precision, weights=self.embedding_weights)
            self._pipe = (key, FeaturizePipeline(gen, speech, device_id=self.device_id), gen)
        return self._pipe[1], self._pipe[2]

    def _source_ragged(self, n: int, start: int):
        """Source clips ``start .. start + n`` as ``(count, RaggedClips)`` pieces (views of the source's buffer when it has one)."""
        from heybuddy_b200.pipeline import RaggedClips

        if hasattr(self.source, "ragged"):
            yield from self.source.ragged(n, start)
            return
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
        yield n, RaggedClips.from_list(out)

    def _source_clips(self, n: int, start: int) -> List[np.ndarray]:
        out = []
        for _, part in self._source_ragged(n, start):
            out += [part.samples[part.offsets[i]:part.offsets[i + 1]] for i in range(len(part))]
        return out

    @staticmethod
    def _split(testing: bool, validation: bool) -> str:
        return "validation" if validation else ("test" if testing else "train")

    def _cursors_at(self, gen, batch_index: int, seed: int) -> Tuple[int, int]:
        """
        Noise-stream / RIR cursors before augmentation batch ``batch_index``: a prefix over the batch coins (one vectorised
        Philox call for the whole range) resumed from the furthest prefix computed so far -- linear over a long run.
        """
        from heybuddy_b200.dataset.draws import advance_noise_cursor, batch_coins

        nb, rb = gen.noise_bank, gen.rir_bank
        if batch_index == 0 or (nb is None and rb is None):
            return 0, 0
        g0, noise_cursor, rir_cursor = self._cursor_cache.get(seed, (0, 0, 0))
        if g0 > batch_index:
            g0, noise_cursor, rir_cursor = 0, 0, 0
        cfg = gen.cfg
        bg, rev = batch_coins(seed, np.arange(g0, batch_index, dtype=np.uint64), cfg, nb is not None, rb is not None)
        rir_cursor += int(rev.sum())
        n_bg = int(bg.sum())
        if n_bg:
            need = cfg.batch_size * cfg.target_samples      # every batch before the last of a dataset is full
            lengths = nb.clip_lengths
            if (lengths == lengths[0]).all():
                noise_cursor = (noise_cursor + n_bg * (-(-need // int(lengths[0])))) % len(nb)
            else:
                for _ in range(n_bg):
                    noise_cursor = advance_noise_cursor(noise_cursor, need, nb.clip_starts)
        self._cursor_cache[seed] = (batch_index, noise_cursor, rir_cursor)
        return noise_cursor, rir_cursor

    def _items(self, first_sample: int, num_samples: int, sink_of: Callable[[int, int], Any], validation: bool, testing: bool):
        """
        Lazily yields ``(clips, table, sink)`` items for :meth:`FeaturizePipeline.featurize_stream`: one per contiguous piece of a
        super-batch of at most ``sample_batch_size`` source clips (features.py:492-535), each with the vectorised draw table of
        its augmentation batches.  ``sink_of(row_lo, row_hi)`` returns the sink for output rows [row_lo, row_hi) of this call.
        """
        from heybuddy_b200.dataset.draws import DrawTable

        split = self._split(testing, validation)
        pipe, gen = self._pipeline(augment=not validation)
        b, t = self.augment_batch_size, gen.target_num_samples
        seed = self.seed + SPLIT_SEED_SALT[split]
        nb, rb = gen.noise_bank, gen.rir_bank
        size = max(b, (self.sample_batch_size // b) * b)
        assert first_sample % b == 0, "a dataset range must start on an augmentation-batch boundary"
        noise_cursor, rir_cursor = self._cursors_at(gen, first_sample // b, seed) if not validation else (0, 0)
        done = 0
        while done < num_samples:
            want = min(size, num_samples - done)
            pending = None      # a piece that does not end on a batch boundary is merged with the next one
            for _, part in self._source_ragged(want, SPLIT_SOURCE_OFFSET[split] + first_sample + done):
                if pending is not None:
                    from heybuddy_b200.pipeline import RaggedClips

                    part = RaggedClips(np.concatenate([pending.samples, part.samples]),
                                       np.concatenate([pending.offsets, pending.offsets[-1] + part.offsets[1:]]))
                    pending = None
                n = len(part)
                whole = n if done + n >= num_samples else (n // b) * b
                if whole < n:
                    pending = part.slice(whole, n)
                    part = part.slice(0, whole)
                if whole == 0:
                    continue
                row = first_sample + done
                table = DrawTable.build(part.lengths, gen.cfg, seed, nb.clip_lengths if nb is not None else None,
                                        len(rb) if rb is not None else 0, first_batch=row // b, noise_cursor=noise_cursor,
                                        rir_cursor=rir_cursor)
                noise_cursor, rir_cursor = table.final_noise_cursor, table.final_rir_cursor
                if validation:
                    # validation path: centre pad only (features.py:413-427), truncating long clips to T explicitly
                    table.pad_before = np.maximum((t - part.lengths) // 2, 0).astype(np.int32)
                yield part, table, sink_of(done, done + whole)
                done += whole
            assert pending is None, "the source returned fewer clips than asked for"
        if not validation:
            self._cursor_cache[seed] = ((first_sample + num_samples + b - 1) // b, noise_cursor, rir_cursor)

    def _run(self, first_sample: int, num_samples: int, sink_of, validation: bool, testing: bool, writer_threads: int = 4,
             wait: bool = True) -> None:
        pipe, gen = self._pipeline(augment=not validation)
        b = self.augment_batch_size
        chunk = max(b, (self.chunk_clips // b) * b)
        self.last_h2d_bytes, self.last_d2h_bytes = pipe.featurize_stream(
            self._items(first_sample, num_samples, sink_of, validation, testing), chunk, writer_threads=writer_threads, wait=wait)

    def generate(self, num_samples: int, sample_save_path: Optional[str] = None, augmented_sample_save_path: Optional[str] = None,
                 testing: bool = False, validation: bool = False, first_sample: Optional[int] = None, out: Optional[Any] = None) -> np.ndarray:
        """
        Generates ``num_samples`` clips and computes their embeddings -> ``f32 [num_samples, 16, 96]`` (features.py:360-490).
        Rows are samples ``first_sample ..`` of this generator's stream (default: where the previous call stopped, rounded up to
        an augmentation-batch boundary), so successive calls -- and a cache extension -- never repeat a sample.
        ``out``: optional destination, a numpy array or a pinned CPU torch tensor ``[num_samples, 16, 96]`` f32 (the device -> host
        copies then land in it directly).
        """
        if self.use_autoconfigure:
            self.autoconfigure()
        b = self.augment_batch_size
        start = self._generated if first_sample is None else first_sample
        start = -(-start // b) * b
        if out is None:
            out = np.empty((num_samples, len(spec.embedding_frame_offsets(spec.CLIP_SAMPLES)), spec.EMB_DIM), dtype=np.float32)
        assert tuple(out.shape) == (num_samples, len(spec.embedding_frame_offsets(spec.CLIP_SAMPLES)), spec.EMB_DIM)
        self._run(start, num_samples, lambda lo, hi: out[lo:hi], validation, testing)
        self._generated = start + num_samples
        return out.numpy() if hasattr(out, "is_pinned") else out

    def __call__(self, num_samples: int, sample_save_path: Optional[str] = None, augmented_sample_save_path: Optional[str] = None,
                 testing: bool = False, validation: bool = False) -> np.ndarray:
        """
        ``num_samples`` NEW samples (features.py:492-535).  The reference forks one child per super-batch of ``sample_batch_size``
        so that PyTorch's host memory dies with it; here the super-batches stream through one pipeline with fixed buffers.
        """
        return self.generate(num_samples, sample_save_path, augmented_sample_save_path, testing, validation)

    def generate_sharded(self, num_samples: int, path: str, barrier: Optional[Callable[[], None]] = None,
                         validation: bool = False, testing: bool = False, writer_threads: int = 4, defer: bool = False):
        """
        Multi-GPU form: this rank featurizes its contiguous block of augmentation batches and writes rows
        [lo*B, hi*B) of the shared ``.npy`` at ``path`` (rank 0 creates it).  Rows go from the pipeline's pinned D2H slots
        straight into the file (``pwrite`` from worker threads), overlapping the kernels of the following chunks.
        Returns the row range written -- or, with ``defer=True``, a ``finish()`` callable that waits for the last rows to reach
        the file, closes it, runs the final barrier and returns the row range (lets the caller start the next file's kernels while
        this one's tail is still being written: ``get_training_features`` does so for the positive / adversarial pair).
        """
        from heybuddy_b200.dataset.precalculated import NpyRowWriter

        if self.use_autoconfigure:
            self.autoconfigure()
        b = self.augment_batch_size
        n_batches = math.ceil(num_samples / b)
        lo_b, hi_b = shard_batches(n_batches, self.rank, self.world_size)
        lo, hi = lo_b * b, min(hi_b * b, num_samples)
        shape = (num_samples, len(spec.embedding_frame_offsets(spec.CLIP_SAMPLES)), spec.EMB_DIM)
        writer = NpyRowWriter(path, shape, create=self.rank == 0, barrier=barrier, shared=self.world_size > 1)
        mapped = None
        try:
            if hi > lo:
                # opt-in on memory-backed file systems: the D2H copies land in the file's own (registered) pages; otherwise rows go
                # from the pipeline's pinned slots into the file with pwrite
                mapped = writer.map_pinned(lo, hi, populate_threads=writer_threads)
                self.last_sink = "pinned file mapping (D2H straight into the page cache)" if mapped is not None else f"{writer.mode} from pinned slots"
                if mapped is not None:
                    sink_of = lambda r0, r1: mapped[r0:r1]
                else:
                    sink_of = lambda r0, r1: (lambda a, z, rows, base=lo + r0: writer.write(base + a, rows))
                self._run(lo, hi - lo, sink_of, validation, testing, writer_threads=writer_threads, wait=False)
        except BaseException:
            writer.close()
            raise

        def finish() -> Tuple[int, int]:
            nonlocal mapped
            try:
                if self._pipe is not None:
                    self._pipe[1].finish_stream()
                mapped = None
            finally:
                writer.close()
            if barrier is not None:
                barrier()
            return lo, hi

        return finish if defer else finish()

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
                  defer: bool = False, **call_kwargs: Any):
        """
        Reuse ``<name>.npy`` when it has enough rows, otherwise generate the missing rows (features.py:686-760).  The
        reference concatenates old + new in memory and rewrites the whole file; here the new rows are appended in place
        (``util/npy_append.py``: the header gets spare digits once, then only the row count changes) -- same file contents.
        """
        existing, have = cls._cached(name, directory, use_cache)
        if existing is not None and have >= want:
            return (lambda: existing) if defer else existing
        gen = make()
        if have > 0:
            from heybuddy_b200.util.npy_append import AppendableNumpyArrayFile, AppendableNumpyHeaderInfo

            gen._generated = have      # rows [0, have) exist: the new ones continue the generator's stream (never a repeat)
            new_rows = gen(want - have, **call_kwargs)
            path = os.path.join(directory, f"{name}.npy")
            del existing                       # drop the read-only memmap before the file changes under it
            AppendableNumpyHeaderInfo.ensure_appendable(path, in_place=True)
            with AppendableNumpyArrayFile(path) as out:
                out.append(new_rows)
            it = PrecalculatedDatasetIterator(name, directory=directory)
            return (lambda: it) if defer else it
        # nothing cached: every rank streams its rows straight into `<name>.npy` (same bytes as the reference's np.save of the
        # whole array, precalculated.py:486, without ever holding the whole array)
        barrier = None
        if gen.world_size > 1:
            import torch.distributed as dist

            assert dist.is_initialized(), "world_size > 1 needs an initialised torch.distributed process group (file barrier)"
            barrier = dist.barrier
        finish = gen.generate_sharded(want, os.path.join(directory, f"{name}.npy"), barrier=barrier, defer=True, **call_kwargs)

        def open_it() -> PrecalculatedDatasetIterator:
            finish()
            return PrecalculatedDatasetIterator(name, directory=directory, use_mem_map=not keep_in_memory)

        return open_it if defer else open_it()

    @classmethod
    def get_training_features(cls, wake_phrase: str, num_positive_samples: int, num_adversarial_samples: int, testing: bool = False,
                              use_cache: bool = True, save_samples: bool = False, keep_in_memory: bool = False,
                              directory: Optional[str] = None, **kwargs: Any) -> Tuple[PrecalculatedDatasetIterator, PrecalculatedDatasetIterator]:
        """(positive, adversarial) iterators over ``<name>.npy`` / ``<name>_adv.npy`` (features.py:628-838)."""
        directory = directory or LOCAL_DIR
        name = cls.get_wake_phrase_file_name(wake_phrase, testing=testing)
        # the adversarial file's kernels start while the tail of the positive file is still on its way to the page cache
        positive = cls._features(name, num_positive_samples, directory, use_cache, keep_in_memory,
                                 lambda: cls.default(wake_phrase, **kwargs), defer=True, testing=testing)
        adversarial = cls._features(f"{name}_adv", num_adversarial_samples, directory, use_cache, keep_in_memory,
                                    lambda: cls.default(wake_phrase, adversarial=True, **kwargs), defer=True, testing=testing)
        return positive(), adversarial()

    @classmethod
    def get_validation_features(cls, wake_phrase: str, num_positive_samples: int, use_cache: bool = True, save_samples: bool = False,
                                keep_in_memory: bool = False, directory: Optional[str] = None, **kwargs: Any) -> PrecalculatedDatasetIterator:
        """Un-augmented, centre-padded positives in ``<name>_val.npy`` (features.py:840-908)."""
        directory = directory or LOCAL_DIR
        name = cls.get_wake_phrase_file_name(wake_phrase) + "_val"
        return cls._features(name, num_positive_samples, directory, use_cache, keep_in_memory,
                             lambda: cls.default(wake_phrase, **kwargs), validation=True)
