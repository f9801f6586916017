"""
In-the-wild extractor: arbitrary recordings -> 1.44 s pieces -> mel -> embeddings ->
``<output_dir>/<name>/<k>.npy`` chunk files of ``[rows, 16, 96]`` (or ``[rows, 17, 96]`` with a row of token ids).

Replaces ``PrecalculatedTrainingDatasetGenerator`` / ``PrecalculatedLabeledTrainingDatasetGenerator``
(reference ``dataset/precalculated.py:40-363``) -- this is how the reference's large negative sets are made.
Same constructor arguments, same piece / batch / file bookkeeping (file names, row counts and row order are
identical: ``tests/golden/extractor.npz`` comes from the reference's own class), different engine:

* the reference featurizes one 128-piece batch at a time through ``SpeechEmbeddings.__call__`` (4 overlapping mel windows
  and 16 embedding windows per piece, ORT batches of 32);
* here pieces are collected into groups of ``gpu_pieces`` (default 4096 = 32 batches), uploaded once, and featurized by the
  fused device path (``SpeechEmbeddings.embed_device``: one mel per piece, fully-convolutional embedding); the per-batch
  bookkeeping (label hook, NaN rows dropped, flush when the buffer holds ``samples_per_file`` rows) is then replayed on the
  host in the reference's order, so the files come out the same.

Input boundary: ``dataset_path`` may be an iterable of samples (dicts shaped like HF ``datasets`` rows) or a path handed to
``datasets.load_dataset`` when that package and the data are available (the hub is a network service: out of scope here).
Resampling of non-16 kHz recordings uses ``torchaudio.transforms.Resample`` on the host, like the reference (input decode,
not the hot path).
"""
from __future__ import annotations

import os
from math import ceil, log10
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple, Union

import numpy as np

from heybuddy_b200.dataset.precalculated import LOCAL_DIR
from heybuddy_b200.util import logger

__all__ = ["PrecalculatedTrainingDatasetGenerator", "PrecalculatedLabeledTrainingDatasetGenerator"]

Sample = Dict[str, Any]
Batch = List[Tuple[np.ndarray, Sample]]


class PrecalculatedTrainingDatasetGenerator:
    """Takes existing in-the-wild audio and creates packed datasets of features."""

    def __init__(self, dataset_path: Union[str, Iterable[Sample]], config_name: Optional[str] = None, split: str = "train",
                 audio_key: str = "audio", audio_array_key: Optional[str] = "array",
                 audio_sample_rate_key: Optional[str] = "sampling_rate", device_id: Optional[int] = None, sample_rate: int = 16000,
                 seconds_per_batch: float = 1.44, process_batch_size: int = 128, embedding_batch_size: int = 32,
                 gpu_pieces: int = 4096, precision: Optional[str] = None) -> None:
        self.dataset_path = dataset_path
        self.config_name = config_name
        self.split = split
        self.audio_key = audio_key
        self.audio_array_key = audio_array_key
        self.audio_sample_rate_key = audio_sample_rate_key
        self.device_id = device_id
        self.sample_rate = sample_rate
        self.seconds_per_batch = seconds_per_batch
        self.process_batch_size = process_batch_size
        self.embedding_batch_size = embedding_batch_size   # kept for signature parity; the device path has no use for it
        self.gpu_pieces = max(int(gpu_pieces), process_batch_size)
        self.precision = precision

    @property
    def samples_per_batch(self) -> int:
        return int(self.sample_rate * self.seconds_per_batch)

    @property
    def speech_embeddings(self):
        if not hasattr(self, "_speech_embeddings"):
            from heybuddy_b200.embeddings import SpeechEmbeddings
            self._speech_embeddings = SpeechEmbeddings(device_id=self.device_id, precision=self.precision)
        return self._speech_embeddings

    # -- the two hooks --------------------------------------------------------------------------------------
    def embed_pieces(self, pieces: np.ndarray) -> np.ndarray:
        """f32 ``[n, samples_per_batch]`` in [-1, 1] -> f32 ``[n, 16, 96]`` (NaNs kept: the caller drops those rows)."""
        import torch

        speech = self.speech_embeddings
        host = torch.from_numpy(np.ascontiguousarray(pieces, dtype=np.float32))
        if not host.is_pinned():
            host = host.pin_memory()
        emb = speech.embed_device(host.to(speech.device, non_blocking=True))
        return emb.cpu().numpy()

    def label_embeddings(self, embeddings: np.ndarray, batch: Batch) -> np.ndarray:
        """Adds any additional rows to the batch's embeddings; the base class returns them as they are."""
        return embeddings

    # -- input side --------------------------------------------------------------------------------------------
    def _samples(self, dataset_streaming: bool, trust_remote_code: bool) -> Iterable[Sample]:
        if not isinstance(self.dataset_path, str):
            return self.dataset_path
        try:
            from datasets import load_dataset
        except ImportError as ex:  # pragma: no cover
            raise RuntimeError("dataset_path is a name/path but the `datasets` package is not installed; pass an iterable of samples") from ex
        return load_dataset(self.dataset_path, self.config_name, split=self.split, streaming=dataset_streaming,
                            trust_remote_code=trust_remote_code)

    def _pieces(self, samples: Iterable[Sample]):
        """(piece f32[samples_per_batch], sample) in the reference's order: consecutive pieces, the last one right-padded."""
        resamplers: Dict[int, Any] = {}
        spb = self.samples_per_batch
        for sample in samples:
            sample = dict(sample)
            audio = sample.pop(self.audio_key)
            rate = None
            if self.audio_sample_rate_key is not None:
                try:
                    rate = audio[self.audio_sample_rate_key]
                except (KeyError, TypeError, IndexError):
                    rate = sample.get(self.audio_sample_rate_key)
            if self.audio_array_key is not None:
                audio = audio[self.audio_array_key]
            audio = np.asarray(audio)
            if rate is not None and rate != self.sample_rate:
                import torch
                import torchaudio

                if rate not in resamplers:
                    resamplers[rate] = torchaudio.transforms.Resample(rate, self.sample_rate).to(dtype=torch.float32)
                audio = resamplers[rate](torch.tensor(audio).to(dtype=torch.float32)).numpy()
            audio = audio.astype(np.float32)
            for i in range(0, len(audio), spb):
                piece = audio[i:i + spb]
                if piece.shape[0] < spb:
                    piece = np.concatenate([piece, np.zeros(spb - piece.shape[0], dtype=np.float32)])
                yield piece, sample

    # -- driver ------------------------------------------------------------------------------------------------
    def __call__(self, name: str, output_dir: str = LOCAL_DIR, max_hours: float = 1000.0, dataset_streaming: bool = True,
                 trust_remote_code: bool = False, samples_per_file: int = 10000,
                 on_progress: Optional[Callable[[int, int], None]] = None) -> List[str]:
        """Writes the chunk files; returns their paths (the reference returns None)."""
        output_dir = os.path.join(output_dir, name)
        os.makedirs(output_dir, exist_ok=True)
        pbs = self.process_batch_size
        max_batches = int(max_hours * 3600 / self.seconds_per_batch / pbs)
        num_files = ceil((max_batches * pbs) / samples_per_file)
        digits = int(log10(num_files)) + 1 if num_files > 0 else 1
        logger.info(f"Will generate up to {max_batches * pbs} samples from {max_hours} hours of data. Writing {num_files} files to \"{output_dir}\".")

        data_files: List[str] = []
        buffer: Optional[np.ndarray] = None
        formed: List[Batch] = []           # complete batches waiting for the next device pass
        num_batches = 0                    # batches formed so far (the reference's counter, processing there is immediate)
        done_batches = 0

        def flush_buffer() -> None:
            nonlocal buffer
            path = os.path.join(output_dir, f"{len(data_files):0{digits}d}.npy")
            np.save(path, buffer)
            data_files.append(path)
            buffer = None

        def run_formed() -> None:
            """One device pass over every waiting batch, then the reference's per-batch bookkeeping in order."""
            nonlocal buffer, done_batches
            if not formed:
                return
            pieces = np.stack([a for b in formed for (a, _) in b])
            emb = self.embed_pieces(pieces)
            lo = 0
            for b in formed:
                e = self.label_embeddings(embeddings=emb[lo:lo + len(b)], batch=b)
                lo += len(b)
                keep = ~np.isnan(e).any(axis=(1, 2))
                if not keep.all():
                    logger.warning(f"Removed {int((~keep).sum())} samples due to NaN values in embeddings.")
                e = e[keep]
                buffer = e if buffer is None else np.concatenate([buffer, e])
                done_batches += 1
                if on_progress is not None:
                    on_progress(done_batches, max_batches)
                if buffer is not None and buffer.shape[0] >= samples_per_file:
                    flush_buffer()
            formed.clear()

        batch: Batch = []
        for piece, sample in self._pieces(self._samples(dataset_streaming, trust_remote_code)):
            batch.append((piece, sample))
            if len(batch) >= pbs:
                formed.append(batch)
                batch = []
                num_batches += 1
                if sum(len(b) for b in formed) >= self.gpu_pieces:
                    run_formed()
            if num_batches >= max_batches:
                break
        if len(batch) > 0 and num_batches < max_batches:
            formed.append(batch)
            num_batches += 1
        run_formed()
        if buffer is not None:
            flush_buffer()
        return data_files


class PrecalculatedLabeledTrainingDatasetGenerator(PrecalculatedTrainingDatasetGenerator):
    """
    Features plus labels: row 16 of every ``[17, 96]`` sample holds the token ids of the recording's transcript (as f32).
    ``tokenizer``: callable ``text -> int array [tokenizer_max_length]``.  The reference's default is a BERT WordPiece tokenizer
    whose vocabulary is a download; pass one in (e.g. ``tokenizers.Tokenizer.from_file``-based) -- there is no silent default.
    """

    def __init__(self, dataset_path: Union[str, Iterable[Sample]], config_name: Optional[str] = None, split: str = "train",
                 audio_key: str = "audio", audio_array_key: Optional[str] = "array",
                 audio_sample_rate_key: Optional[str] = "sampling_rate", transcript_key: str = "transcript",
                 device_id: Optional[int] = None, sample_rate: int = 16000, seconds_per_batch: float = 1.44,
                 process_batch_size: int = 128, embedding_batch_size: int = 32, tokenizer_max_length: int = 96,
                 tokenizer: Optional[Callable[[str], Any]] = None, **kwargs: Any) -> None:
        super().__init__(dataset_path=dataset_path, config_name=config_name, split=split, audio_key=audio_key,
                         audio_array_key=audio_array_key, audio_sample_rate_key=audio_sample_rate_key, device_id=device_id,
                         sample_rate=sample_rate, seconds_per_batch=seconds_per_batch, process_batch_size=process_batch_size,
                         embedding_batch_size=embedding_batch_size, **kwargs)
        self.transcript_key = transcript_key
        self.tokenizer_max_length = tokenizer_max_length
        if tokenizer is not None:
            self._tokenizer = tokenizer

    @property
    def tokenizer(self) -> Callable[[str], Any]:
        if not hasattr(self, "_tokenizer"):
            raise RuntimeError("PrecalculatedLabeledTrainingDatasetGenerator needs a tokenizer: pass tokenizer=callable(text) -> "
                               f"int[{self.tokenizer_max_length}] (the reference's BERT vocabulary is a network download)")
        return self._tokenizer

    def tokenize(self, text: str) -> np.ndarray:
        if getattr(self, "_last_text", None) == text:
            return self._last_tokens
        tokens = self.tokenizer(text)
        tokens = tokens.numpy() if hasattr(tokens, "numpy") else np.asarray(tokens)
        self._last_text, self._last_tokens = text, tokens
        return tokens

    def label_embeddings(self, embeddings: np.ndarray, batch: Batch) -> np.ndarray:
        seen: Dict[str, np.ndarray] = {}
        rows = []
        for _, sample in batch:
            text = sample[self.transcript_key]
            if text not in seen:
                seen[text] = self.tokenize(text)[np.newaxis, ...]
            rows.append(seen[text])
        return np.stack([np.concatenate([e, t], axis=0) for e, t in zip(embeddings, rows)]).astype(np.float32)
