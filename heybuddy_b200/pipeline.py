"""
The fused featurization pipeline: ragged int16 clips -> length fix -> augmentation -> log-mel ->
speech embeddings, everything resident on one GPU between the first H2D copy and the final D2H of
the ``[n, 16, 96]`` embeddings.

This is what ``TrainingFeaturesGenerator.generate`` (reference dataset/features.py:360-490) does
with a python loop per clip, a D2H per clip and one ORT call per 32 items; here a chunk of
augmentation batches is a handful of kernel launches on one stream:

    hb_colored_bases (the chunk's coloured-noise patterns, from the draw table's counters)
    hb_featurize_i16 = hb_augment_clips_i16 (length fix + augmentation) -> hb_mel_f32 -> hb_embed_clips

``featurize_stream`` is the host side: per chunk ONE pinned metadata blob (offsets, pad offsets, per-clip draw
records, coloured batch ids) and the int16 samples go up on a copy stream, results come back on a third stream
into pinned slots, and a small thread pool hands them to the caller's sink (a numpy array, a pinned tensor, or a
``.npy`` row writer) -- so uploads, kernels, downloads and file writes of different chunks all overlap.
"""
from __future__ import annotations

import time
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from heybuddy_b200 import _native, spec
from heybuddy_b200.dataset.augmented import AugmentedAudioGenerator
from heybuddy_b200.dataset.draws import DrawTable
from heybuddy_b200.embeddings import SpeechEmbeddings

__all__ = ["FeaturizePipeline", "RaggedClips", "DeviceChunk"]


@dataclass
class RaggedClips:
    """n int16 clips as one concatenated sample array + offsets (the shape TTS output arrives in)."""
    samples: np.ndarray   # int16 [total]
    offsets: np.ndarray   # int64 [n + 1]
    pinned: Optional["object"] = None  # optional torch int16 tensor over the same samples in pinned host memory

    @classmethod
    def from_list(cls, clips: Sequence[np.ndarray]) -> "RaggedClips":
        lengths = np.array([c.shape[0] for c in clips], dtype=np.int64)
        for c in clips:
            if c.dtype != np.int16:
                raise TypeError("RaggedClips holds int16 clips (Piper's output type, piper/pretrained.py:406-408)")
        return cls(np.concatenate(clips) if len(clips) else np.zeros(0, np.int16),
                   np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64))

    def __len__(self) -> int:
        return int(self.offsets.shape[0] - 1)

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.offsets)

    def slice(self, lo: int, hi: int) -> "RaggedClips":
        o = self.offsets[lo:hi + 1]
        return RaggedClips(self.samples[o[0]:o[-1]], (o - o[0]).astype(np.int64),
                           self.pinned[int(o[0]):int(o[-1])] if self.pinned is not None else None)

    def pin(self) -> "RaggedClips":
        """Moves the samples into pinned host memory so uploads need no staging copy."""
        import torch

        t = torch.from_numpy(self.samples).pin_memory()
        return RaggedClips(t.numpy(), self.offsets, t)


@dataclass
class DeviceChunk:
    """One chunk's device-resident inputs (what ``value`` in bench.py times over)."""
    samples: "object"       # cuda int16 [total]
    offsets: "object"       # cuda int64 [n + 1]
    pad_before: "object"    # cuda int32 [n]
    params: "object"        # cuda uint8 [n, 32]  (hb_clip_aug records)
    colored_ids: Optional["object"]     # cuda int64 [k] global batch ids of the coloured batches, or None
    colored_f_decay: Optional["object"]  # cuda f32 [k]
    n: int
    seed: int = 0
    k9: Optional[dict] = None           # K9Draws.pack(): name -> device tensor (ps_ratios / ps_counts stay host arrays)


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class _nvtx:
    """NVTX range around a pipeline stage (tracing) -- shows up in nsys / `ncu --nvtx`; a no-op without a profiler."""

    def __init__(self, name: str) -> None:
        self.name = name

    def __enter__(self):
        import torch

        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        import torch

        torch.cuda.nvtx.range_pop()


class FeaturizePipeline:
    def __init__(self, augment: AugmentedAudioGenerator, speech: SpeechEmbeddings, device_id: Optional[int] = None):
        if augment is None:
            raise ValueError("FeaturizePipeline needs an AugmentedAudioGenerator (set every probability to 0 for a plain length fix)")
        self.augment = augment
        self.speech = speech
        self.device = _native.require_cuda(device_id if device_id is not None else speech.spectrogram.device_id)
        self.t = augment.target_num_samples
        self.slot_offsets = np.asarray(spec.embedding_frame_offsets(self.t), dtype=np.int32)
        self._bufs: Dict[Tuple[str, int], "object"] = {}
        self.stage_ms: Dict[str, float] = {}
        self._events: List[Tuple[str, object, object]] = []
        self.profile = False
        self.last_stream_wait_s = 0.0
        self._writers: Optional[ThreadPoolExecutor] = None

    # -- host-side packing -------------------------------------------------------------------------
    def pack_params(self, table: DrawTable) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Draw table -> (pad_before i32[n], hb_clip_aug records [n], coloured batch ids i64[k], their f_decay f32[k])."""
        params = self.augment.clip_params(table)
        _, ids, f_decay = table.colored_slots()
        return np.ascontiguousarray(table.pad_before, dtype=np.int32), params, ids, f_decay

    @staticmethod
    def _k9_pack(table: DrawTable) -> Optional[Dict[str, np.ndarray]]:
        """The table's packed K9 draws (built once per table: the band-stop FIRs are designed here), or None when nothing was drawn."""
        if table.k9 is None or not table.k9.any():
            return None
        pk = getattr(table, "_k9_packed", None)
        if pk is None:
            pk = table.k9.pack()
            table._k9_packed = pk
        return pk

    def _meta_layout(self, n: int, k: int, k9: Optional[Dict[str, np.ndarray]] = None) -> Tuple[Dict[str, Tuple[int, int]], int]:
        """
        Byte ranges of the per-chunk metadata blob: offsets i64[n+1], pads i32[n], params [n][32], ids i64[k], f_decay f32[k] and,
        when K9 transforms were drawn, the device-side arrays of ``K9Draws.pack()`` (index lists, biquads, band-stop FIR rows).
        """
        from heybuddy_b200.dataset.k9 import HOST_ONLY

        fields = [("offsets", 8 * (n + 1)), ("pads", 4 * n), ("params", 32 * n), ("ids", 8 * k), ("fd", 4 * k)]
        if k9 is not None:
            fields += [(name, int(arr.nbytes)) for name, arr in k9.items() if name not in HOST_ONLY]
        lay, at = {}, 0
        for name, nbytes in fields:
            lay[name] = (at, nbytes)
            at = _align(at + nbytes)
        return lay, max(at, 256)

    def _fill_meta(self, blob: np.ndarray, clips: RaggedClips, table: DrawTable) -> Tuple[Dict[str, Tuple[int, int]], int, int]:
        from heybuddy_b200.dataset.k9 import HOST_ONLY

        pads, params, ids, fd = self.pack_params(table)
        n, k = len(clips), int(ids.shape[0])
        assert params.shape[0] == n == pads.shape[0], (params.shape, pads.shape, n)
        k9 = self._k9_pack(table)
        lay, total = self._meta_layout(n, k, k9)
        assert blob.nbytes >= total
        arrays = [("offsets", clips.offsets), ("pads", pads), ("params", params), ("ids", ids), ("fd", fd)]
        if k9 is not None:
            arrays += [(name, arr) for name, arr in k9.items() if name not in HOST_ONLY]
        for name, arr in arrays:
            at, nbytes = lay[name]
            if nbytes:
                blob[at:at + nbytes] = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        return lay, total, k

    @staticmethod
    def _chunk_from_meta(samples_dev, meta_dev, lay, n: int, k: int, seed: int, k9_host: Optional[Dict[str, np.ndarray]] = None) -> DeviceChunk:
        import torch

        from heybuddy_b200.dataset.k9 import HOST_ONLY

        def view(name, dtype, shape):
            at, nbytes = lay[name]
            return meta_dev[at:at + nbytes].view(dtype).view(shape)

        k9 = None
        if k9_host is not None:
            k9 = {name: (arr if name in HOST_ONLY else view(name, torch.from_numpy(arr[:0]).dtype, arr.shape)) for name, arr in k9_host.items()}
        return DeviceChunk(
            samples=samples_dev, offsets=view("offsets", torch.int64, (n + 1,)), pad_before=view("pads", torch.int32, (n,)),
            params=view("params", torch.uint8, (n, 32)), colored_ids=view("ids", torch.int64, (k,)) if k else None,
            colored_f_decay=view("fd", torch.float32, (k,)) if k else None, n=n, seed=seed, k9=k9)

    def upload(self, clips: RaggedClips, table: DrawTable) -> DeviceChunk:
        """Synchronous upload of one chunk (tests, the bench's device-resident leg)."""
        import torch

        n = len(clips)
        _, total = self._meta_layout(n, int(np.count_nonzero(table.colored_apply)), self._k9_pack(table))
        blob = np.zeros(total, dtype=np.uint8)
        lay, total, k = self._fill_meta(blob, clips, table)
        dev = self.device
        return self._chunk_from_meta(torch.from_numpy(np.ascontiguousarray(clips.samples)).to(dev), torch.from_numpy(blob).to(dev),
                                     lay, n, k, table.seed, self._k9_pack(table))

    # -- device path -----------------------------------------------------------------------------------
    def _buf(self, name: str, shape, dtype):
        import torch

        key = (name, int(np.prod(shape)))
        buf = self._bufs.get(key)
        if buf is None or buf.dtype != dtype:
            for k in [k for k in self._bufs if k[0] == name]:
                del self._bufs[k]
            buf = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[key] = buf
        return buf.view(shape)

    def _grow(self, name: str, numel: int, dtype):
        """A device buffer of at least ``numel`` elements that only ever grows (sizes that change per chunk)."""
        import torch

        buf = self._bufs.get((name, -1))
        if buf is None or buf.numel() < numel or buf.dtype != dtype:
            buf = torch.empty(max(int(numel * 1.25), 1), dtype=dtype, device=self.device)
            self._bufs[(name, -1)] = buf
        return buf

    def _mark(self, name: str):
        import torch

        if not self.profile:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        self._events.append((name, ev))
        return ev

    def collect_stage_times(self) -> Dict[str, float]:
        """Sums CUDA-event durations per stage since the last call (events are on the compute stream)."""
        out: Dict[str, float] = {}
        for (n0, e0), (n1, e1) in zip(self._events[:-1], self._events[1:]):
            if n1 == "begin":
                continue
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        self._events = []
        return out

    def run_device(self, chunk: DeviceChunk, out=None, keep_audio: bool = False):
        """Device-resident chunk -> cuda f32 ``[n, 16, 96]`` embeddings (optionally also the augmented audio)."""
        import torch

        lib = _native.load()
        n, t, dev = chunk.n, self.t, self.device
        aug = self.augment
        with torch.cuda.device(dev):
            st = _native.stream_ptr(dev)
            nb, rb = aug.noise_bank, aug.rir_bank
            self._mark("begin")
            bases_ptr = None
            if chunk.colored_ids is not None:
                k = int(chunk.colored_ids.numel())
                bases = self._grow("bases", k * spec.COLORED_BASE_SAMPLES, torch.float32)
                with _nvtx("hb/colored_bases"):
                    _native.check(lib.hb_colored_bases(chunk.seed & (2 ** 64 - 1), chunk.colored_ids.data_ptr(), chunk.colored_f_decay.data_ptr(),
                                                       k, bases.data_ptr(), st), "hb_colored_bases")
                bases_ptr = bases.data_ptr()
            self._mark("colored")
            banks = (nb.stream.data_ptr() if nb is not None else None, bases_ptr, rb.spec.data_ptr() if rb is not None else None)
            if t == spec.CLIP_SAMPLES and not keep_audio and not self.profile and chunk.k9 is None:
                # one C-ABI call for the whole path (hb_featurize_i16 = augment_i16 -> mel -> embed on one workspace)
                emb_model = self.speech.embeddings
                if not emb_model.loaded:
                    emb_model.load()
                _native.ensure_tables(dev)
                offs = np.ascontiguousarray(self.slot_offsets, dtype=np.int32)
                if out is None:
                    out = torch.empty((n, offs.size, spec.EMB_DIM), dtype=torch.float32, device=dev)
                nbytes = lib.hb_featurize_workspace_bytes(n, t, emb_model.mode)
                _native.check(nbytes, "hb_featurize_workspace_bytes")
                ws = self._grow("featurize_ws", int(nbytes), torch.uint8)
                with _nvtx(f"hb/featurize_i16[{n}]"):
                    _native.check(lib.hb_featurize_i16(
                        emb_model._handle, emb_model.mode, chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(),
                        *banks, chunk.params.data_ptr(), offs.ctypes.data, offs.size,
                        out.data_ptr(), n, t, ws.data_ptr(), int(ws.numel()), st), "hb_featurize_i16")
                return out
            audio = self._buf("audio", (n, t), torch.float32) if not keep_audio else torch.empty((n, t), dtype=torch.float32, device=dev)
            if t == spec.CLIP_SAMPLES and chunk.k9 is None:
                # length fix fused into the augmentation kernel's load: int16 samples -> shared memory, no f32 intermediate
                _native.check(lib.hb_augment_clips_i16(chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(),
                                                       *banks, chunk.params.data_ptr(), audio.data_ptr(), n, t, st), "hb_augment_clips_i16")
            else:
                fixed = self._buf("fixed", (n, t), torch.float32)
                _native.check(lib.hb_fix_length_i16(chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(),
                                                    fixed.data_ptr(), n, t, st), "hb_fix_length_i16")
                if chunk.k9 is not None:
                    # K9: the reference's per-clip numpy transforms (augmented.py:325-328: SevenBandParametricEQ, TanhDistortion) and
                    # the head of its batch Compose (:369-372: PitchShift, BandStopFilter), in place on the selected clips
                    from heybuddy_b200.dataset.k9 import apply_packed

                    with _nvtx("hb/k9"):
                        apply_packed(fixed, chunk.k9, lambda name, numel, dtype: self._grow("k9_" + name, numel, dtype))
                    self._mark("k9")
                _native.check(lib.hb_augment_clips_f32(fixed.data_ptr(), *banks, chunk.params.data_ptr(), audio.data_ptr(), n, t, st),
                              "hb_augment_clips_f32")
            self._mark("augment")
            mel = self._buf("mel", (n, spec.mel_frames(t), spec.N_MELS), torch.float32)
            with _nvtx("hb/mel"):
                self.speech.spectrogram.run_device(audio, scale=spec.AUDIO_SCALE, out=mel)
            self._mark("mel")
            if out is None:
                out = torch.empty((n, self.slot_offsets.size, spec.EMB_DIM), dtype=torch.float32, device=dev)
            with _nvtx("hb/embed"):
                self.speech.embeddings.run_clips_device(mel, self.slot_offsets, out=out)
            self._mark("embed")
        return (out, audio) if keep_audio else out

    def run_fused_front(self, chunk: DeviceChunk, mel_out=None):
        """Only the production-mode front end: ``hb_colored_bases`` + ``hb_augment_mel_i16`` (length fix + augmentation + mel in ONE
        kernel) -> cuda f32 ``[n, 141, 32]``.  Used by the parity test against the staged pair and by bench.py's stage timing."""
        import torch

        lib = _native.load()
        n, t, dev = chunk.n, self.t, self.device
        aug = self.augment
        _native.ensure_tables(dev)
        with torch.cuda.device(dev):
            st = _native.stream_ptr(dev)
            nb, rb = aug.noise_bank, aug.rir_bank
            bases_ptr = None
            if chunk.colored_ids is not None:
                k = int(chunk.colored_ids.numel())
                bases = self._grow("bases", k * spec.COLORED_BASE_SAMPLES, torch.float32)
                _native.check(lib.hb_colored_bases(chunk.seed & (2 ** 64 - 1), chunk.colored_ids.data_ptr(), chunk.colored_f_decay.data_ptr(),
                                                   k, bases.data_ptr(), st), "hb_colored_bases")
                bases_ptr = bases.data_ptr()
            if mel_out is None:
                mel_out = torch.empty((n, spec.mel_frames(t), spec.N_MELS), dtype=torch.float32, device=dev)
            _native.check(lib.hb_augment_mel_i16(
                chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(), nb.stream.data_ptr() if nb is not None else None,
                bases_ptr, rb.spec.data_ptr() if rb is not None else None, chunk.params.data_ptr(), float(spec.AUDIO_SCALE), mel_out.data_ptr(),
                n, t, st), "hb_augment_mel_i16")
        return mel_out

    # -- host path (the call a user makes) -----------------------------------------------------------------
    def featurize_host(self, clips: RaggedClips, tables: Union[DrawTable, Sequence[DrawTable]], chunk_clips: int,
                       out: Optional[np.ndarray] = None):
        """
        Host int16 clips -> host f32 ``[n, 16, 96]``.  ``tables``: the draw table of all the clips (or one table per chunk of
        ``chunk_clips`` clips, a multiple of the augmentation batch size).  ``out`` may be a numpy array, a pinned CPU torch
        tensor (the D2H then lands in it directly) or a callable ``(lo, hi, rows)``.  Returns (embeddings, h2d_bytes, d2h_bytes).
        """
        n_slots = self.slot_offsets.size
        if out is None:
            out = np.empty((len(clips), n_slots, spec.EMB_DIM), dtype=np.float32)
        h2d, d2h = self.featurize_stream([(clips, tables, out)], chunk_clips)
        return (out.numpy() if hasattr(out, "is_pinned") else out), h2d, d2h

    def featurize_stream(self, items, chunk_clips: int, writer_threads: int = 4, wait: bool = True):
        """
        Streaming form of :meth:`featurize_host`: ``items`` is an iterable (pulled lazily, e.g. a generator that fetches one
        super-batch of source clips at a time) of ``(clips, table, sink)`` host datasets.
        ``table``: a :class:`DrawTable` covering the item's clips (sliced per chunk here) or a list of per-chunk tables.
        ``sink``: numpy array / pinned CPU torch tensor ``[len(clips), 16, 96]``, or a callable ``sink(lo, hi, rows)`` that is
        handed every finished row range (``rows``: f32 numpy view of a pinned slot, valid until the call returns; called from
        worker threads, possibly out of order).  The upload of the next chunk -- of the same or of the next item -- always
        overlaps the compute of the current one, so a long run pays the pipeline fill (first H2D) and drain (last D2H) once,
        not once per item.  Returns (h2d_bytes, d2h_bytes); with ``wait=False`` the call returns as soon as the last chunk is
        enqueued and ``self.finish_stream()`` waits for the downloads and the sinks (lets the caller start the next dataset's
        kernels while this one's rows are still being written).
        """
        import torch

        dev = self.device
        n_slots = self.slot_offsets.size
        b = self.augment.batch_size
        self.finish_stream()          # a previous deferred call must have drained: it owns the pinned output slots
        assert chunk_clips % b == 0, "chunks are whole augmentation batches"
        for key in ("copy_stream", "d2h_stream"):
            if (key, 0) not in self._bufs:
                self._bufs[(key, 0)] = torch.cuda.Stream(device=dev)
        copy_stream = self._bufs[("copy_stream", 0)]
        d2h_stream = self._bufs[("d2h_stream", 0)]     # results drain on their own stream, off the compute stream's critical path
        compute = torch.cuda.current_stream(dev)
        if self._writers is None or getattr(self._writers, "_max_workers", 0) != writer_threads:
            self._writers = ThreadPoolExecutor(max_workers=writer_threads, thread_name_prefix="hb-sink")
        h2d = d2h = 0
        # The host stages up to DEPTH chunks ahead of the chunk whose kernels it is enqueueing (the metadata goes through N_SLOTS
        # pinned staging blobs; the samples come straight from the caller's pinned memory when they are pinned), so a host hiccup
        # of a few milliseconds does not starve the GPU.  At most MAX_INFLIGHT chunks are uploaded-but-not-finished at any time
        # (bounds the device memory of a long run: uploads are faster than the kernels and would otherwise run ahead without limit).
        DEPTH, N_SLOTS, MAX_INFLIGHT, OUT_SLOTS = 3, 4, 5, 4
        stage_events = [None] * N_SLOTS
        out_tasks = [None] * OUT_SLOTS
        wait_s = 0.0    # host time spent blocked on the device (the rest of the call is host work)
        stats = {"items_s": 0.0, "stage_s": 0.0, "enqueue_s": 0.0, "wait_slot_s": 0.0, "wait_sink_s": 0.0, "sink_s": 0.0, "chunks": 0}

        def work_units():
            """(sink, clips of the item, lo, hi, table of the chunk) for every chunk of every item, items pulled lazily."""
            for clips, tables, out in items:
                n = len(clips)
                if hasattr(out, "is_pinned"):
                    assert out.is_pinned() and tuple(out.shape) == (n, n_slots, spec.EMB_DIM) and out.dtype == torch.float32
                elif isinstance(out, np.ndarray):
                    assert tuple(out.shape) == (n, n_slots, spec.EMB_DIM) and out.dtype == np.float32
                else:
                    assert callable(out), "sink: numpy array, pinned tensor or callable(lo, hi, rows)"
                n_chunks = (n + chunk_clips - 1) // chunk_clips
                per_chunk = isinstance(tables, (list, tuple))
                if per_chunk:
                    assert len(tables) >= n_chunks, "one draw table per chunk"
                else:
                    assert tables.n_clips == n, (tables.n_clips, n)
                for ci in range(n_chunks):
                    lo, hi = ci * chunk_clips, min(n, (ci + 1) * chunk_clips)
                    yield out, clips, lo, hi, (tables[ci] if per_chunk else tables.slice(lo // b, (hi + b - 1) // b))

        units = work_units()
        done_events = deque()   # compute-finished events of the chunks enqueued so far (bounded)
        n_staged = 0

        def blocked(ev):
            nonlocal wait_s
            if ev is not None and not ev.query():
                t0 = time.perf_counter()
                ev.synchronize()
                wait_s += time.perf_counter() - t0

        def pinned(name: str, slot: int, numel: int, dtype):
            key = (f"pin_{name}_{slot}", 0)
            pin = self._bufs.get(key)
            if pin is None or pin.numel() < numel or pin.dtype != dtype:
                pin = torch.empty(max(int(numel * 1.25), 1), dtype=dtype).pin_memory()
                self._bufs[key] = pin
            return pin

        def stage(unit):
            nonlocal h2d, n_staged
            sink, clips, lo, hi, table = unit
            part = clips.slice(lo, hi)
            n = hi - lo
            k = n_staged
            n_staged += 1
            slot = k % N_SLOTS
            blocked(stage_events[slot])          # the H2D that last read this slot's pinned staging buffers must have finished
            while len(done_events) > 0 and k - done_events[0][0] >= MAX_INFLIGHT:
                blocked(done_events.popleft()[1])
            torch.cuda.nvtx.range_push(f"hb/stage_chunk[{n}]")
            _, total = self._meta_layout(n, int(np.count_nonzero(table.colored_apply)), self._k9_pack(table))
            meta_pin = pinned("meta", slot, total, torch.uint8)
            lay, total, kc = self._fill_meta(meta_pin.numpy(), part, table)
            if part.pinned is not None:
                samples_pin = part.pinned        # already pinned: DMA straight from the caller's buffer
            else:
                samples_pin = pinned("samples", slot, part.samples.size, torch.int16)[:part.samples.size]
                np.copyto(samples_pin.numpy(), part.samples)
            with torch.cuda.stream(copy_stream):
                samples_dev = torch.empty(part.samples.size, dtype=torch.int16, device=dev)
                samples_dev.copy_(samples_pin, non_blocking=True)
                meta_dev = torch.empty(total, dtype=torch.uint8, device=dev)
                meta_dev.copy_(meta_pin[:total], non_blocking=True)
                h2d += part.samples.size * 2 + total
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            stage_events[slot] = ev
            torch.cuda.nvtx.range_pop()
            return self._chunk_from_meta(samples_dev, meta_dev, lay, n, kc, table.seed, self._k9_pack(table)), meta_dev, ev, sink, lo, hi

        def deliver(sink, lo, hi, pin, done_ev):
            done_ev.synchronize()
            t0 = time.perf_counter()
            rows = pin[:(hi - lo) * n_slots * spec.EMB_DIM].numpy().reshape(hi - lo, n_slots, spec.EMB_DIM)
            if isinstance(sink, np.ndarray):
                np.copyto(sink[lo:hi], rows)
            else:
                sink(lo, hi, rows)
            stats["sink_s"] += time.perf_counter() - t0     # summed over the worker threads

        staged = deque()
        exhausted = False

        def fill():
            nonlocal exhausted
            while not exhausted and len(staged) < DEPTH:
                t0 = time.perf_counter()
                unit = next(units, None)
                t1 = time.perf_counter()
                stats["items_s"] += t1 - t0
                if unit is None:
                    exhausted = True
                    break
                staged.append(stage(unit))
                stats["stage_s"] += time.perf_counter() - t1

        fill()
        k = -1
        while staged:
            k += 1
            chunk, meta_dev, ev, sink, lo, hi = staged.popleft()
            t_enq = time.perf_counter()
            compute.wait_event(ev)
            chunk.samples.record_stream(compute)
            meta_dev.record_stream(compute)
            emb = self.run_device(chunk)
            ready = torch.cuda.Event()
            ready.record(compute)
            done_events.append((k, ready))
            d2h_stream.wait_event(ready)
            emb.record_stream(d2h_stream)
            d2h += emb.numel() * 4
            if hasattr(sink, "is_pinned"):
                with torch.cuda.stream(d2h_stream):
                    sink[lo:hi].copy_(emb, non_blocking=True)
            else:
                oslot = k % OUT_SLOTS
                if out_tasks[oslot] is not None:   # the task that last read this pinned slot must have handed its rows over
                    t0 = time.perf_counter()
                    out_tasks[oslot].result()
                    wait_s += time.perf_counter() - t0
                    stats["wait_sink_s"] += time.perf_counter() - t0
                    t_enq += time.perf_counter() - t0
                pin = pinned("out", oslot, emb.numel(), torch.float32)
                with torch.cuda.stream(d2h_stream):
                    pin[:emb.numel()].copy_(emb.reshape(-1), non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(d2h_stream)
                out_tasks[oslot] = self._writers.submit(deliver, sink, lo, hi, pin, done)
            del chunk, emb, meta_dev
            stats["enqueue_s"] += time.perf_counter() - t_enq
            stats["chunks"] += 1
            fill()
        def finish():
            t0 = time.perf_counter()
            for task in out_tasks:
                if task is not None:
                    task.result()
            d2h_stream.synchronize()
            stats["wait_s"] = wait_s + time.perf_counter() - t0
            self.last_stream_wait_s = stats["wait_s"]   # host time spent blocked on the device / the sinks (diagnostic)

        self.last_stats = stats            # host-side time breakdown of the call (diagnostic; bench.py prints it)
        self._pending_finish = finish
        if wait:
            self.finish_stream()
        return h2d, d2h

    def finish_stream(self) -> None:
        """Waits for the downloads and sink tasks of the last ``featurize_stream(..., wait=False)``."""
        fin, self._pending_finish = getattr(self, "_pending_finish", None), None
        if fin is not None:
            fin()
