"""
The fused featurization pipeline: ragged int16 clips -> length fix -> augmentation -> log-mel ->
speech embeddings, everything resident on one GPU between the first H2D copy and the final D2H of
the ``[n, 16, 96]`` embeddings.

This is what ``TrainingFeaturesGenerator.generate`` (reference dataset/features.py:360-490) does
with a python loop per clip, a D2H per clip and one ORT call per 32 items; here a chunk of
augmentation batches is five kinds of kernel launches on one stream:

    hb_augment_clips_i16 (length fix + augmentation) -> hb_mel_f32 -> hb_embed_clips (trunk + tail)

``featurize_host`` adds the pinned-memory H2D / D2H copies on a second stream so chunk i+1 uploads
while chunk i computes (double buffering).
"""
from __future__ import annotations

import time
from collections import deque
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

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
    bases: Optional["object"]  # cuda f32 [k, 16000] coloured patterns or None
    n: int


class FeaturizePipeline:
    def __init__(self, augment: Optional[AugmentedAudioGenerator], speech: SpeechEmbeddings, device_id: Optional[int] = None):
        self.augment = augment
        self.speech = speech
        self.device = _native.require_cuda(device_id if device_id is not None else speech.spectrogram.device_id)
        self.t = spec.CLIP_SAMPLES if augment is None else augment.target_num_samples
        self.slot_offsets = np.asarray(spec.embedding_frame_offsets(self.t), dtype=np.int32)
        self._bufs: Dict[Tuple[str, int], "object"] = {}
        self.stage_ms: Dict[str, float] = {}
        self._events: List[Tuple[str, object, object]] = []
        self.profile = False

    # -- host-side packing -------------------------------------------------------------------------
    def pack_params(self, table: DrawTable) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
        """Draw table -> (pad_before i32[n], hb_clip_aug records, coloured bases f32[k,16000] | None)."""
        aug = self.augment
        bases = [d.colored_base for d in table.batches if d.colored_apply]
        slots, k = [], 0
        for d in table.batches:
            slots.append(k if d.colored_apply else -1)
            k += int(d.colored_apply)
        params = aug.clip_params(table.batches, table.noise_clip_cursor, table.rir_index, slots)
        pads = np.concatenate([d.pad_before for d in table.batches]).astype(np.int32)
        return pads, params, (np.stack(bases) if bases else None)

    def upload(self, clips: RaggedClips, table: DrawTable) -> DeviceChunk:
        import torch

        pads, params, bases = self.pack_params(table)
        dev = self.device
        return DeviceChunk(
            samples=torch.from_numpy(clips.samples).to(dev), offsets=torch.from_numpy(clips.offsets).to(dev),
            pad_before=torch.from_numpy(pads).to(dev),
            params=torch.from_numpy(params.view(np.uint8).reshape(len(clips), -1)).to(dev),
            bases=torch.from_numpy(bases).to(dev) if bases is not None else None, n=len(clips))

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
            if t == spec.CLIP_SAMPLES and not keep_audio and not self.profile:
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
                ws = self._buf("featurize_ws", (int(nbytes),), torch.uint8)
                _native.check(lib.hb_featurize_i16(
                    emb_model._handle, emb_model.mode, chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(),
                    nb.stream.data_ptr() if nb is not None else None, chunk.bases.data_ptr() if chunk.bases is not None else None,
                    rb.spec.data_ptr() if rb is not None else None, chunk.params.data_ptr(), offs.ctypes.data, offs.size,
                    out.data_ptr(), n, t, ws.data_ptr(), int(nbytes), st), "hb_featurize_i16")
                return out
            self._mark("begin")
            audio = self._buf("audio", (n, t), torch.float32) if not keep_audio else torch.empty((n, t), dtype=torch.float32, device=dev)
            banks = (nb.stream.data_ptr() if nb is not None else None, chunk.bases.data_ptr() if chunk.bases is not None else None,
                     rb.spec.data_ptr() if rb is not None else None)
            if t == spec.CLIP_SAMPLES:
                # length fix fused into the augmentation kernel's load: int16 samples -> shared memory, no f32 intermediate
                _native.check(lib.hb_augment_clips_i16(chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(),
                                                       *banks, chunk.params.data_ptr(), audio.data_ptr(), n, t, st), "hb_augment_clips_i16")
            else:
                fixed = self._buf("fixed", (n, t), torch.float32)
                _native.check(lib.hb_fix_length_i16(chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(),
                                                    fixed.data_ptr(), n, t, st), "hb_fix_length_i16")
                _native.check(lib.hb_augment_clips_f32(fixed.data_ptr(), *banks, chunk.params.data_ptr(), audio.data_ptr(), n, t, st),
                              "hb_augment_clips_f32")
            self._mark("augment")
            mel = self._buf("mel", (n, spec.mel_frames(t), spec.N_MELS), torch.float32)
            self.speech.spectrogram.run_device(audio, scale=spec.AUDIO_SCALE, out=mel)
            self._mark("mel")
            if out is None:
                out = torch.empty((n, self.slot_offsets.size, spec.EMB_DIM), dtype=torch.float32, device=dev)
            self.speech.embeddings.run_clips_device(mel, self.slot_offsets, out=out)
            self._mark("embed")
        return (out, audio) if keep_audio else out

    # -- host path (the call a user makes) -----------------------------------------------------------------
    def featurize_host(self, clips: RaggedClips, tables: Sequence[DrawTable], chunk_clips: int, out: Optional[np.ndarray] = None):
        """
        Host int16 clips -> host f32 ``[n, 16, 96]``.  ``tables[i]`` holds the draws of chunk i
        (``chunk_clips`` clips each, a multiple of the augmentation batch size).  Uploads run on a side
        stream from pinned staging buffers, double-buffered against the compute stream; results drain on a
        third stream.  ``out`` may be a numpy array or a pinned CPU torch tensor (the D2H then lands in it
        directly).  Returns (embeddings, h2d_bytes, d2h_bytes).
        """
        n_slots = self.slot_offsets.size
        if out is None:
            out = np.empty((len(clips), n_slots, spec.EMB_DIM), dtype=np.float32)
        h2d, d2h = self.featurize_stream([(clips, tables, out)], chunk_clips)
        return (out.numpy() if hasattr(out, "is_pinned") else out), h2d, d2h

    def featurize_stream(self, items, chunk_clips: int):
        """
        Streaming form of :meth:`featurize_host`: ``items`` is a sequence of ``(clips, tables, out)`` host datasets
        (``out``: numpy array or pinned CPU torch tensor ``[len(clips), 16, 96]``).  The upload of the next chunk --
        of the same or of the next item -- always overlaps the compute of the current one, so a long run pays the
        pipeline fill (first H2D) and drain (last D2H) once, not once per item.  Returns (h2d_bytes, d2h_bytes).
        """
        import torch

        dev = self.device
        n_slots = self.slot_offsets.size
        for key in ("copy_stream", "d2h_stream"):
            if (key, 0) not in self._bufs:
                self._bufs[(key, 0)] = torch.cuda.Stream(device=dev)
        copy_stream = self._bufs[("copy_stream", 0)]
        d2h_stream = self._bufs[("d2h_stream", 0)]     # results drain on their own stream, off the compute stream's critical path
        compute = torch.cuda.current_stream(dev)
        h2d = d2h = 0
        # The host stages up to DEPTH chunks ahead of the chunk whose kernels it is enqueueing (small draw-table arrays go through
        # N_SLOTS pinned staging buffers; the samples come straight from the caller's pinned memory), so a host hiccup of a few
        # milliseconds does not starve the GPU.  At most MAX_INFLIGHT chunks are uploaded-but-not-finished at any time (bounds the
        # device memory of a long run: uploads are faster than the kernels and would otherwise run ahead without limit).
        DEPTH, N_SLOTS, MAX_INFLIGHT = 3, 4, 5
        stage_events = [None] * N_SLOTS
        wait_s = 0.0    # host time spent blocked on the device (the rest of the call is host work)

        work = []   # (item index, chunk index, lo, hi)
        sinks = []  # per item: (numpy view, pinned torch tensor or None)
        for ii, (clips, tables, out) in enumerate(items):
            n = len(clips)
            out_t = out if hasattr(out, "is_pinned") else None
            if out_t is not None:
                assert out_t.is_pinned() and tuple(out_t.shape) == (n, n_slots, spec.EMB_DIM) and out_t.dtype == torch.float32
            else:
                assert tuple(out.shape) == (n, n_slots, spec.EMB_DIM) and out.dtype == np.float32
            sinks.append((out if out_t is None else out_t.numpy(), out_t))
            n_chunks = (n + chunk_clips - 1) // chunk_clips
            assert len(tables) >= n_chunks, "one draw table per chunk"
            for ci in range(n_chunks):
                work.append((ii, ci, ci * chunk_clips, min(n, (ci + 1) * chunk_clips)))
        done_events = [None] * len(work)   # compute of chunk k finished

        def blocked(ev):
            nonlocal wait_s
            if ev is not None and not ev.query():
                t0 = time.perf_counter()
                ev.synchronize()
                wait_s += time.perf_counter() - t0

        def stage(k: int):
            nonlocal h2d
            ii, ci, lo, hi = work[k]
            clips, tables, _ = items[ii]
            part = clips.slice(lo, hi)
            pads, params, bases = self.pack_params(tables[ci])
            slot = k % N_SLOTS
            blocked(stage_events[slot])          # the H2D that last read this slot's pinned staging buffers must have finished
            if k >= MAX_INFLIGHT:
                blocked(done_events[k - MAX_INFLIGHT])
            host = {}
            for name, arr in (("samples", part.samples), ("offsets", part.offsets), ("pads", pads),
                              ("params", params.view(np.uint8).reshape(hi - lo, -1)), ("bases", bases)):
                if arr is None:
                    host[name] = None
                    continue
                if name == "samples" and part.pinned is not None:
                    host[name] = (part.pinned, arr.shape)  # already pinned: DMA straight from the caller's buffer
                    continue
                key = (f"pin_{name}_{slot}", 0)
                pin = self._bufs.get(key)
                if pin is None or pin.numel() < arr.size or pin.dtype != torch.from_numpy(arr[:0]).dtype:
                    pin = torch.empty(max(arr.size, 1), dtype=torch.from_numpy(arr[:0]).dtype).pin_memory()
                    self._bufs[key] = pin
                pin[:arr.size].copy_(torch.from_numpy(np.ascontiguousarray(arr).reshape(-1)))
                host[name] = (pin, arr.shape)
            with torch.cuda.stream(copy_stream):
                devs = {}
                for name, v in host.items():
                    if v is None:
                        devs[name] = None
                        continue
                    pin, shape = v
                    cnt = int(np.prod(shape))
                    d = torch.empty(cnt, dtype=pin.dtype, device=dev)
                    d.copy_(pin[:cnt], non_blocking=True)
                    h2d += cnt * pin.element_size()
                    devs[name] = d.view(shape)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            stage_events[slot] = ev
            return DeviceChunk(devs["samples"], devs["offsets"], devs["pads"], devs["params"], devs["bases"], hi - lo), ev

        pending = None  # staged path: (pinned buffer, numpy sink, lo, hi, event)

        def drain():
            nonlocal pending
            if pending is not None:
                p_pin, p_sink, p_lo, p_hi, p_ev = pending
                blocked(p_ev)
                p_sink[p_lo:p_hi] = p_pin[:(p_hi - p_lo) * n_slots * spec.EMB_DIM].numpy().reshape(p_hi - p_lo, n_slots, spec.EMB_DIM)
                pending = None

        staged = deque()
        next_stage = 0

        def fill():
            nonlocal next_stage
            while next_stage < len(work) and len(staged) < DEPTH:
                staged.append(stage(next_stage))
                next_stage += 1

        fill()
        for k in range(len(work)):
            chunk, ev = staged.popleft()
            ii, ci, lo, hi = work[k]
            compute.wait_event(ev)
            for v in (chunk.samples, chunk.offsets, chunk.pad_before, chunk.params, chunk.bases):
                if v is not None:
                    v.record_stream(compute)
            emb = self.run_device(chunk)
            ready = torch.cuda.Event()
            ready.record(compute)
            done_events[k] = ready
            if k >= MAX_INFLIGHT + 1:
                done_events[k - MAX_INFLIGHT - 1] = None
            d2h_stream.wait_event(ready)
            emb.record_stream(d2h_stream)
            sink, out_t = sinks[ii]
            d2h += emb.numel() * 4
            if out_t is not None:
                with torch.cuda.stream(d2h_stream):
                    out_t[lo:hi].copy_(emb, non_blocking=True)
                del chunk, emb
                fill()
                continue
            key = (f"pin_out_{k % 2}", 0)
            pin = self._bufs.get(key)
            if pin is None or pin.numel() < emb.numel():
                pin = torch.empty(emb.numel(), dtype=torch.float32).pin_memory()
                self._bufs[key] = pin
            drain()   # the previous chunk's pinned buffer (the other slot) -> its numpy sink
            with torch.cuda.stream(d2h_stream):
                pin[:emb.numel()].copy_(emb.reshape(-1), non_blocking=True)
                done = torch.cuda.Event()
                done.record(d2h_stream)
            pending = (pin, sink, lo, hi, done)
            del chunk, emb
            fill()
        drain()
        t0 = time.perf_counter()
        d2h_stream.synchronize()
        wait_s += time.perf_counter() - t0
        self.last_stream_wait_s = wait_s   # host time spent blocked on the device during the call (diagnostic)
        return h2d, d2h
