"""
``SpeechEmbeddings`` / ``SpeechEmbeddingModel`` -- B200 replacement of reference
``heybuddy/embeddings.py``.

Public surface and semantics follow the reference (embeddings.py:23-243): audio -> sliding
17280-sample windows (stride 1920) -> 32-bin log-mel -> 76-frame windows (stride 8) -> 96-d
embeddings, windows concatenated on axis 1, NaN repair, optional spectrogram return.

What differs underneath: ``__call__`` keeps the whole batch on the GPU, computes ONE mel per
clip (frame f of audio window w is global frame 12 w + f) and evaluates the conv
stack fully convolutionally once per clip (``hb_embed_clips``), writing the embeddings in the
reference's slot order.  The reference's per-window methods ``audio_to_spectrograms`` /
``spectrograms_to_embeddings`` are kept with their signatures and run the same kernels.

Reference quirks fixed on purpose: the caller's tensor is not scaled in
place; a trailing embed batch of size 1 keeps its batch axis; NaN repair is seedable.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np

from heybuddy_b200 import _native, spec
from heybuddy_b200.spectrogram import MelSpectrogramModel
from heybuddy_b200.util import PretrainedNativeModel, audio_to_bct_tensor, logger

__all__ = ["SpeechEmbeddingModel", "SpeechEmbeddings", "get_speech_embeddings", "DEFAULT_EMBED_PRECISION"]

DEFAULT_EMBED_PRECISION = os.environ.get("HEYBUDDY_B200_EMBED_PRECISION", "f16")
# clips per device chunk of the fused path (bounds the activation workspace)
DEFAULT_CLIP_CHUNK = int(os.environ.get("HEYBUDDY_B200_CLIP_CHUNK", "1024"))


RANDOM_INIT = "random-init"
_warned_random_init = False


def _warn_random_init() -> None:
    global _warned_random_init
    if not _warned_random_init:
        _warned_random_init = True
        logger.warning(
            "SpeechEmbeddingModel: NO pretrained weights given -- using the seeded RANDOM init of the embedding CNN. Embeddings from "
            "an untrained network are not interchangeable with the reference's speech-embedding.onnx features: classifiers trained "
            "on them, and reference .pt checkpoints scored with them, are meaningless. Pass weights=<dict | .npz path>, set "
            "HEYBUDDY_B200_EMBED_WEIGHTS, or opt in with weights='random-init' / HEYBUDDY_B200_ALLOW_RANDOM_INIT=1.")


def pack_embedding_weights(weights: Dict[str, np.ndarray]) -> np.ndarray:
    """Layer-table order, each kernel ``[kh, kw, cin, cout]`` then its bias (``hb_embed_create`` layout)."""
    parts = []
    for (name, kh, kw, cin, cout, *_r) in spec.EMBEDDING_LAYERS:
        w = np.asarray(weights[f"{name}.weight"], dtype=np.float32)
        b = np.asarray(weights[f"{name}.bias"], dtype=np.float32)
        assert w.shape == (kh, kw, cin, cout), (name, w.shape)
        assert b.shape == (cout,), (name, b.shape)
        parts += [w.ravel(), b.ravel()]
    return np.ascontiguousarray(np.concatenate(parts))


class SpeechEmbeddingModel(PretrainedNativeModel):
    """
    Compute speech embeddings from spectrograms: f32 ``[n, 76, 32, 1]`` (input name ``input_1``)
    -> ``conv2d_19`` ``[n, 1, 1, 96]``; ``__call__`` returns ``[n, 96]`` (embeddings.py:32-42).

    ``precision``: ``"f16"`` = tcgen05 implicit GEMM, fp16 operands / fp32 TMEM accumulation
    (TF32-class mantissa; tolerance stated in tests/test_embed_gpu.py), ``"fp32"`` = CUDA-core
    parity mode.

    Weights: ``weights`` = a dict of ``<layer>.weight`` (HWIO) / ``<layer>.bias`` arrays or the path of an ``.npz`` holding
    them (``from_file(path)`` and the environment variable ``HEYBUDDY_B200_EMBED_WEIGHTS`` do the same);
    ``python -m heybuddy_b200.util.onnx_wire speech-embedding.onnx out.npz`` converts the reference's artefact
    (embeddings.py:29-30) once it is at hand.  The artefact cannot be downloaded offline, so with NO weights given the model
    falls back to the seeded RANDOM init of ``spec.init_embedding_weights`` -- an untrained CNN whose features are not
    interchangeable with the reference's -- and says so with a warning unless the caller opted in
    (``weights="random-init"`` or ``HEYBUDDY_B200_ALLOW_RANDOM_INIT=1``; tests and the bench do).
    """

    input_name = "input_1"

    def __init__(self, device_id: Optional[int] = None, load: bool = False, precision: Optional[str] = None,
                 weights: Optional[Union[Dict[str, np.ndarray], str]] = None) -> None:
        self.precision = precision or DEFAULT_EMBED_PRECISION
        if self.precision not in _native.EMBED_MODES:
            raise ValueError(f"precision must be one of {sorted(_native.EMBED_MODES)}, got {self.precision!r}")
        self._weights = weights
        self._handle = None
        self._workspace = None
        super().__init__(device_id=device_id, load=load)

    @property
    def mode(self) -> int:
        return _native.EMBED_MODES[self.precision]

    def _load(self) -> None:
        import ctypes

        import torch

        weights = self._weights
        path = weights if isinstance(weights, str) and weights != RANDOM_INIT else None
        path = path or self.pretrained_model_path or (os.environ.get("HEYBUDDY_B200_EMBED_WEIGHTS") if weights is None else None)
        if path:
            with np.load(path) as z:
                weights = {k: z[k] for k in z.files}
        elif weights is None or isinstance(weights, str):
            if weights != RANDOM_INIT and os.environ.get("HEYBUDDY_B200_ALLOW_RANDOM_INIT", "") not in ("1", "true", "yes"):
                _warn_random_init()
            weights = spec.init_embedding_weights()
        packed = pack_embedding_weights(weights)
        lib = _native.load()
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(lib.hb_embed_create(ctypes.byref(handle), packed.ctypes.data, packed.size), "hb_embed_create")
        self._handle = handle

    def _unload(self) -> None:
        if self._handle is not None:
            _native.load().hb_embed_destroy(self._handle)
            self._handle = None
        self._workspace = None

    def __del__(self) -> None:
        try:
            self._unload()
        except Exception:
            pass

    def _get_workspace(self, nbytes: int, device):
        import torch

        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != device:
            self._workspace = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        return self._workspace

    # -- device-level entries -------------------------------------------------------------------
    def run_clips_device(self, mel_dev, slot_offsets, out=None):
        """cuda f32 ``[B, F, 32]`` -> cuda f32 ``[B, n_slots, 96]``; slot s = frames [off[s], off[s]+76)."""
        import torch

        if not self.loaded:
            self.load()
        assert mel_dev.is_cuda and mel_dev.dtype == torch.float32 and mel_dev.is_contiguous()
        b, f, m = mel_dev.shape
        assert m == spec.N_MELS
        offs = np.ascontiguousarray(slot_offsets, dtype=np.int32)
        if out is None:
            out = torch.empty((b, offs.size, spec.EMB_DIM), dtype=torch.float32, device=mel_dev.device)
        lib = _native.load()
        with torch.cuda.device(mel_dev.device):
            nbytes = lib.hb_embed_clips_workspace_bytes(b, f, self.mode)
            _native.check(nbytes, "hb_embed_clips_workspace_bytes")
            ws = self._get_workspace(nbytes, mel_dev.device)
            _native.check(
                lib.hb_embed_clips(self._handle, self.mode, mel_dev.data_ptr(), b, f, offs.ctypes.data, offs.size,
                                   out.data_ptr(), ws.data_ptr(), ws.numel(), _native.stream_ptr(mel_dev.device)),
                "hb_embed_clips",
            )
        return out

    def run_windows_device(self, windows_dev):
        """cuda f32 ``[n, 76, 32]`` -> cuda f32 ``[n, 96]``."""
        assert windows_dev.shape[1:] == (spec.EMB_WINDOW, spec.N_MELS)
        return self.run_clips_device(windows_dev, [0])[:, 0, :]

    def activation_device(self, mel_dev, layer: int):
        """Parity hook: activation after conv ``layer`` as cuda f32 NHWC ``[B, T_l, F_l, C_l]``."""
        import torch

        if not self.loaded:
            self.load()
        b, f, _ = mel_dev.shape
        t_l, f_l, c_l = spec.embedding_layer_shapes(f, spec.N_MELS)[layer]
        out = torch.empty((b, t_l, f_l, c_l), dtype=torch.float32, device=mel_dev.device)
        lib = _native.load()
        with torch.cuda.device(mel_dev.device):
            nbytes = lib.hb_embed_clips_workspace_bytes(b, f, self.mode)
            ws = self._get_workspace(nbytes, mel_dev.device)
            n = lib.hb_embed_activation(self._handle, self.mode, mel_dev.data_ptr(), b, f, layer, out.data_ptr(),
                                        out.numel(), ws.data_ptr(), ws.numel(), _native.stream_ptr(mel_dev.device))
            _native.check(n, "hb_embed_activation")
        assert n == out.numel(), (n, out.shape)
        return out

    # -- reference surface ------------------------------------------------------------------------
    def _run(self, **inputs: np.ndarray) -> List[np.ndarray]:
        import torch

        spectrograms = np.ascontiguousarray(inputs[self.input_name], dtype=np.float32)
        assert spectrograms.ndim == 4 and spectrograms.shape[1:] == (spec.EMB_WINDOW, spec.N_MELS, 1), spectrograms.shape
        dev = torch.from_numpy(spectrograms[..., 0]).to(self.device)
        out = self.run_windows_device(dev)
        return [out.cpu().numpy()[:, None, None, :]]

    def __call__(self, spectrograms: np.ndarray, **kwargs: Any) -> np.ndarray:  # type: ignore[override]
        out = super().__call__(input_1=spectrograms, **kwargs)[0]
        return out.reshape(out.shape[0], spec.EMB_DIM)  # never squeeze the batch axis (Appendix B)


class SpeechEmbeddings:
    """A class to compute embeddings from audio (reference embeddings.py:44-234)."""

    def __init__(self, device_id: Optional[int] = None, load: bool = False, precision: Optional[str] = None,
                 clip_chunk: int = DEFAULT_CLIP_CHUNK, seed: Optional[int] = None,
                 weights: Optional[Union[Dict[str, np.ndarray], str]] = None) -> None:
        self.spectrogram = MelSpectrogramModel(device_id=device_id, load=load)
        self.embeddings = SpeechEmbeddingModel(device_id=device_id, load=load, precision=precision, weights=weights)
        self.clip_chunk = clip_chunk
        self._rng = np.random.default_rng(seed)

    @property
    def device(self):
        return self.spectrogram.device

    # -- reference per-window methods -----------------------------------------------------------
    def audio_to_spectrograms(
        self,
        audio,
        batch_size: int = 128,
        mel_bins: int = 32,
        on_progress: Optional[Callable[[int, int], None]] = None,
    ) -> np.ndarray:
        """``[b, t]`` tensor (int16-range floats) -> np f32 ``[b, ceil(t/160 - 3), 32]`` (embeddings.py:56-84)."""
        import torch

        assert mel_bins == spec.N_MELS
        if isinstance(audio, np.ndarray):
            audio = torch.from_numpy(audio)
        b, t = audio.shape
        n_frames = spec.reference_frames(t)
        n_total = b * n_frames
        out = np.empty((b, n_frames, mel_bins), dtype=np.float32)
        for i in range(0, b, batch_size):
            chunk = audio[i:i + batch_size].detach().to(self.device, dtype=torch.float32).contiguous()
            mel = self.spectrogram.run_device(chunk)
            out[i:i + batch_size] = mel[:, :n_frames].cpu().numpy()
            if on_progress is not None:
                on_progress(min(i + batch_size, b), n_total)
        if on_progress is not None:
            on_progress(n_total, n_total)
        return out

    def spectrograms_to_embeddings(
        self,
        spectrograms: np.ndarray,
        batch_size: int = 128,
        embedding_dim: int = 96,
        window_size: int = 76,
        window_stride: int = 8,
        on_progress: Optional[Callable[[int, int], None]] = None,
    ) -> np.ndarray:
        """np f32 ``[b, t, 32]`` -> ``[b, (t-76)//8 + 1, 96]`` (embeddings.py:86-151)."""
        import torch

        assert embedding_dim == spec.EMB_DIM and window_size == spec.EMB_WINDOW
        b, t, m = spectrograms.shape
        assert t >= window_size, f"Time dimension {t} must be at least {window_size}"
        n_frames = (t - window_size) // window_stride + 1
        n_total = b * n_frames
        offsets = [j * window_stride for j in range(n_frames)]
        out = np.empty((b, n_frames, embedding_dim), dtype=np.float32)
        clip_batch = max(1, batch_size // max(1, n_frames))
        fully_conv = all(o % 4 == 0 for o in offsets)
        done = 0
        for i in range(0, b, clip_batch):
            mel = torch.from_numpy(np.ascontiguousarray(spectrograms[i:i + clip_batch], dtype=np.float32)).to(self.device)
            if fully_conv:
                emb = self.embeddings.run_clips_device(mel, offsets)
            else:
                win = torch.stack([mel[:, o:o + window_size] for o in offsets], dim=1).reshape(-1, window_size, m)
                emb = self.embeddings.run_windows_device(win.contiguous()).reshape(mel.shape[0], n_frames, embedding_dim)
            out[i:i + clip_batch] = emb.cpu().numpy()
            done += mel.shape[0] * n_frames
            if on_progress is not None:
                on_progress(done, n_total)
        if on_progress is not None:
            on_progress(n_total, n_total)
        return out

    # -- fused device path ---------------------------------------------------------------------------
    def embed_device(self, audio_dev, scale: float = spec.AUDIO_SCALE, slot_offsets=None, return_mel: bool = False):
        """cuda f32 ``[B, T]`` in [-1, 1] -> cuda f32 ``[B, n_slots, 96]`` (reference slot order)."""
        t = audio_dev.shape[1]
        if slot_offsets is None:
            slot_offsets = spec.embedding_frame_offsets(t)
        mel = self.spectrogram.run_device(audio_dev, scale=scale)
        emb = self.embeddings.run_clips_device(mel, slot_offsets)
        return (emb, mel) if return_mel else emb

    def __call__(
        self,
        audio: Any,
        spectrogram_batch_size: int = 32,
        mel_bins: int = 32,
        embedding_batch_size: int = 32,
        embedding_dim: int = 96,
        window_size: int = 76,
        window_stride: int = 8,
        audio_window_size: int = 17280,
        audio_window_stride: int = 1920,
        on_spectrogram_progress: Optional[Callable[[int, int], None]] = None,
        on_embedding_progress: Optional[Callable[[int, int], None]] = None,
        remove_nan: bool = True,
        return_spectrograms: bool = False,
    ) -> Union[np.ndarray, Tuple[np.ndarray, np.ndarray]]:
        """Compute embeddings from audio (embeddings.py:153-234)."""
        import torch

        audio_tensor, _ = audio_to_bct_tensor(audio, sample_rate=16000)
        audio_tensor = audio_tensor.to(torch.float32)
        if audio_tensor.shape[1] > 1:
            audio_tensor = audio_tensor.mean(dim=1, keepdim=True)
        audio_tensor = audio_tensor[:, 0, :]  # batch, time (the x32767 is applied inside the mel kernel)
        b, t = audio_tensor.shape

        starts = list(range(0, t - audio_window_size + 1, audio_window_stride))
        if not starts:
            raise ValueError("need at least one array to concatenate")  # what np.concatenate([]) raises upstream
        offsets = spec.embedding_frame_offsets(t, audio_window_size, audio_window_stride, window_size, window_stride)
        aligned = audio_window_stride % spec.HOP == 0 and all(o % 4 == 0 for o in offsets)

        if not aligned:
            # generic (unaligned) geometry: reference control flow, window by window, same kernels
            embeddings_list, spectrograms_list = [], []
            for i in starts:
                s = self.audio_to_spectrograms(audio_tensor[:, i:i + audio_window_size] * spec.AUDIO_SCALE,
                                               batch_size=spectrogram_batch_size, mel_bins=mel_bins,
                                               on_progress=on_spectrogram_progress)
                embeddings_list.append(self.spectrograms_to_embeddings(
                    s, batch_size=embedding_batch_size, embedding_dim=embedding_dim, window_size=window_size,
                    window_stride=window_stride, on_progress=on_embedding_progress))
                spectrograms_list.append(s)
            embeddings = np.concatenate(embeddings_list, axis=1)
            spectrograms = np.concatenate(spectrograms_list, axis=1) if return_spectrograms else None
        else:
            n_slots = len(offsets)
            embeddings = np.empty((b, n_slots, embedding_dim), dtype=np.float32)
            frames_per_window = spec.mel_frames(audio_window_size)
            frame_index = None
            if return_spectrograms:
                frame_index = torch.as_tensor(
                    [s // spec.HOP + f for s in starts for f in range(frames_per_window)], dtype=torch.long)
                spectrograms = np.empty((b, frame_index.numel(), mel_bins), dtype=np.float32)
            else:
                spectrograms = None
            n_total = b * n_slots
            for i in range(0, b, self.clip_chunk):
                chunk = audio_tensor[i:i + self.clip_chunk].to(self.device, non_blocking=True).contiguous()
                emb, mel = self.embed_device(chunk, slot_offsets=offsets, return_mel=True)
                embeddings[i:i + self.clip_chunk] = emb.cpu().numpy()
                if return_spectrograms:
                    spectrograms[i:i + self.clip_chunk] = mel[:, frame_index.to(mel.device)].cpu().numpy()
                done = min(i + self.clip_chunk, b)
                if on_spectrogram_progress is not None:
                    on_spectrogram_progress(done, b)
                if on_embedding_progress is not None:
                    on_embedding_progress(done * n_slots, n_total)

        if remove_nan:
            nan_rows = np.nonzero(np.isnan(embeddings).any(axis=(1, 2)))[0]
            if nan_rows.size:
                logger.warning(f"Replacing {nan_rows.size} NaN embeddings with random embeddings.")
                keep = np.setdiff1d(np.arange(len(embeddings)), nan_rows)
                if keep.size == 0:
                    logger.warning("All embeddings are NaN, returning zero embeddings.")
                    return np.zeros(embeddings.shape, dtype=np.float32)
                for i in nan_rows:
                    embeddings[i] = embeddings[self._rng.choice(keep)]

        if return_spectrograms:
            tt = spectrograms.shape[1]
            truncated_t = tt - ((tt - window_size) % window_stride)
            return embeddings, spectrograms[:, :truncated_t]
        return embeddings


GLOBAL_EMBEDDINGS: Dict[Optional[int], SpeechEmbeddings] = {}


def get_speech_embeddings(device_id: Optional[int] = None, weights: Optional[Union[Dict[str, np.ndarray], str]] = None) -> SpeechEmbeddings:
    """
    Get a SpeechEmbeddings instance for a given device_id (embeddings.py:236-243).  ``weights`` (dict, ``.npz`` path or
    ``"random-init"``) applies when the instance is first created.
    """
    if device_id not in GLOBAL_EMBEDDINGS:
        GLOBAL_EMBEDDINGS[device_id] = SpeechEmbeddings(device_id=device_id, weights=weights)
    return GLOBAL_EMBEDDINGS[device_id]
