"""
``MelSpectrogramModel`` -- B200 replacement of reference ``heybuddy/spectrogram.py``.

Same callable contract (spectrogram.py:23-32): ``audio`` is a float32 numpy array ``[T]`` or
``[B, T]`` (already in int16 range, embeddings.py:182), the return value is
``squeeze(model(audio)) / 10 + 2`` with shape ``[B, F, 32]`` (``[F, 32]`` for one clip), where
the model's own output is ``[B, 1, F, 32]`` under the input name ``input``.

The model is the fused STFT-power -> mel -> log kernel ``hb_mel_f32``
(``csrc/mel.cu``); the ``/10 + 2`` is folded into the kernel.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from heybuddy_b200 import _native, spec
from heybuddy_b200.util import PretrainedNativeModel

__all__ = ["MelSpectrogramModel", "get_mel_spectrogram_model"]


class MelSpectrogramModel(PretrainedNativeModel):
    """Compute the (scaled) log-mel spectrogram of an audio signal on the GPU."""

    input_name = "input"

    def _load(self) -> None:
        _native.ensure_tables(self.device)

    # -- device-level entry (used by the fused pipeline; no host copies) --------------------
    def run_device(self, audio_dev, scale: float = 1.0, out=None):
        """``audio_dev`` cuda f32 ``[B, T]`` (row-contiguous) -> cuda f32 ``[B, F, 32]`` (already ``/10 + 2``)."""
        import torch

        if not self.loaded:
            self.load()
        assert audio_dev.is_cuda and audio_dev.dtype == torch.float32 and audio_dev.dim() == 2
        assert audio_dev.stride(1) == 1
        b, t = audio_dev.shape
        f = spec.mel_frames(t)
        if out is None:
            out = torch.empty((b, f, spec.N_MELS), dtype=torch.float32, device=audio_dev.device)
        if b == 0 or f == 0:
            return out
        lib = _native.load()
        with torch.cuda.device(audio_dev.device):
            _native.check(
                lib.hb_mel_f32(audio_dev.data_ptr(), audio_dev.stride(0) if b > 1 else t, float(scale), out.data_ptr(), b, t,
                               _native.stream_ptr(audio_dev.device)),
                "hb_mel_f32",
            )
        return out

    def _run(self, **inputs: np.ndarray) -> List[np.ndarray]:
        """ORT-style: named input ``input`` f32 ``[B, T]`` -> ``[out]`` with out ``[B, 1, F, 32]`` in dB."""
        import torch

        audio = inputs[self.input_name]
        audio_dev = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).to(self.device)
        mel = self.run_device(audio_dev)
        # undo the folded post-scale so the raw "model output" is dB like the ONNX graph's
        db = (mel - spec.MEL_POST_ADD) * spec.MEL_POST_DIV
        return [db[:, None, :, :].cpu().numpy()]

    def __call__(self, audio: np.ndarray, **kwargs: Any) -> np.ndarray:  # type: ignore[override]
        assert isinstance(audio, np.ndarray)
        if audio.ndim == 1:
            audio = audio[np.newaxis, :]
        assert audio.ndim == 2, f"Audio must be a 1D or 2D array, got {audio.ndim}D"
        import torch

        if not self.loaded:
            self.load()
        audio_dev = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).to(self.device)
        mel = self.run_device(audio_dev)
        return np.squeeze(mel.cpu().numpy())


GLOBAL_MEL_MODELS: Dict[Optional[int], MelSpectrogramModel] = {}


def get_mel_spectrogram_model(device_id: Optional[int] = None) -> MelSpectrogramModel:
    if device_id not in GLOBAL_MEL_MODELS:
        GLOBAL_MEL_MODELS[device_id] = MelSpectrogramModel(device_id=device_id, load=True)
    return GLOBAL_MEL_MODELS[device_id]
