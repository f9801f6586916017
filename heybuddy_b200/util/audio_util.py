"""
``audio_to_bct_tensor`` for in-memory audio (reference util/audio_util.py:73-145).

Same conventions: a list is a batch whose items are truncated to the shortest; 1-D = one
mono clip; 2-D = (channels, time); 3-D = (batch, channels, time); int16 -> /32768,
int8 -> (x-128)/128.  Paths / encoded bytes (torchaudio.load) and resampling are out of
scope for the hot path and raise.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import numpy as np

__all__ = ["audio_to_bct_tensor"]


def audio_to_bct_tensor(input_data: Any, sample_rate: Optional[int] = None, target_sample_rate: Optional[int] = None):
    import torch

    if isinstance(input_data, (list, tuple)):
        recursed: List[Tuple[Any, Optional[int]]] = [audio_to_bct_tensor(d, sample_rate) for d in input_data]
        min_frames = min(d.shape[-1] for d, _ in recursed)
        rates = [sr for _, sr in recursed if sr is not None]
        if rates and sample_rate is None:
            sample_rate = rates[0]
        return torch.cat([d[..., :min_frames] for d, _ in recursed], dim=0), sample_rate

    if isinstance(input_data, np.ndarray):
        waveform = torch.from_numpy(input_data)
    elif isinstance(input_data, torch.Tensor):
        waveform = input_data
    else:
        raise ValueError(
            f"Unsupported input type {type(input_data)}: the B200 hot path takes in-memory waveforms "
            "(numpy arrays, torch tensors or lists of them)"
        )
    if sample_rate is None:
        raise ValueError("No sample rate provided. Please provide a sample rate.")
    if waveform.dtype is torch.int16:
        waveform = waveform.float() / 32768.0
    elif waveform.dtype is torch.int8:
        waveform = (waveform.float() - 128) / 128.0
    if target_sample_rate is not None and sample_rate != target_sample_rate:
        raise ValueError("resampling is outside the B200 hot path; provide 16 kHz audio")
    if waveform.dim() == 1:
        waveform = waveform.unsqueeze(0)
    if waveform.dim() == 2:
        waveform = waveform.unsqueeze(0)
    return waveform, sample_rate
