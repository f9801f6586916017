"""
Oracle: speech-embedding CNN (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows the reference call site ``SpeechEmbeddingModel.__call__``
(/root/reference/src/python/heybuddy/embeddings.py:32-42): input f32
``[n, 76, 32, 1]`` named ``input_1``, output ``conv2d_19`` ``[n, 1, 1, 96]``
squeezed to ``[n, 96]``.  The model file (``speech-embedding.onnx``, sha256
70d16429...5c1f, embeddings.py:29-30) is not on disk: I/O is pinned, the interior
(``heybuddy_b200.spec.EMBEDDING_LAYERS``) is the proposed table of SURVEY.md A.6
-> **parity unpinned**.  Weights are random-init from ``spec.init_embedding_weights``.

Plain PyTorch CPU convolutions; ``dtype=torch.float64`` gives the exact answer,
``float32`` what a straight fp32 implementation sees.  ``operand_rounding``
emulates a tensor-core operand format (fp32 accumulate) so tests can state a
tolerance for the tf32 / f16 / bf16 modes of the CUDA path.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from heybuddy_b200 import spec


def _round_operand(x: torch.Tensor, mode: Optional[str]) -> torch.Tensor:
    if mode is None:
        return x
    if mode == "bf16":
        return x.to(torch.bfloat16).to(x.dtype)
    if mode == "f16":
        return x.to(torch.float16).to(x.dtype)
    if mode == "tf32":
        # round-to-nearest-even to 10 explicit mantissa bits
        xi = x.to(torch.float32).contiguous().view(torch.int32)
        lsb = (xi >> 13) & 1
        xi = (xi + 0x0FFF + lsb) & ~0x1FFF
        return xi.view(torch.float32).to(x.dtype)
    raise ValueError(mode)


def embed_strip(
    mel: np.ndarray,
    weights: Dict[str, np.ndarray],
    dtype: torch.dtype = torch.float32,
    operand_rounding: Optional[str] = None,
) -> np.ndarray:
    """
    Fully-convolutional evaluation: ``mel [n, T, 32]`` (T >= 76) -> ``[n, T_out, 96]``
    where T_out = number of stride-8 windows starting at frame 0.  For T == 76 this
    is exactly the reference model (``[n, 1, 96]``).
    """
    x = torch.from_numpy(np.ascontiguousarray(mel)).to(dtype)[:, None, :, :]  # NCHW: [n,1,T,F]
    for li, (name, kh, kw, cin, cout, pad, act, pool) in enumerate(spec.EMBEDDING_LAYERS):
        w = torch.from_numpy(weights[f"{name}.weight"]).to(dtype).permute(3, 2, 0, 1).contiguous()  # OIHW
        b = torch.from_numpy(weights[f"{name}.bias"]).to(dtype)
        padding = (0, kw // 2) if pad == "same" else (0, 0)
        if li == 0:
            xin, win = x, w  # the first (Cin=1) conv runs on fp32 CUDA cores in every mode
        else:
            xin, win = _round_operand(x, operand_rounding), _round_operand(w, operand_rounding)
        x = F.conv2d(xin, win, b, padding=padding)
        if act:
            x = F.leaky_relu(x, spec.LEAKY_SLOPE)
        if pool is not None:
            x = F.max_pool2d(x, kernel_size=pool, stride=pool)
    # [n, 96, T_out, 1] -> [n, T_out, 96]
    return x[:, :, :, 0].permute(0, 2, 1).contiguous().to(torch.float32).numpy()


def speech_embedding_model(
    spectrograms: np.ndarray,
    weights: Dict[str, np.ndarray],
    dtype: torch.dtype = torch.float32,
    operand_rounding: Optional[str] = None,
) -> np.ndarray:
    """``[n, 76, 32, 1] -> [n, 96]`` (the reference's ring-1 callable, un-squeezed on axis 0)."""
    spectrograms = np.asarray(spectrograms, dtype=np.float32)
    assert spectrograms.ndim == 4 and spectrograms.shape[1:] == (spec.EMB_WINDOW, spec.N_MELS, 1), spectrograms.shape
    return embed_strip(spectrograms[..., 0], weights, dtype=dtype, operand_rounding=operand_rounding)[:, 0, :]
