"""
ctypes binding of the C-ABI in ``include/heybuddy_b200.h``.

There is no CPU fallback: if the shared library is missing or CUDA is unavailable the
product path raises.  ``torch`` is used only for device memory, streams and pinned
buffers; every kernel on the path comes from ``libheybuddy_b200.so``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import numpy as np

from heybuddy_b200 import spec

HB_OK = 0
HB_EMBED_FP32 = 0
HB_EMBED_F16 = 1
EMBED_MODES = {"fp32": HB_EMBED_FP32, "f16": HB_EMBED_F16}

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libheybuddy_b200.so")
_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()
_tables_ready = set()


class NativeError(RuntimeError):
    """Raised when a C-ABI call returns a negative status."""


class hb_clip_aug(ctypes.Structure):
    _fields_ = [
        ("noise_offset", ctypes.c_int64),
        ("colored_index", ctypes.c_int32),
        ("rir_index", ctypes.c_int32),
        ("colored_snr_db", ctypes.c_float),
        ("gain", ctypes.c_float),
        ("noise_snr_db", ctypes.c_float),
        ("reserved", ctypes.c_int32),
    ]


# numpy view of the same 32-byte record
CLIP_AUG_DTYPE = np.dtype([
    ("noise_offset", "<i8"), ("colored_index", "<i4"), ("rir_index", "<i4"),
    ("colored_snr_db", "<f4"), ("gain", "<f4"), ("noise_snr_db", "<f4"), ("reserved", "<i4"),
])
assert CLIP_AUG_DTYPE.itemsize == ctypes.sizeof(hb_clip_aug) == 32


def lib_path() -> str:
    return _LIB_PATH


def _declare(lib: ctypes.CDLL) -> None:
    c_int, c_i64, c_f, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
    sig = {
        "hb_abi_version": (c_int, []),
        "hb_last_error": (ctypes.c_char_p, []),
        "hb_launch_count": (c_i64, []),
        "hb_init_tables": (c_int, [c_vp, c_vp]),
        "hb_mel_frames": (c_int, [c_int]),
        "hb_mel_f32": (c_int, [c_vp, c_i64, c_f, c_vp, c_int, c_int, c_vp]),
        "hb_embed_create": (c_int, [ctypes.POINTER(c_vp), c_vp, c_i64]),
        "hb_embed_destroy": (c_int, [c_vp]),
        "hb_embed_num_params": (c_i64, []),
        "hb_embed_windows_workspace_bytes": (c_i64, [c_int, c_int]),
        "hb_embed_windows": (c_int, [c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
        "hb_embed_clips_workspace_bytes": (c_i64, [c_int, c_int, c_int]),
        "hb_embed_clips": (c_int, [c_vp, c_int, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_i64, c_vp]),
        "hb_embed_activation": (c_i64, [c_vp, c_int, c_vp, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_vp]),
        "hb_rir_spectrum": (c_int, [c_vp, c_vp, c_int, c_int, c_vp]),
        "hb_check_kernels": (c_int, []),
        "hb_linear_tf32x3": (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp]),
        "hb_colored_bases": (c_int, [ctypes.c_uint64, c_vp, c_vp, c_int, c_vp, c_vp]),
        "hb_augment_clips_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
        "hb_k9_eq_f32": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
        "hb_k9_tanh_f32": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
        "hb_k9_bandstop_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
        "hb_pitch_plan_create": (c_int, [ctypes.POINTER(c_vp), c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp]),
        "hb_pitch_plan_destroy": (c_int, [c_vp]),
        "hb_k9_pitch_workspace_bytes": (c_i64, [c_vp, c_int]),
        "hb_k9_pitch_f32": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
        "hb_fix_length_i16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
        "hb_featurize_workspace_bytes": (c_i64, [c_int, c_int, c_int]),
        "hb_featurize_i16": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_int, c_vp, c_i64, c_vp]),
        "hb_augment_clips_i16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
        "hb_augment_mel_i16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f, c_vp, c_int, c_int, c_vp]),
        "hb_mlp_num_params": (c_i64, []),
        "hb_mlp_create": (c_int, [ctypes.POINTER(c_vp), c_vp, c_i64]),
        "hb_mlp_destroy": (c_int, [c_vp]),
        "hb_mlp_get_params": (c_int, [c_vp, c_vp, c_i64]),
        "hb_mlp_set_params": (c_int, [c_vp, c_vp, c_i64]),
        "hb_mlp_workspace_bytes": (c_i64, [c_int, c_int]),
        "hb_mlp_forward": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
        "hb_mlp_train_step": (c_int, [c_vp, c_vp, c_vp, c_int, c_f, c_f, c_f, c_int, c_vp, c_vp, c_vp, c_i64, c_vp]),
        "hb_mlp_select": (c_int, [c_vp, c_vp, c_vp, c_int, c_f, c_vp, c_vp, c_vp, c_i64, c_vp]),
        "hb_mlp_backward": (c_int, [c_vp, c_vp, c_vp, c_int, c_f, c_f, c_vp, c_int, c_vp, c_vp, c_vp, c_i64, c_vp]),
        "hb_mlp_grads_copy": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp]),
        "hb_mlp_local_step": (c_int, [c_vp, c_vp, c_vp, c_int, c_f, c_f, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
        "hb_mlp_apply_exchange": (c_int, [c_vp, c_vp, c_f, c_int, c_vp, c_vp]),
        "hb_mlp_adam": (c_int, [c_vp, c_f, c_vp, c_vp]),
        "hb_mlp_set_loss_scale": (c_int, [c_vp, c_f]),
        "hb_mlp_get_adam": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64]),
        "hb_mlp_set_adam": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64]),
        "hb_mlp_dropout": (c_int, [c_vp, c_vp, c_i64, c_f, ctypes.c_uint64, ctypes.c_uint64, c_vp]),
        "hb_mlp_get_grads": (c_int, [c_vp, c_vp, c_i64]),
        "hb_mlp_multi_workspace_bytes": (c_i64, [c_int, c_int]),
        "hb_mlp_forward_multi": (c_int, [c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args


DECLARED_SYMBOLS = None  # filled on load (names bound above)


def load() -> ctypes.CDLL:
    """Load the shared library (no CUDA call is made).  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise NativeError(
                    f"{_LIB_PATH} is missing: run `python -m heybuddy_b200.build` (or __graft_entry__.build()). "
                    "heybuddy_b200 has no CPU fallback."
                )
            lib = ctypes.CDLL(_LIB_PATH)
            _declare(lib)
            if lib.hb_abi_version() != 1:
                raise NativeError(f"ABI mismatch: library reports {lib.hb_abi_version()}, binding expects 1")
            _lib = lib
        return _lib


def check(status: int, what: str = "") -> int:
    if status < 0:
        msg = load().hb_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'heybuddy_b200'} failed ({status}): {msg}")
    return status


def require_cuda(device_id: Optional[int] = None):
    """Returns the torch device to run on; raises when there is no GPU (no CPU path exists)."""
    import torch

    if not torch.cuda.is_available():
        raise NativeError("heybuddy_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device_id is None else int(device_id)
    return torch.device(f"cuda:{idx}")


def ensure_tables(device) -> None:
    """Uploads the Hann window / mel filterbank constant tables once per device."""
    import torch

    lib = load()
    key = device.index
    if key in _tables_ready:
        return
    with _lock:
        if key in _tables_ready:
            return
        hann = np.ascontiguousarray(spec.hann_window_padded(), dtype=np.float32)
        fb = np.ascontiguousarray(spec.mel_filterbank(), dtype=np.float32)
        with torch.cuda.device(device):
            check(lib.hb_init_tables(hann.ctypes.data, fb.ctypes.data), "hb_init_tables")
        _tables_ready.add(key)


def stream_ptr(device=None) -> int:
    import torch

    return int(torch.cuda.current_stream(device).cuda_stream)
