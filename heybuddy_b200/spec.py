"""
Single source of truth for every constant on the featurization hot path.

Both the CPU oracle (``oracle/``, test infrastructure) and the CUDA path
(``heybuddy_b200/csrc``) read the numbers below; nothing else in the tree may
hard-code them.  Each constant cites the reference file:line that pins it
(paths relative to ``/root/reference/src/python/heybuddy`` unless stated).

Pinned by reference source
--------------------------
* ``SAMPLE_RATE``, ``TARGET_LENGTH_S`` -> ``CLIP_SAMPLES`` = int(1.44*16000) = 23040
  (dataset/augmented.py:31,123-128)
* audio window 17280, stride 1920 (embeddings.py:162-163)
* mel hop 160 and the frame-count formula ceil(t/160-3) (embeddings.py:66-67)
* 32 mel bins, 76-frame embedding window, stride 8, 96-d (embeddings.py:88-92)
* mel post-scale ``x/10 + 2`` (spectrogram.py:32)
* ``audio *= 32767`` before the mel (embeddings.py:182)
* classifier architecture (wakeword.py:171-348, modules/multi_layer_perceptron.py:76-124)

Restated from third-party behaviour (not under /root/reference; "parity unpinned",
see SURVEY.md 8c / Appendix A)
* mel = torchaudio MelSpectrogram(n_fft=512, win_length=400, hop_length=160,
  center=False, n_mels=32, f_min=60, f_max=3800, HTK, norm=None, power=2) -> 10*log10(max(P,1e-10))
* embedding CNN interior (layer table below): I/O pinned, interior proposed.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# Audio / windowing constants
# ----------------------------------------------------------------------------
SAMPLE_RATE = 16000
TARGET_LENGTH_S = 1.44
CLIP_SAMPLES = int(TARGET_LENGTH_S * SAMPLE_RATE)  # 23040
AUDIO_WINDOW = 17280
AUDIO_STRIDE = 1920
AUDIO_SCALE = 32767.0  # embeddings.py:182

N_FFT = 512
WIN_LENGTH = 400
HOP = 160
N_FREQ = N_FFT // 2 + 1  # 257
N_MELS = 32
MEL_FMIN = 60.0
MEL_FMAX = 3800.0
MEL_FLOOR = 1e-10  # amin of AmplitudeToDB(power)
MEL_POST_DIV = 10.0  # spectrogram.py:32
MEL_POST_ADD = 2.0

EMB_WINDOW = 76
EMB_STRIDE = 8
EMB_DIM = 96
LEAKY_SLOPE = 0.2

# frame f of audio-window w == global frame (AUDIO_STRIDE/HOP)*w + f
FRAMES_PER_AUDIO_STRIDE = AUDIO_STRIDE // HOP  # 12


def mel_frames(num_samples: int) -> int:
    """Frames the mel model returns for ``num_samples`` (no centring/padding)."""
    if num_samples < N_FFT:
        return 0
    return 1 + (num_samples - N_FFT) // HOP


def reference_frames(num_samples: int) -> int:
    """The reference's own frame-count formula, embeddings.py:67."""
    return int(np.ceil(num_samples / 160 - 3))


def audio_window_starts(num_samples: int, window: int = AUDIO_WINDOW, stride: int = AUDIO_STRIDE) -> List[int]:
    """embeddings.py:190 ``range(0, T - window + 1, stride)``."""
    return list(range(0, num_samples - window + 1, stride))


def embedding_frame_offsets(
    num_samples: int,
    audio_window: int = AUDIO_WINDOW,
    audio_stride: int = AUDIO_STRIDE,
    window: int = EMB_WINDOW,
    stride: int = EMB_STRIDE,
) -> List[int]:
    """
    Global mel-frame offset of every embedding slot, in the reference's output
    order (embeddings.py:190-209 + :131-137).  For 23040 samples this is
    [0,8,16,24, 12,20,28,36, 24,32,40,48, 36,44,52,60].
    """
    offs: List[int] = []
    frames_per_window = mel_frames(audio_window)
    for start in audio_window_starts(num_samples, audio_window, audio_stride):
        assert start % HOP == 0, "audio stride must be a multiple of the mel hop"
        base = start // HOP
        n = (frames_per_window - window) // stride + 1
        offs.extend(base + stride * j for j in range(n))
    return offs


# ----------------------------------------------------------------------------
# Mel front-end tables (fp64 math, rounded once to fp32)
# ----------------------------------------------------------------------------
def hann_window_padded() -> np.ndarray:
    """
    Periodic Hann(400) zero-padded symmetrically to 512 (torch.stft convention
    for win_length < n_fft): 56 zeros each side.
    """
    n = np.arange(WIN_LENGTH, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / WIN_LENGTH)
    out = np.zeros(N_FFT, dtype=np.float64)
    left = (N_FFT - WIN_LENGTH) // 2
    out[left:left + WIN_LENGTH] = w
    return out.astype(np.float32)


def _hz_to_mel_htk(f: np.ndarray) -> np.ndarray:
    return 2595.0 * np.log10(1.0 + f / 700.0)


def _mel_to_hz_htk(m: np.ndarray) -> np.ndarray:
    return 700.0 * (10.0 ** (m / 2595.0) - 1.0)


def mel_filterbank() -> np.ndarray:
    """
    HTK triangular filterbank, norm=None, shape [257, 32] float32; restates
    ``torchaudio.functional.melscale_fbanks(257, 60, 3800, 32, 16000)``.
    """
    all_freqs = np.linspace(0.0, SAMPLE_RATE // 2, N_FREQ, dtype=np.float64)
    m_min = _hz_to_mel_htk(np.asarray(MEL_FMIN, dtype=np.float64))
    m_max = _hz_to_mel_htk(np.asarray(MEL_FMAX, dtype=np.float64))
    m_pts = np.linspace(m_min, m_max, N_MELS + 2, dtype=np.float64)
    f_pts = _mel_to_hz_htk(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]  # [257, 34]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    return fb.astype(np.float32)


def mel_band_limits() -> Tuple[int, int]:
    """[lo, hi) range of FFT bins with a non-zero filterbank row."""
    fb = mel_filterbank()
    nz = np.nonzero(fb.sum(axis=1) > 0)[0]
    return int(nz[0]), int(nz[-1]) + 1


# ----------------------------------------------------------------------------
# Speech-embedding CNN layer table (I/O pinned by
# embeddings.py:32-42 / src/ts/src/models/speech-embedding.ts:137-152, interior proposed)
# ----------------------------------------------------------------------------
# (name, kh(time), kw(freq), cin, cout, padding, leaky_relu, pool_after (time, freq) | None)
EMBEDDING_LAYERS: List[Tuple[str, int, int, int, int, str, bool, Optional[Tuple[int, int]]]] = [
    ("conv2d",    1, 3,  1, 24, "same",  True, None),
    ("conv2d_1",  3, 1, 24, 24, "valid", True, None),
    ("conv2d_2",  1, 3, 24, 24, "same",  True, None),
    ("conv2d_3",  3, 1, 24, 24, "valid", True, (2, 2)),
    ("conv2d_4",  1, 3, 24, 48, "same",  True, None),
    ("conv2d_5",  3, 1, 48, 48, "valid", True, None),
    ("conv2d_6",  1, 3, 48, 48, "same",  True, None),
    ("conv2d_7",  3, 1, 48, 48, "valid", True, (1, 2)),
    ("conv2d_8",  1, 3, 48, 72, "same",  True, None),
    ("conv2d_9",  3, 1, 72, 72, "valid", True, None),
    ("conv2d_10", 1, 3, 72, 72, "same",  True, None),
    ("conv2d_11", 3, 1, 72, 72, "valid", True, (2, 2)),
    ("conv2d_12", 1, 3, 72, 96, "same",  True, None),
    ("conv2d_13", 3, 1, 96, 96, "valid", True, None),
    ("conv2d_14", 1, 3, 96, 96, "same",  True, None),
    ("conv2d_15", 3, 1, 96, 96, "valid", True, (2, 2)),
    ("conv2d_16", 1, 2, 96, 96, "valid", True, None),
    ("conv2d_17", 3, 1, 96, 96, "valid", True, None),
    ("conv2d_18", 1, 1, 96, 96, "valid", True, None),
    ("conv2d_19", 3, 1, 96, 96, "valid", False, None),
]
EMBEDDING_WEIGHT_SEED = 3001


def embedding_layer_shapes(t_in: int = EMB_WINDOW, f_in: int = N_MELS) -> List[Tuple[int, int, int]]:
    """(time, freq, channels) after every layer (after its pool, if any)."""
    shapes = []
    t, f = t_in, f_in
    for (_, kh, kw, _cin, cout, pad, _act, pool) in EMBEDDING_LAYERS:
        if pad == "valid":
            t, f = t - kh + 1, f - kw + 1
        if pool is not None:
            t, f = t // pool[0], f // pool[1]
        shapes.append((t, f, cout))
    return shapes


def embedding_macs_per_window(t_in: int = EMB_WINDOW, f_in: int = N_MELS) -> int:
    """Multiply-accumulates of one forward pass over a [t_in, f_in] strip."""
    macs = 0
    t, f = t_in, f_in
    for (_, kh, kw, cin, cout, pad, _act, pool) in EMBEDDING_LAYERS:
        if pad == "valid":
            t, f = t - kh + 1, f - kw + 1
        macs += t * f * cout * kh * kw * cin
        if pool is not None:
            t, f = t // pool[0], f // pool[1]
    return macs


def embedding_macs_per_clip(frames: int = 141, f_in: int = N_MELS) -> int:
    """
    MACs of the fully-convolutional evaluation of one clip (``hb_embed_clips``): conv2d..conv2d_15 once over
    the whole ``frames``-frame strip, then the last 2x2 pool and block 5 once per pool phase (window offsets
    = 0 and 4 mod 8).  This is the algorithmic work behind ``roofline.achieved`` in bench.py.
    """
    macs = 0
    t, f = frames, f_in
    for li, (_, kh, kw, cin, cout, pad, _act, pool) in enumerate(EMBEDDING_LAYERS[:16]):
        if pad == "valid":
            t, f = t - kh + 1, f - kw + 1
        macs += t * f * cout * kh * kw * cin
        if pool is not None and li != 15:
            t, f = t // pool[0], f // pool[1]
    for phase in (0, 1):
        tp, fp = (t - phase) // 2, f // 2
        for (_, kh, kw, cin, cout, pad, _act, pool) in EMBEDDING_LAYERS[16:]:
            tp, fp = tp - kh + 1, fp - kw + 1
            macs += tp * fp * cout * kh * kw * cin
    return macs


def embedding_num_params() -> int:
    return sum(kh * kw * cin * cout + cout for (_, kh, kw, cin, cout, *_r) in EMBEDDING_LAYERS)


def init_embedding_weights(seed: int = EMBEDDING_WEIGHT_SEED) -> Dict[str, np.ndarray]:
    """
    Random-init weights of the table above: ``randn * fan_in**-0.5`` kernels in
    HWIO layout ``[kh, kw, cin, cout]`` and ``0.1*randn`` biases, from a seeded
    numpy Generator so the oracle and the CUDA path load identical bits.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    out: Dict[str, np.ndarray] = {}
    for (name, kh, kw, cin, cout, *_r) in EMBEDDING_LAYERS:
        fan_in = kh * kw * cin
        out[f"{name}.weight"] = (rng.standard_normal((kh, kw, cin, cout)) * fan_in ** -0.5).astype(np.float32)
        out[f"{name}.bias"] = (0.1 * rng.standard_normal((cout,))).astype(np.float32)
    return out


# ----------------------------------------------------------------------------
# Wake-word classifier (wakeword.py:171-348; fully pinned)
# ----------------------------------------------------------------------------
CLS_FRAMES = 16
CLS_IN = CLS_FRAMES * EMB_DIM  # 1536
CLS_DIM = 96
CLS_LAYERS = 2
LN_EPS = 1e-5


def normalized_dim(dim: int, multiple_of: int = 8, down_ratio: float = 2 / 3) -> int:
    """util/modeling_util.py:42-72."""
    d = int(dim * down_ratio)
    if d % multiple_of == 0:
        return d
    return d + multiple_of - (d % multiple_of)


CLS_HIDDEN = normalized_dim(CLS_DIM)  # 64


def classifier_param_shapes(layer_dim: int = CLS_DIM, num_layers: int = CLS_LAYERS) -> List[Tuple[str, Tuple[int, ...]]]:
    """
    state_dict keys and shapes in ``nn.Module`` registration order
    (== ONNX initializer names of src/ts/models/*.onnx).
    """
    h = normalized_dim(layer_dim)
    shapes: List[Tuple[str, Tuple[int, ...]]] = [
        ("norm_in.weight", (CLS_IN,)), ("norm_in.bias", (CLS_IN,)),
        ("mlp_in.hidden.weight", (h, CLS_IN)), ("mlp_in.hidden.bias", (h,)),
        ("mlp_in.output.weight", (layer_dim, h)), ("mlp_in.output.bias", (layer_dim,)),
        ("mlp_in.gate.weight", (h, CLS_IN)), ("mlp_in.gate.bias", (h,)),
    ]
    for l in range(num_layers):
        shapes += [
            (f"layers.{l}.0.weight", (layer_dim,)), (f"layers.{l}.0.bias", (layer_dim,)),
            (f"layers.{l}.1.hidden.weight", (h, layer_dim)), (f"layers.{l}.1.hidden.bias", (h,)),
            (f"layers.{l}.1.output.weight", (layer_dim, h)), (f"layers.{l}.1.output.bias", (layer_dim,)),
            (f"layers.{l}.1.gate.weight", (h, layer_dim)), (f"layers.{l}.1.gate.bias", (h,)),
        ]
    shapes += [
        ("norm_out.weight", (layer_dim,)), ("norm_out.bias", (layer_dim,)),
        ("mlp_out.hidden.weight", (h, layer_dim)), ("mlp_out.hidden.bias", (h,)),
        ("mlp_out.output.weight", (1, h)), ("mlp_out.output.bias", (1,)),
        ("mlp_out.gate.weight", (h, layer_dim)), ("mlp_out.gate.bias", (h,)),
    ]
    return shapes


def classifier_num_params() -> int:
    return sum(int(np.prod(s)) for _, s in classifier_param_shapes())


def init_classifier_weights(seed: int = 5002) -> Dict[str, np.ndarray]:
    """nn.Linear-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) init; LayerNorm = (1, 0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out: Dict[str, np.ndarray] = {}
    for name, shape in classifier_param_shapes():
        if name.startswith("norm") or name.endswith(".0.weight") or name.endswith(".0.bias"):
            out[name] = (np.ones(shape) if name.endswith("weight") else np.zeros(shape)).astype(np.float32)
        else:
            fan_in = shape[-1] if len(shape) == 2 else None
            if fan_in is None:
                # bias: fan_in of the matching weight
                fan_in = out[name.replace(".bias", ".weight")].shape[1]
            bound = 1.0 / math.sqrt(fan_in)
            out[name] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    return out


# ----------------------------------------------------------------------------
# Augmentation draw table: one row per
# augmentation batch, consumed identically by the oracle and the CUDA kernel.
# ----------------------------------------------------------------------------
COLORED_BASE_SAMPLES = SAMPLE_RATE  # 1 s pattern tiled to the clip length
GAIN_MIN_DB = -18.0  # torch_audiomentations.Gain defaults (augmented.py:116-120 passes only p)
GAIN_MAX_DB = 6.0
REVERB_EPS = 1e-14  # speechbrain reverberate rescale epsilon
