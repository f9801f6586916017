"""
Streaming sliding-window inference for many wake-word models (BASELINE config 5).

Semantics = the browser runtime (``src/ts/src/hey-buddy.ts:382-469``): every 1920 new samples the last
17280 samples go through mel -> 4 embeddings (window-relative frame offsets 0, 8, 16, 24); a FIFO of the last 4
results forms the ``[16, 96]`` classifier input, and every wake-word model runs on the SAME buffer.  (The Python
reference's ``WakeWordModelThread`` re-featurizes the audio once per model, ``util/model_util.py:62-93``.)

Here the stream is featurized ONCE, offline-batched: frame f of step s is global mel frame 12 s + f, so a strip of
K consecutive steps is one fully-convolutional ``hb_embed_clips`` call over ``12 (K-1) + 105`` frames with
slot offsets ``12 k + 8 j``; the FIFO is a gather; ``hb_mlp_forward_multi`` evaluates all models on the buffer.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from heybuddy_b200 import spec
from heybuddy_b200.embeddings import SpeechEmbeddings
from heybuddy_b200.wakeword import MultiWakeWordModel, WakeWordMLPModel

__all__ = ["stream_step_embeddings", "stream_predict", "num_stream_steps", "WakeWordStreamService"]

STEPS_PER_STRIP = 32


def num_stream_steps(num_samples: int) -> int:
    """Steps at which a full 17280-sample window is available: ``len(range(0, n - 17280 + 1, 1920))``."""
    return len(spec.audio_window_starts(num_samples))


def stream_step_embeddings(speech: SpeechEmbeddings, audio, strips_per_call: int = 512):
    """
    ``audio``: 1-D float tensor / array in [-1, 1] (a long 16 kHz stream) -> cuda f32 ``[steps, 4, 96]``: the 4
    embeddings the runtime computes at every 1920-sample step.
    """
    import torch

    if isinstance(audio, np.ndarray):
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    audio = audio.reshape(-1).to(torch.float32)
    steps = num_stream_steps(audio.numel())
    if steps <= 0:
        raise ValueError("the stream is shorter than one 17280-sample window")
    dev = speech.device
    k = STEPS_PER_STRIP
    n_strips = (steps + k - 1) // k
    strip_samples = spec.AUDIO_WINDOW + spec.AUDIO_STRIDE * (k - 1)
    offsets = [spec.FRAMES_PER_AUDIO_STRIDE * s + spec.EMB_STRIDE * j for s in range(k) for j in range(4)]
    out = torch.empty((n_strips * k, 4, spec.EMB_DIM), dtype=torch.float32, device=dev)
    # pad the tail so the last (partial) strip is addressable; its extra steps are dropped below
    need = (n_strips - 1) * k * spec.AUDIO_STRIDE + strip_samples
    if need > audio.numel():
        audio = torch.cat([audio, torch.zeros(need - audio.numel(), dtype=torch.float32, device=audio.device)])
    for lo in range(0, n_strips, strips_per_call):
        hi = min(n_strips, lo + strips_per_call)
        # overlapping strips as a strided view of the stream (no copy on the host), one H2D per call
        view = audio.as_strided((hi - lo, strip_samples), (k * spec.AUDIO_STRIDE, 1), lo * k * spec.AUDIO_STRIDE)
        chunk = view.to(dev, non_blocking=True).contiguous()
        emb = speech.embed_device(chunk, slot_offsets=offsets)      # [strips, 4k, 96]
        out[lo * k:hi * k] = emb.reshape(-1, 4, spec.EMB_DIM)
    return out[:steps]


def stream_predict(models: Sequence[WakeWordMLPModel], audio, speech: Optional[SpeechEmbeddings] = None,
                   device_id: Optional[int] = None, precision: Optional[str] = None):
    """
    All ``models`` evaluated at every step of the stream -> cuda f32 ``[M, steps - 3]`` (the first classifier input is
    available once the FIFO holds 4 results; column c is step c + 3).
    """
    import torch

    speech = speech or SpeechEmbeddings(device_id=device_id, precision=precision)
    step_emb = stream_step_embeddings(speech, audio)                 # [steps, 4, 96]
    steps = step_emb.shape[0]
    if steps < 4:
        raise ValueError("need at least 4 steps (17280 + 3 * 1920 samples) to fill the 16 x 96 buffer")
    idx = torch.arange(steps - 3, device=step_emb.device)[:, None] + torch.arange(4, device=step_emb.device)[None, :]
    windows = step_emb[idx].reshape(steps - 3, 16, spec.EMB_DIM).contiguous()   # FIFO of the last 4 results
    return MultiWakeWordModel(list(models))(windows)



class WakeWordStreamService:
    """
    Batched multi-stream form of the browser runtime's loop (``HeyBuddy.process``, src/ts/src/hey-buddy.ts:382-469; SURVEY.md 8f
    row 4): ``num_streams`` live audio streams advance in lockstep and every one of them is scored by every wake-word model.

    ``push(audio[S, n])`` appends ``n`` new samples (a multiple of the 1920-sample batch interval) to every stream and returns the
    probabilities of the window steps that became complete: step i covers global samples ``[1920 i, 1920 i + 17280)`` exactly like
    the runtime's batcher, its 4 embeddings join a FIFO of the last 4 steps (``embeddingBufferArray``) and, once that holds 16
    frames, all M models run on it (before that the runtime reports probability 0; so does this, with ``valid`` False).  State per
    stream = the samples of the next, still incomplete window (< 17280) and the last 3 steps' embeddings -- both stay on the
    device.  Featurization of all streams and steps of a push is ONE fully-convolutional pass; all models x all streams x all
    steps are ONE ``hb_mlp_forward_multi`` call.  (Voice-activity gating, recording and callbacks are the runtime's UI layer,
    SURVEY.md 2: out of scope -- every complete step is scored.)
    """

    def __init__(self, models: Sequence[WakeWordMLPModel], num_streams: int, speech: Optional[SpeechEmbeddings] = None,
                 device_id: Optional[int] = None, precision: Optional[str] = None) -> None:
        import torch

        self.speech = speech or SpeechEmbeddings(device_id=device_id, precision=precision)
        self.multi = MultiWakeWordModel(list(models))
        self.num_streams = int(num_streams)
        dev = self.speech.device
        self.pending = torch.zeros((self.num_streams, 0), dtype=torch.float32, device=dev)            # samples of the next window so far
        self.fifo = torch.zeros((self.num_streams, 0, 4, spec.EMB_DIM), dtype=torch.float32, device=dev)  # last <= 3 steps' embeddings
        self.steps_emitted = 0

    def push(self, audio):
        """
        ``audio``: float ``[S, n]`` (tensor / array, [-1, 1]), n a multiple of 1920.  Returns ``(probs, valid)``: cuda f32
        ``[M, S, k]`` for the k steps completed by this push (k may be 0) and a bool list of length k (False while the 16-frame
        buffer is still filling: those columns are 0, as in the runtime).
        """
        import torch

        if isinstance(audio, np.ndarray):
            audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
        assert audio.dim() == 2 and audio.shape[0] == self.num_streams and audio.shape[1] % spec.AUDIO_STRIDE == 0, tuple(audio.shape)
        dev = self.speech.device
        x = torch.cat([self.pending, audio.to(dev, dtype=torch.float32)], dim=1)
        k = len(spec.audio_window_starts(x.shape[1]))
        m = len(self.multi.models)
        if k == 0:
            self.pending = x
            return torch.zeros((m, self.num_streams, 0), dtype=torch.float32, device=dev), []
        emb = _batched_step_embeddings(self.speech, x, k)                          # [S, k, 4, 96]
        self.pending = x[:, k * spec.AUDIO_STRIDE:].contiguous()
        hist = torch.cat([self.fifo, emb], dim=1)                                  # [S, f + k, 4, 96]
        f = self.fifo.shape[1]
        valid = [f + j + 1 >= 4 for j in range(k)]
        probs = torch.zeros((m, self.num_streams, k), dtype=torch.float32, device=dev)
        first = next((j for j, v in enumerate(valid) if v), None)
        if first is not None:
            idx = torch.arange(f + first - 3, f + k - 3, device=dev)[:, None] + torch.arange(4, device=dev)[None, :]     # [k', 4] step indices
            windows = hist[:, idx].reshape(self.num_streams * (k - first), 16, spec.EMB_DIM).contiguous()
            probs[:, :, first:] = self.multi(windows).reshape(m, self.num_streams, k - first)
        self.fifo = hist[:, -3:].contiguous()
        self.steps_emitted += k
        return probs, valid


def _batched_step_embeddings(speech: SpeechEmbeddings, x, k: int):
    """cuda f32 ``[S, L]`` whose first ``k`` 17280-sample windows (stride 1920) are complete -> ``[S, k, 4, 96]``."""
    s = x.shape[0]
    strip = STEPS_PER_STRIP
    n_strips = (k + strip - 1) // strip
    strip_samples = spec.AUDIO_WINDOW + spec.AUDIO_STRIDE * (strip - 1)
    need = (n_strips - 1) * strip * spec.AUDIO_STRIDE + strip_samples
    if need > x.shape[1]:
        import torch

        x = torch.cat([x, torch.zeros((s, need - x.shape[1]), dtype=x.dtype, device=x.device)], dim=1)
    x = x.contiguous()
    offsets = [spec.FRAMES_PER_AUDIO_STRIDE * i + spec.EMB_STRIDE * j for i in range(strip) for j in range(4)]
    view = x.as_strided((s, n_strips, strip_samples), (x.stride(0), strip * spec.AUDIO_STRIDE, 1))
    emb = speech.embed_device(view.reshape(s * n_strips, strip_samples).contiguous(), slot_offsets=offsets)   # [S * strips, 4 * strip, 96]
    return emb.reshape(s, n_strips * strip, 4, spec.EMB_DIM)[:, :k]
