"""
Oracle: windowing / ordering of the audio -> embeddings pipeline
(TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the reference's own Python control flow:

* ``SpeechEmbeddings.__call__``               embeddings.py:153-234
* ``SpeechEmbeddings.audio_to_spectrograms``  embeddings.py:56-84
* ``SpeechEmbeddings.spectrograms_to_embeddings`` embeddings.py:86-151
* ``audio_to_bct_tensor`` (list / ndarray / int16 handling)  util/audio_util.py:73-145

as executed -- 4 overlapping 17280-sample windows per 23040-sample clip, 4
embedding windows per audio window, results concatenated on axis 1 -- i.e. with
all of the reference's redundant work, so it doubles as the CPU baseline's
control flow.  PINNED against the unmodified reference code driven with the
same injected callables (``tests/golden/pipeline_order.npz``).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

from heybuddy_b200 import spec

MelFn = Callable[[np.ndarray], np.ndarray]      # f32 [b, t] (x32767)      -> f32 [b, F, 32]
EmbedFn = Callable[[np.ndarray], np.ndarray]    # f32 [n, 76, 32, 1]       -> f32 [n, 96]


def audio_to_bt(audio: Union[np.ndarray, Sequence[np.ndarray]]) -> np.ndarray:
    """
    util/audio_util.py:73-145 + embeddings.py:182-186 for array inputs: a list is a
    batch (items truncated to the shortest, :90-101); 1-D = one mono clip; 2-D =
    (channels, time); 3-D = (batch, channels, time); int16 -> /32768.  Returns the
    mono ``f32 [B, T]`` the mel model sees, scaled by 32767.
    """
    def one(a) -> np.ndarray:
        a = np.asarray(a)
        if a.dtype == np.int16:
            a = a.astype(np.float32) / 32768.0
        a = a.astype(np.float32)
        if a.ndim == 1:
            a = a[None, None, :]
        elif a.ndim == 2:
            a = a[None, :, :]
        assert a.ndim == 3
        return a

    if isinstance(audio, (list, tuple)):
        items = [one(a) for a in audio]
        t = min(i.shape[-1] for i in items)
        bct = np.concatenate([i[..., :t] for i in items], axis=0)
    else:
        bct = one(audio)
    bct = bct * np.float32(spec.AUDIO_SCALE)
    if bct.shape[1] > 1:
        bct = bct.mean(axis=1, keepdims=True)
    return bct[:, 0, :]


def audio_to_spectrograms(audio_bt: np.ndarray, mel_fn: MelFn, batch_size: int = 32) -> np.ndarray:
    b, t = audio_bt.shape
    n_frames = spec.reference_frames(t)
    out = np.empty((b, n_frames, spec.N_MELS), dtype=np.float32)
    for i in range(0, max(b, batch_size), batch_size):
        chunk = audio_bt[i:i + batch_size]
        if chunk.shape[0] == 0:
            break
        out[i:i + batch_size] = mel_fn(chunk).reshape(chunk.shape[0], n_frames, spec.N_MELS)
    return out


def spectrograms_to_embeddings(spectrograms: np.ndarray, embed_fn: EmbedFn, batch_size: int = 32) -> np.ndarray:
    b, t, _ = spectrograms.shape
    n_frames = (t - spec.EMB_WINDOW) // spec.EMB_STRIDE + 1
    out = np.empty((b, n_frames, spec.EMB_DIM), dtype=np.float32)
    pending: List[Tuple[int, int, np.ndarray]] = []

    def flush() -> None:
        if not pending:
            return
        windows = np.stack([w for _, _, w in pending])[..., None]
        res = embed_fn(windows).reshape(len(pending), spec.EMB_DIM)
        for k, (i, j, _) in enumerate(pending):
            out[i, j // spec.EMB_STRIDE] = res[k]
        pending.clear()

    for i in range(b):
        for j in range(0, t, spec.EMB_STRIDE):
            w = spectrograms[i, j:j + spec.EMB_WINDOW]
            if w.shape[0] < spec.EMB_WINDOW:
                break
            pending.append((i, j, w))
            if len(pending) >= batch_size:
                flush()
    flush()
    return out


def speech_embeddings(
    audio,
    mel_fn: MelFn,
    embed_fn: EmbedFn,
    spectrogram_batch_size: int = 32,
    embedding_batch_size: int = 32,
    return_spectrograms: bool = False,
):
    """``SpeechEmbeddings.__call__`` as executed (embeddings.py:153-234), remove_nan left out."""
    x = audio_to_bt(audio)
    embs, mels = [], []
    for start in spec.audio_window_starts(x.shape[1]):
        m = audio_to_spectrograms(x[:, start:start + spec.AUDIO_WINDOW], mel_fn, spectrogram_batch_size)
        embs.append(spectrograms_to_embeddings(m, embed_fn, embedding_batch_size))
        mels.append(m)
    e = np.concatenate(embs, axis=1)
    if return_spectrograms:
        s = np.concatenate(mels, axis=1)
        t = s.shape[1]
        return e, s[:, : t - ((t - spec.EMB_WINDOW) % spec.EMB_STRIDE)]
    return e


def augment_table(clips, table, noise_stream: Optional[np.ndarray], noise_clip_starts: Optional[np.ndarray], rir_kernels, dtype=np.float64):
    """
    The oracle's augmentation of ``clips`` (ragged int16 / float arrays) under a draw table
    (``heybuddy_b200.dataset.draws.DrawTable``): per augmentation batch, length fix with the table's pad offsets, then
    ``oracle.augment.augment_batch`` with the batch's draws -- the noise rows are consecutive slices of the contiguous noise
    stream starting at the batch's first noise clip (augmented.py:246-267), the RIR is the batch's (augmented.py:389-392).
    Returns ``f32 [n, T]``.
    """
    from oracle import augment as oaug

    from oracle import k9 as ok9

    t = table.cfg.target_samples
    out, i0 = [], 0
    for k, (d, ncur, ridx) in enumerate(zip(table.batches, table.noise_clip_cursor, table.rir_index)):
        b = len(d.pad_before)
        fixed = np.stack([oaug.to_target_length(c, int(p), t) for c, p in zip(clips[i0:i0 + b], d.pad_before)])
        if table.k9 is not None:      # the per-clip numpy transforms come first (augmented.py:325-328)
            fixed = ok9.apply_table(fixed, table.slice(k, k + 1))
        noise = None
        if d.background_apply:
            off = int(noise_clip_starts[ncur])
            noise = noise_stream[off:off + b * t].reshape(b, t)
        out.append(oaug.augment_batch(
            fixed, colored_base=d.colored_base if d.colored_apply else None, colored_snr_db=d.colored_snr_db,
            gain_db=d.gain_db if d.gain_apply else None, noise=noise, noise_snr_db=d.noise_snr_db,
            rir=rir_kernels[ridx] if d.reverb_apply else None, dtype=dtype))
        i0 += b
    return np.concatenate(out)


def well_conditioned_slots(mel_true: np.ndarray, floor_db: float = 80.0, num_samples: int = spec.CLIP_SAMPLES) -> np.ndarray:
    """
    ``[n, F, 32]`` reference-scale log-mel (``log10(P) + 2``, floor -8) -> bool ``[n, slots]``: True where the slot's 76-frame
    window holds no (frame, mel bin) within ``floor_db`` of the 1e-10 power floor.  log10 of digital silence / of reverb tails that
    decayed below fp32 round-off is implementation noise, and an embedding window that contains such frames inherits it.
    """
    thr = np.log10(spec.MEL_FLOOR) + floor_db / 10.0 + spec.MEL_POST_ADD
    frame_min = mel_true.min(axis=2)
    assert mel_true.shape[1] == spec.mel_frames(num_samples), (mel_true.shape, num_samples)
    offs = spec.embedding_frame_offsets(num_samples)
    return np.stack([frame_min[:, o:o + spec.EMB_WINDOW].min(axis=1) > thr for o in offs], axis=1)
