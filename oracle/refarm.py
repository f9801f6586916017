"""
Reference arm harness (TEST INFRASTRUCTURE, see oracle/__init__.py): runs the REFERENCE'S OWN Python for the hot path --
``heybuddy.dataset.augmented.AugmentedAudioGenerator.__call__`` / ``execute_augment_batch`` (augmented.py:297-427) and
``heybuddy.embeddings.SpeechEmbeddings.__call__`` (embeddings.py:153-234), unmodified, from the scratch copy
``baseline/_ref/heybuddy`` that ``__graft_entry__.build()`` makes of ``/root/reference/src/python/heybuddy`` -- on the host
cores.

What the reference delegates to things that cannot exist offline is served by the oracle's CPU restatements, injected at the
reference's own seams:

    onnxruntime sessions (mel-spectrogram.onnx, speech-embedding.onnx)   -> oracle.mel / oracle.embed behind the ring-1 callables
                                                                            ``SpeechEmbeddings().spectrogram`` / ``.embeddings``
    audiomentations.Compose([SevenBandParametricEQ, TanhDistortion])     -> stub module, identity (probabilities 0 in the configs)
    torch_audiomentations.Compose([PitchShift, BandStopFilter,           -> stub module: per_batch coin + parameters drawn per call,
        AddColoredNoise, Gain])                                             oracle.augment.add_colored_noise / gain
    speechbrain.processing.signal_processing.reverberate                 -> stub module, oracle.augment.reverberate
    av, soundfile, numpy.compat, piper_phonemize                         -> empty stub modules (import-time only)

``torchaudio.functional.add_noise`` (augmented.py:272-276) is the real dependency.  Datasets are plain in-memory row lists with
the two methods the reference calls on a HF ``Dataset`` (``shuffle()``, iteration).  The reference's RNG is unseeded
(np.random / torch.rand): this harness is for TIMING and for shape / ordering checks, not for value parity.
"""
from __future__ import annotations

import os
import sys
import types
from typing import Any, Dict, List, Optional

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "heybuddy", "embeddings.py"))


class Rows(list):
    """In-memory stand-in for a HF ``datasets.Dataset`` of audio rows: the reference only calls ``shuffle()`` and iterates."""

    def shuffle(self, *a: Any, **k: Any) -> "Rows":
        return self


def audio_rows(clips, sampling_rate: int = 16000) -> Rows:
    return Rows({"audio": {"array": c, "sampling_rate": sampling_rate}} for c in clips)


def _install_stubs() -> None:
    import torch

    from oracle import augment as oaug

    for m in ("av", "soundfile", "numpy.compat", "piper_phonemize"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["piper_phonemize"].phonemize_espeak = lambda *a, **k: []
    np.compat = sys.modules["numpy.compat"]

    # ---- audiomentations: per-clip numpy transforms (probabilities are 0 in the bench configs) ----
    am = types.ModuleType("audiomentations")

    class _PerClip:
        def __init__(self, p: float = 0.0, **kw: Any) -> None:
            self.p, self.kw = p, kw

        def __call__(self, samples, sample_rate):
            if self.p and np.random.rand() < self.p:
                raise NotImplementedError(f"{type(self).__name__}: not restated in the reference arm (set its probability to 0)")
            return samples

    class _ComposeNp:
        def __init__(self, transforms, **kw: Any) -> None:
            self.transforms = transforms

        def __call__(self, samples, sample_rate):
            for t in self.transforms:
                samples = t(samples, sample_rate)
            return samples

    am.Compose = _ComposeNp
    am.SevenBandParametricEQ = type("SevenBandParametricEQ", (_PerClip,), {})
    am.TanhDistortion = type("TanhDistortion", (_PerClip,), {})
    sys.modules["audiomentations"] = am

    # ---- torch_audiomentations: batch transforms, mode="per_batch"  ----
    tam = types.ModuleType("torch_audiomentations")

    class _Batch:
        def __init__(self, p: float = 0.0, **kw: Any) -> None:
            self.p, self.kw = p, kw

        def apply(self, x: np.ndarray) -> np.ndarray:
            raise NotImplementedError(f"{type(self).__name__}: not restated in the reference arm (set its probability to 0)")

        def __call__(self, batch, sample_rate):
            if self.p and float(torch.rand(())) < self.p:
                x = batch[:, 0, :].cpu().numpy()
                return torch.from_numpy(self.apply(x).astype(np.float32))[:, None, :].to(batch.device)
            return batch

    class AddColoredNoise(_Batch):
        def apply(self, x):
            snr = float(torch.empty(()).uniform_(self.kw["min_snr_in_db"], self.kw["max_snr_in_db"]))
            f_decay = float(torch.empty(()).uniform_(self.kw["min_f_decay"], self.kw["max_f_decay"]))
            base = oaug.colored_noise_base(torch.randn(16000).numpy(), f_decay, dtype=np.float32)
            return oaug.add_colored_noise(x, base, snr, dtype=np.float32)

    class Gain(_Batch):
        def apply(self, x):
            return oaug.gain(x, float(torch.empty(()).uniform_(-18.0, 6.0)), dtype=np.float32)

    class _ComposeT:
        def __init__(self, transforms, **kw: Any) -> None:
            self.transforms = transforms

        def __call__(self, batch, sample_rate):
            for t in self.transforms:
                batch = t(batch, sample_rate)
            return batch

    tam.Compose = _ComposeT
    tam.PitchShift = type("PitchShift", (_Batch,), {})
    tam.BandStopFilter = type("BandStopFilter", (_Batch,), {})
    tam.AddColoredNoise = AddColoredNoise
    tam.Gain = Gain
    sys.modules["torch_audiomentations"] = tam

    # ---- speechbrain reverberate ----
    sb = types.ModuleType("speechbrain")
    sbp = types.ModuleType("speechbrain.processing")
    sbs = types.ModuleType("speechbrain.processing.signal_processing")

    def reverberate(waveforms, rir_waveform, rescale_amp: str = "avg"):
        y = oaug.reverberate(waveforms.cpu().numpy(), rir_waveform.cpu().numpy().reshape(-1), dtype=np.float32)
        return torch.from_numpy(y.astype(np.float32)).to(waveforms.device)

    sbs.reverberate = reverberate
    sb.processing, sbp.signal_processing = sbp, sbs
    sys.modules.update({"speechbrain": sb, "speechbrain.processing": sbp, "speechbrain.processing.signal_processing": sbs})


_loaded: Optional[Dict[str, Any]] = None


def load() -> Dict[str, Any]:
    """Imports the reference package from ``baseline/_ref`` (never from /root/reference in place)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError(f"{REF_DIR}/heybuddy is missing: run __graft_entry__.build() in the build container")
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    _install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from heybuddy.dataset.augmented import AugmentedAudioGenerator  # noqa: E402  (reference)
    from heybuddy.embeddings import SpeechEmbeddings  # noqa: E402  (reference)

    _loaded = {"AugmentedAudioGenerator": AugmentedAudioGenerator, "SpeechEmbeddings": SpeechEmbeddings}
    return _loaded


def speech_embeddings(weights=None, mel_dtype=np.float32):
    """The reference's ``SpeechEmbeddings`` with the oracle's mel / embedding callables behind its ring-1 attributes."""
    from heybuddy_b200 import spec
    from oracle import embed as oembed, mel as omel

    ref = load()
    weights = weights if weights is not None else spec.init_embedding_weights()
    s = ref["SpeechEmbeddings"]()

    def mel_fn(audio):
        m = omel.mel_spectrogram(audio, dtype=mel_dtype)
        return m if m.shape[0] > 1 else m[0]          # the reference squeezes (spectrogram.py:32)

    s.spectrogram = mel_fn
    s.embeddings = lambda windows: oembed.speech_embedding_model(windows, weights).squeeze()   # embeddings.py:42
    return s


def featurize(source_clips: List[np.ndarray], noise_clips: List[np.ndarray], rirs: List[np.ndarray], augment_batch: int = 8,
              f_decay=(0.0, 0.0), weights=None) -> np.ndarray:
    """
    ``len(source_clips)`` clips through the reference's own generator + embeddings: ``AugmentedAudioGenerator(...)(n)`` rows ->
    ``SpeechEmbeddings()(list of clips)`` -> ``f32 [n, 16, 96]`` (what ``TrainingFeaturesGenerator.generate`` does between the
    TTS stage and the return, features.py:431-490, without the debug-sample and tqdm lines).
    """
    ref = load()
    gen = ref["AugmentedAudioGenerator"](
        audio_rows(source_clips), device_id=None, augmentation_dataset=audio_rows(noise_clips), impulse_response_dataset=audio_rows(rirs),
        batch_size=augment_batch, seven_band_aug_prob=0.0, tanh_distortion_prob=0.0, pitch_shift_prob=0.0, band_stop_prob=0.0,
        colored_noise_min_f_decay=f_decay[0], colored_noise_max_f_decay=f_decay[1])
    speech = speech_embeddings(weights)
    clips = [row["audio"]["array"] for row in gen(len(source_clips))]        # features.py:445-447
    return speech(clips, spectrogram_batch_size=32, embedding_batch_size=32)  # features.py:485-490, constants.py:138-139
