"""
heybuddy_b200 -- B200-native (sm_100a) implementation of Hey Buddy's featurization hot path:
batched augmentation -> 32-bin log-mel -> 96-d speech embeddings -> .npy memmaps, plus the
wake-word classifier that consumes them.  Module names mirror the reference package
(``heybuddy.embeddings`` -> ``heybuddy_b200.embeddings`` ...); see INTEGRATION.md.
"""
__version__ = "0.1.0"

_LAZY = {
    "SpeechEmbeddings": "heybuddy_b200.embeddings",
    "SpeechEmbeddingModel": "heybuddy_b200.embeddings",
    "get_speech_embeddings": "heybuddy_b200.embeddings",
    "MelSpectrogramModel": "heybuddy_b200.spectrogram",
    "get_mel_spectrogram_model": "heybuddy_b200.spectrogram",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(_LAZY[name]), name)
    raise AttributeError(name)
