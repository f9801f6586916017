"""Dataset side of the hot path (mirrors reference ``heybuddy/dataset/``)."""
_LAZY = {
    "AugmentedAudioGenerator": "heybuddy_b200.dataset.augmented",
    "NoiseBank": "heybuddy_b200.dataset.augmented",
    "RirBank": "heybuddy_b200.dataset.augmented",
    "TrainingFeaturesGenerator": "heybuddy_b200.dataset.features",
    "PrecalculatedDatasetIterator": "heybuddy_b200.dataset.precalculated",
    "TrainingDatasetIterator": "heybuddy_b200.dataset.training",
    "WakeWordTrainingDatasetIterator": "heybuddy_b200.dataset.training",
    # the reference's tests import this stale name (tests/test_training_dataset_generator.py:1)
    "TrainingDatasetGenerator": "heybuddy_b200.dataset.training",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        mod = importlib.import_module(_LAZY[name])
        return getattr(mod, "WakeWordTrainingDatasetIterator" if name == "TrainingDatasetGenerator" else name)
    raise AttributeError(name)
