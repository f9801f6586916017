"""Seeded inputs shared by make_golden_extractor.py (reference run) and the extractor tests."""
import numpy as np


def synthetic_recordings(seed: int = 7001):
    """Stand-in for an in-the-wild dataset: ragged recordings, one at 8 kHz, one with a NaN sample."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lengths = [30000, 23040, 61000, 9000, 52000, 47000]
    recs = []
    for i, n in enumerate(lengths):
        rate = 8000 if i == 3 else 16000
        x = (0.1 * rng.standard_normal(n)).clip(-1, 1).astype(np.float32)
        if i == 4:
            x[30000] = np.nan                       # poisons the second 1.44 s piece of this recording
        recs.append({"audio": {"array": x, "sampling_rate": rate}, "transcript": f"utterance number {i % 3}", "id": i})
    return recs


def fake_tokens(text: str, length: int = 96) -> np.ndarray:
    """Deterministic stand-in for BERTTokenizer(length=96)(text): int64 [96]."""
    out = np.zeros(length, dtype=np.int64)
    codes = [101] + [1000 + (ord(c) * 7) % 2000 for c in text][: length - 2] + [102]
    out[: len(codes)] = codes
    return out
