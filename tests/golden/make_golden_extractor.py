"""
Golden fixture for the in-the-wild extractor, produced by the REFERENCE's own
``PrecalculatedTrainingDatasetGenerator.__call__`` / ``PrecalculatedLabeledTrainingDatasetGenerator``
(``dataset/precalculated.py:40-363``), unmodified, with
  * ``datasets.load_dataset`` answered by the seeded synthetic recordings below (no network in this image),
  * the oracle's mel / embedding callables behind the reference ``SpeechEmbeddings`` (ring 1, as in make_golden.py),
  * a stand-in tokenizer for the labeled variant (the BERT vocabulary is a download).

Run in the build container only:   python tests/golden/make_golden_extractor.py
Writes ``extractor.npz``: file names, row counts and contents of every chunk file the reference wrote.
"""
from __future__ import annotations

import os
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import ROOT, import_reference  # noqa: E402

sys.path.insert(0, ROOT)


from extractor_inputs import fake_tokens, synthetic_recordings  # noqa: E402


def main() -> None:
    import torch

    import_reference()
    import datasets

    from heybuddy_b200 import spec
    from oracle import embed as oembed
    from oracle import mel as omel
    from heybuddy.dataset.precalculated import (PrecalculatedLabeledTrainingDatasetGenerator,  # reference
                                                PrecalculatedTrainingDatasetGenerator)
    from heybuddy.embeddings import SpeechEmbeddings  # reference

    weights = spec.init_embedding_weights()
    s = SpeechEmbeddings()
    s.spectrogram = lambda audio: (lambda m: m if m.shape[0] > 1 else m[0])(omel.mel_spectrogram(audio))
    s.embeddings = lambda windows: oembed.speech_embedding_model(windows, weights)

    datasets.load_dataset = lambda *a, **k: [dict(r, audio=dict(r["audio"])) for r in synthetic_recordings()]
    out = {}
    t0 = time.time()
    for tag, cls in (("plain", PrecalculatedTrainingDatasetGenerator), ("labeled", PrecalculatedLabeledTrainingDatasetGenerator)):
        gen = cls("synthetic", process_batch_size=4, embedding_batch_size=32)
        gen._speech_embeddings = s
        if tag == "labeled":
            gen._tokenizer = lambda text: torch.from_numpy(fake_tokens(text))
        with tempfile.TemporaryDirectory() as d:
            gen("wild", output_dir=d, samples_per_file=6)
            names = sorted(os.listdir(os.path.join(d, "wild")))
            out[f"{tag}_names"] = np.array(names)
            for n in names:
                out[f"{tag}_{n}"] = np.load(os.path.join(d, "wild", n))
        # a short run: max_hours limits the number of 4-piece batches to 2
        if tag == "plain":
            with tempfile.TemporaryDirectory() as d:
                gen("capped", output_dir=d, samples_per_file=6, max_hours=2 * 4 * 1.44 / 3600 + 1e-9)
                names = sorted(os.listdir(os.path.join(d, "capped")))
                out["capped_names"] = np.array(names)
                for n in names:
                    out[f"capped_{n}"] = np.load(os.path.join(d, "capped", n))
    np.savez_compressed(os.path.join(HERE, "extractor.npz"), **out)
    print("wrote extractor.npz in %.0f s:" % (time.time() - t0), {k: (v.shape if v.dtype.kind != "U" else list(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
