"""
In-the-wild extractor vs files written by the reference's own generator classes
(tests/golden/extractor.npz, produced by tests/golden/make_golden_extractor.py with the oracle's mel / embedding behind the
reference pipeline).  CPU: piece / batch / file bookkeeping with the oracle injected.  GPU: the device path end to end.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from extractor_inputs import fake_tokens, synthetic_recordings  # noqa: E402

from heybuddy_b200 import spec
from heybuddy_b200.dataset.extractor import PrecalculatedLabeledTrainingDatasetGenerator, PrecalculatedTrainingDatasetGenerator

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "extractor.npz"))


def _oracle_embed(pieces):
    from oracle import embed as oembed
    from oracle import mel as omel
    from oracle import pipeline as opipe

    weights = spec.init_embedding_weights()
    return opipe.speech_embeddings([p for p in pieces], lambda a: omel.mel_spectrogram(a),
                                   lambda w: oembed.speech_embedding_model(w, weights)).astype(np.float32)


def _check(paths, tag, rtol):
    names = [os.path.basename(p) for p in paths]
    assert names == list(GOLD[f"{tag}_names"]), names
    for p, n in zip(paths, names):
        got, want = np.load(p), GOLD[f"{tag}_{n}"]
        assert got.shape == want.shape and got.dtype == want.dtype
        scale = np.abs(want[:, :16]).max()
        assert np.abs(got[:, :16] - want[:, :16]).max() <= rtol * scale
        if want.shape[1] == 17:
            assert np.array_equal(got[:, 16], want[:, 16])      # token ids, exact


def test_bookkeeping_matches_reference_files(tmp_path):
    class OracleBacked(PrecalculatedTrainingDatasetGenerator):
        def embed_pieces(self, pieces):
            return _oracle_embed(pieces)

    gen = OracleBacked(synthetic_recordings(), process_batch_size=4, gpu_pieces=8)
    seen = []
    paths = gen("wild", output_dir=str(tmp_path), samples_per_file=6, on_progress=lambda d, t: seen.append(d))
    _check(paths, "plain", 1e-6)
    assert seen == [1, 2, 3, 4]
    # max_hours caps the number of batches (2 here); file-name width follows the planned number of files
    paths = gen("capped", output_dir=str(tmp_path), samples_per_file=6, max_hours=2 * 4 * 1.44 / 3600 + 1e-9)
    _check(paths, "capped", 1e-6)
    # one device pass over everything (gpu_pieces large) changes nothing
    paths = OracleBacked(synthetic_recordings(), process_batch_size=4, gpu_pieces=4096)("wild2", output_dir=str(tmp_path), samples_per_file=6)
    for p, n in zip(paths, GOLD["plain_names"]):
        assert np.allclose(np.load(p), GOLD[f"plain_{n}"], atol=1e-6)


def test_labeled_variant_and_tokenizer_requirement(tmp_path):
    class OracleBacked(PrecalculatedLabeledTrainingDatasetGenerator):
        def embed_pieces(self, pieces):
            return _oracle_embed(pieces)

    gen = OracleBacked(synthetic_recordings(), process_batch_size=4, tokenizer=fake_tokens)
    _check(gen("wild", output_dir=str(tmp_path), samples_per_file=6), "labeled", 1e-6)
    with pytest.raises(RuntimeError, match="needs a tokenizer"):
        OracleBacked(synthetic_recordings(), process_batch_size=4)("x", output_dir=str(tmp_path), samples_per_file=6)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("f16", 2e-3)])
def test_device_path_matches_reference_files(tmp_path, cuda_device, precision, rtol):
    gen = PrecalculatedTrainingDatasetGenerator(synthetic_recordings(), process_batch_size=4, gpu_pieces=8, device_id=0, precision=precision)
    _check(gen("wild", output_dir=str(tmp_path), samples_per_file=6), "plain", rtol)
    lab = PrecalculatedLabeledTrainingDatasetGenerator(synthetic_recordings(), process_batch_size=4, device_id=0, precision=precision,
                                                       tokenizer=fake_tokens)
    _check(lab("wild_l", output_dir=str(tmp_path), samples_per_file=6), "labeled", rtol)
