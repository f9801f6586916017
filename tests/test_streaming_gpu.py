"""GPU: streaming sliding-window inference with several wake-word models (BASELINE config 5)."""
import numpy as np
import pytest
import torch

from heybuddy_b200 import spec
from oracle import classifier as ocls
from oracle import embed as oembed
from oracle import mel as omel

pytestmark = pytest.mark.gpu


def _oracle_steps(stream, weights):
    """Browser semantics (hey-buddy.ts:382-413): every 1920 samples, last 17280 samples -> mel -> 4 embeddings."""
    out = []
    for start in spec.audio_window_starts(stream.shape[0]):
        m = omel.mel_spectrogram(stream[None, start:start + spec.AUDIO_WINDOW] * np.float32(spec.AUDIO_SCALE))[0]
        wins = np.stack([m[o:o + 76] for o in (0, 8, 16, 24)])[..., None]
        out.append(oembed.speech_embedding_model(wins, weights, dtype=torch.float64))
    return np.stack(out)  # [steps, 4, 96]


@pytest.mark.parametrize("seconds", [2.5, 9.0])
def test_stream_matches_per_step_oracle(cuda_device, seconds):
    from heybuddy_b200.embeddings import SpeechEmbeddings
    from heybuddy_b200.streaming import num_stream_steps, stream_predict, stream_step_embeddings
    from heybuddy_b200.wakeword import WakeWordMLPModel

    rng = np.random.Generator(np.random.PCG64(5001))
    n = int(seconds * 16000)
    stream = (0.1 * rng.standard_normal(n)).clip(-1, 1).astype(np.float32)
    steps = num_stream_steps(n)
    assert steps == (n - 17280) // 1920 + 1
    weights = spec.init_embedding_weights()
    want = _oracle_steps(stream, weights)
    for precision, tol in (("fp32", 1e-4), ("f16", 1e-3)):
        speech = SpeechEmbeddings(device_id=0, precision=precision)
        got = stream_step_embeddings(speech, stream).cpu().numpy()
        assert got.shape == (steps, 4, 96)
        assert np.abs(got - want).max() / np.abs(want).max() < tol, precision
    # N models on the same rolling [16, 96] buffer: column c = step c + 3
    models = [WakeWordMLPModel(device_id=0, seed=5002 + i) for i in range(5)]
    probs = stream_predict(models, stream, speech=speech).cpu().numpy()
    assert probs.shape == (5, steps - 3)
    fifo = np.stack([got[c:c + 4].reshape(16, 96) for c in range(steps - 3)])
    for i in (0, 4):
        ref = ocls.forward(fifo, spec.init_classifier_weights(5002 + i))[:, 0]
        np.testing.assert_allclose(probs[i], ref, rtol=1e-3, atol=1e-6)


def test_long_strip_geometry(cuda_device):
    """A 477-frame strip (32 steps) exercises the time-tiled tail and many-tile trunk of hb_embed_clips."""
    from heybuddy_b200.embeddings import SpeechEmbeddingModel

    rng = np.random.Generator(np.random.PCG64(9))
    audio = (0.1 * rng.standard_normal((2, 512 + 160 * 476))).clip(-1, 1).astype(np.float32) * np.float32(spec.AUDIO_SCALE)
    m = omel.mel_spectrogram(audio)
    offs = [12 * s + 8 * j for s in range(32) for j in range(4)]
    w = spec.init_embedding_weights()
    want = np.stack([oembed.speech_embedding_model(m[:, o:o + 76, :, None], w, dtype=torch.float64) for o in offs], axis=1)
    for precision, tol in (("fp32", 1e-4), ("f16", 1e-3)):
        model = SpeechEmbeddingModel(device_id=0, precision=precision, load=True)
        got = model.run_clips_device(torch.from_numpy(m).cuda(), offs).cpu().numpy()
        assert np.abs(got - want).max() / np.abs(want).max() < tol, precision


def test_stream_service_matches_offline(cuda_device):
    """
    The batched multi-stream service (browser loop semantics, hey-buddy.ts:382-469): three streams pushed in uneven chunks give,
    step for step, what the offline strip evaluation of each whole stream gives; the first three steps report 0 / not valid
    while the 16-frame buffer fills.
    """
    from heybuddy_b200.embeddings import SpeechEmbeddings
    from heybuddy_b200.streaming import WakeWordStreamService, num_stream_steps, stream_predict
    from heybuddy_b200.wakeword import WakeWordMLPModel

    rng = np.random.Generator(np.random.PCG64(77))
    n = 1920 * 60
    streams = (0.1 * rng.standard_normal((3, n))).clip(-1, 1).astype(np.float32)
    speech = SpeechEmbeddings(device_id=0, precision="f16")
    models = [WakeWordMLPModel(device_id=0, seed=5002 + i) for i in range(4)]
    want = np.stack([stream_predict(models, s, speech=speech).cpu().numpy() for s in streams], axis=1)     # [M, S, steps - 3]
    steps = num_stream_steps(n)
    service = WakeWordStreamService(models, num_streams=3, speech=speech)
    got, valid, at = [], [], 0
    for chunk in (1920 * 5, 1920 * 4, 1920, 1920 * 37, 1920 * 2, 1920 * 11):
        p, v = service.push(streams[:, at:at + chunk])
        at += chunk
        got.append(p.cpu().numpy())
        valid += v
    got = np.concatenate(got, axis=2)
    assert at == n and got.shape == (4, 3, steps) and service.steps_emitted == steps
    assert valid == [False] * 3 + [True] * (steps - 3) and not got[:, :, :3].any()
    np.testing.assert_allclose(got[:, :, 3:], want, rtol=0, atol=2e-6)
