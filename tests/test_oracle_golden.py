"""CPU: pin the oracle against fixtures produced by the reference itself (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from heybuddy_b200 import spec
from oracle import augment as oaug
from oracle import classifier as ocls
from oracle import embed as oembed
from oracle import mel as omel
from oracle import pipeline as opipe


def _clips(seed, n):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (0.1 * rng.standard_normal((n, spec.CLIP_SAMPLES))).clip(-1, 1).astype(np.float32)


def test_spec_pinned_shapes():
    # tests/test_embeddings.py:10-15 and mel-spectrogram.ts:38-42 of the reference
    assert spec.CLIP_SAMPLES == 23040
    assert spec.mel_frames(17280) == 105 == spec.reference_frames(17280)
    assert spec.mel_frames(23040) == 141
    assert spec.mel_frames(12640) == 76
    assert spec.embedding_frame_offsets(23040) == [0, 8, 16, 24, 12, 20, 28, 36, 24, 32, 40, 48, 36, 44, 52, 60]
    assert len(spec.embedding_frame_offsets(32000)) == 32  # BASELINE config 1: 8 windows x 4
    assert spec.embedding_layer_shapes()[-1] == (1, 1, 96)
    assert spec.embedding_macs_per_window() == 44_144_640
    assert spec.embedding_num_params() == 274_440
    assert spec.classifier_num_params() == 256_417
    assert spec.CLS_HIDDEN == 64
    assert spec.mel_band_limits() == (2, 122)


def test_pipeline_order_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "pipeline_order.npz"))
    clips = _clips(int(g["seed"]), 3)
    weights = spec.init_embedding_weights()
    emb, mels = opipe.speech_embeddings(
        [c for c in clips], omel.mel_spectrogram, lambda w: oembed.speech_embedding_model(w, weights),
        return_spectrograms=True)
    assert emb.shape == (3, 16, 96) and tuple(g["spectrogram_shape"]) == mels.shape == (3, 420, 32)
    np.testing.assert_allclose(emb, g["embeddings"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(mels[:, :8], g["spectrogram_head"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(mels[:, -8:], g["spectrogram_tail"], rtol=0, atol=1e-6)
    one, one_mel = opipe.speech_embeddings(clips[0, :17280], omel.mel_spectrogram,
                                           lambda w: oembed.speech_embedding_model(w, weights), return_spectrograms=True)
    assert one.shape == (1, 4, 96) and one_mel.shape == tuple(g["one_spectrogram_shape"]) == (1, 100, 32)
    np.testing.assert_allclose(one, g["one_embeddings"], rtol=0, atol=1e-6)
    # the reference called the mel model 4x per batch with [B,17280] and the embed model with <=32 windows
    assert [tuple(c) for c in g["mel_calls"][:4]] == [(3, 17280)] * 4
    assert g["emb_call_sizes"].max() <= 32


def test_fully_convolutional_equals_windows():
    """One clip-level evaluation gives every window's embedding exactly."""
    clips = _clips(11, 2) * spec.AUDIO_SCALE
    m = omel.mel_spectrogram(clips)
    w = spec.init_embedding_weights()
    strip = oembed.embed_strip(m[:, :124], w)
    wins = np.stack([oembed.speech_embedding_model(m[:, o:o + 76, :, None], w) for o in range(0, 49, 8)], axis=1)
    np.testing.assert_array_equal(strip, wins)


def test_classifier_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "classifier_hey_buddy.npz"))
    params = {k[len("param::"):]: g[k] for k in g.files if k.startswith("param::")}
    assert [k for k, _ in spec.classifier_param_shapes()] == list(params.keys())
    for k, shp in spec.classifier_param_shapes():
        assert params[k].shape == shp
    rng = np.random.Generator(np.random.PCG64(int(g["x_seed"])))
    x = rng.standard_normal((64, 16, 96)).astype(np.float32)
    x[:16] += 0.5 * rng.standard_normal((1, 1, 96)).astype(np.float32)
    y = np.zeros(64, dtype=np.int64)
    y[:16] = 1
    prob = ocls.forward(x, params)
    np.testing.assert_allclose(prob, g["prob"], rtol=2e-4, atol=1e-7)
    zero = np.load(os.path.join(golden_dir, "classifier_zero_answers.npz"))
    np.testing.assert_allclose(ocls.forward(np.zeros((1, 16, 96), np.float32), params)[0, 0], zero["hey_buddy"], rtol=1e-4)
    # one training step: loss, selection and gradients
    p2, loss, n_sel, grads = ocls.forward_backward_torch(x, y, params, negative_weight=float(g["negative_weight"]),
                                                         high_loss_threshold=float(g["threshold"]))
    assert n_sel == int(g["n_selected"])
    np.testing.assert_allclose(loss, float(g["loss"]), rtol=1e-5)
    for k in g.files:
        if k.startswith("grad::"):
            np.testing.assert_allclose(grads[k[6:]], g[k], rtol=2e-3, atol=1e-7, err_msg=k)
        elif k.startswith("gradnorm::"):
            np.testing.assert_allclose(np.linalg.norm(grads[k[10:]]), float(g[k]), rtol=1e-4, err_msg=k)


def test_add_noise_matches_torchaudio(golden_dir):
    g = np.load(os.path.join(golden_dir, "add_noise.npz"))
    rng = np.random.Generator(np.random.PCG64(int(g["seed"])))
    wav = rng.standard_normal((4, 4096)).astype(np.float32) * 0.1
    noi = rng.standard_normal((4, 4096)).astype(np.float32) * np.array([[0.01], [0.3], [1.0], [5.0]], dtype=np.float32)
    out = oaug.add_noise(wav, noi, g["snr"])
    np.testing.assert_allclose(out, g["out"], rtol=1e-5, atol=1e-6)


def test_mel_oracle_matches_torchaudio_restatement():
    """The numpy restatement equals torchaudio's MelSpectrogram + AmplitudeToDB."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    clips = _clips(5, 2) * spec.AUDIO_SCALE
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=512, win_length=400, hop_length=160,
                                              center=False, n_mels=32, f_min=60.0, f_max=3800.0, power=2.0)
    ref = ms(torch.from_numpy(clips).double()) if False else ms(torch.from_numpy(clips))
    ref_db = (10.0 * torch.log10(torch.clamp(ref, min=1e-10))).transpose(1, 2).numpy() / 10 + 2
    got = omel.mel_spectrogram(clips)
    assert got.shape == (2, 141, 32)
    np.testing.assert_allclose(got, ref_db, rtol=0, atol=2e-5)


def test_reverb_is_circular_and_amplitude_preserving():
    rng = np.random.Generator(np.random.PCG64(3))
    x = rng.standard_normal((2, spec.CLIP_SAMPLES)).astype(np.float32) * 0.1
    rir = np.exp(-np.arange(4000) / 800.0) * rng.standard_normal(4000)
    rir[37] = 3.0
    y = oaug.reverberate(x, rir.astype(np.float32))
    np.testing.assert_allclose(np.abs(y).mean(axis=1), np.abs(x).mean(axis=1), rtol=1e-6)
    # direct time-domain circular convolution of one output sample
    k = oaug.rotate_rir(rir.astype(np.float32), spec.CLIP_SAMPLES).astype(np.float64)
    n = 1234
    direct = sum(x[0, (n - j) % spec.CLIP_SAMPLES].astype(np.float64) * k[j] for j in np.nonzero(k)[0])
    yy = np.fft.irfft(np.fft.rfft(x[0].astype(np.float64)) * np.fft.rfft(k), n=spec.CLIP_SAMPLES)
    np.testing.assert_allclose(yy[n], direct, rtol=1e-9)
    # a unit impulse at the peak is the identity
    imp = np.zeros(100, dtype=np.float32)
    imp[10] = 1.0
    np.testing.assert_allclose(oaug.reverberate(x, imp), x, rtol=0, atol=1e-7)


def test_to_target_length():
    a = (np.arange(100) - 50).astype(np.int16)
    out = oaug.to_target_length(a, pad_before=7, target=128)
    assert out.dtype == np.float32 and out.shape == (128,)
    np.testing.assert_array_equal(out[7:107], a.astype(np.float32) / 32768.0)
    assert out[:7].sum() == 0 and out[107:].sum() == 0
    long = np.ones(300, dtype=np.float32)
    assert oaug.to_target_length(long, 0, target=128).shape == (128,)


def test_reverb_independent_cross_check():
    """
    `oracle.augment.reverberate` restates speechbrain's reverberate (absent offline).  Independent routes to the same definition
    (SURVEY.md A.3 item 5), none of which touches rotate_rir / numpy's rfft:
      (a) scipy.signal.convolve (direct, time domain) of the clip with the RIR, wrapped modulo T, then rolled left by the peak
          delay d = argmax|rir| -- circular convolution with [rir[d:], 0.., rir[:d]] is the un-rotated circular convolution delayed by -d;
      (b) torch.fft on the whole batch.
    Both followed by the avg-amplitude rescale written out directly.
    """
    torch = pytest.importorskip("torch")
    signal = pytest.importorskip("scipy.signal")
    rng = np.random.Generator(np.random.PCG64(17))
    t = spec.CLIP_SAMPLES
    x = (rng.standard_normal((3, t)) * 0.1).astype(np.float32)
    for ln, peak in ((4000, 37), (24000, 150), (30000, 0)):            # shorter than, about, and longer than the clip (truncated to T)
        rir = (np.exp(-np.arange(ln) / 900.0) * rng.standard_normal(ln)).astype(np.float32)
        rir[peak] = 3.5
        want = oaug.reverberate(x, rir)
        r = rir[:t].astype(np.float64)
        d = int(np.argmax(np.abs(rir)))
        # (a) direct convolution, wrapped, rolled
        full = np.stack([signal.convolve(xi.astype(np.float64), r, mode="full", method="direct") for xi in x[:1]])
        wrapped = full[:, :t].copy()
        wrapped[:, :full.shape[1] - t] += full[:, t:]
        ya = np.roll(wrapped, -d, axis=1)
        ya = ya * np.abs(x[:1].astype(np.float64)).mean(axis=1, keepdims=True) / (np.abs(ya).mean(axis=1, keepdims=True) + 1e-14)
        np.testing.assert_allclose(want[:1], ya, atol=2e-7 * np.abs(ya).max())
        # (b) torch.fft, whole batch
        k = torch.zeros(t, dtype=torch.float64)
        k[:r.shape[0]] = torch.from_numpy(r)
        k = torch.roll(k, -d)
        xb = torch.from_numpy(x).double()
        yb = torch.fft.irfft(torch.fft.rfft(xb) * torch.fft.rfft(k), n=t)
        yb = yb * xb.abs().mean(dim=1, keepdim=True) / (yb.abs().mean(dim=1, keepdim=True) + 1e-14)
        np.testing.assert_allclose(want, yb.numpy(), atol=2e-7 * float(yb.abs().max()))
