"""GPU: hb_mel_f32 (through the reference-shaped MelSpectrogramModel) vs the CPU oracle."""
import numpy as np
import pytest

from heybuddy_b200 import spec
from oracle import mel as omel

pytestmark = pytest.mark.gpu

# north star: mel within 1e-4 relative in fp32.  The output is log10(P)+2 (|values| ~ 1..16); the test
# bounds |got - want| <= 1e-4 * max(1, |want|) against the float64 oracle.
MEL_RTOL = 1e-4


def _assert_mel_close(got, want):
    assert got.shape == want.shape
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    assert err.max() <= MEL_RTOL, f"max rel err {err.max():.3e}"


def _clips(seed, n, t):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (0.1 * rng.standard_normal((n, t))).clip(-1, 1).astype(np.float32) * np.float32(spec.AUDIO_SCALE)


@pytest.mark.parametrize("t,frames", [(17280, 105), (23040, 141), (12640, 76), (32000, 197), (512, 1), (671, 1), (672, 2)])
def test_mel_shapes_and_values(cuda_device, t, frames):
    from heybuddy_b200.spectrogram import MelSpectrogramModel

    model = MelSpectrogramModel(device_id=0)
    audio = _clips(100 + t, 5, t)
    got = model(audio)
    want = omel.mel_spectrogram(audio)
    assert want.shape == (5, frames, 32)
    _assert_mel_close(got, np.squeeze(want))  # the reference squeezes the model output (spectrogram.py:32)


def test_mel_single_clip_squeezes_like_reference(cuda_device):
    from heybuddy_b200.spectrogram import MelSpectrogramModel

    model = MelSpectrogramModel(device_id=0)
    audio = _clips(1, 1, 17280)
    got = model(audio[0])  # 1-D input (spectrogram.py:28-29)
    assert got.shape == (105, 32)
    _assert_mel_close(got[None], omel.mel_spectrogram(audio))
    raw = super(MelSpectrogramModel, model).__call__(input=audio)[0]  # ORT-style named input -> [B,1,F,32] dB
    assert raw.shape == (1, 1, 105, 32)
    np.testing.assert_allclose(np.squeeze(raw) / 10 + 2, got, atol=2e-5)


def test_mel_edge_inputs(cuda_device):
    from heybuddy_b200.spectrogram import MelSpectrogramModel

    model = MelSpectrogramModel(device_id=0)
    # silence -> floor: log10(1e-10) + 2 = -8
    z = model(np.zeros((2, 17280), dtype=np.float32))
    np.testing.assert_allclose(z, -8.0, atol=1e-6)
    # full-scale square wave and a pure tone (large dynamic range across bins)
    t = np.arange(23040)
    sq = (np.sign(np.sin(2 * np.pi * 440 * t / 16000)) * 32767).astype(np.float32)
    tone = (np.sin(2 * np.pi * 1000 * t / 16000) * 20000).astype(np.float32)
    audio = np.stack([sq, tone])
    got, want = model(audio), omel.mel_spectrogram(audio)
    # far-from-the-tone bins hold fp32 leakage 1e-12 relative to the peak; compare in the power domain there
    big = want > want.max() - 6.0
    _assert_mel_close(got[big], want[big])
    assert np.abs(got - want).max() < 5e-2
    # too short for a frame -> empty
    assert model(np.zeros((3, 100), dtype=np.float32)).shape == (3, 0, 32)


def test_mel_window_independence(cuda_device):
    """Frame f of audio window w equals global frame 12 w + f: bit-identical."""
    from heybuddy_b200.spectrogram import MelSpectrogramModel

    model = MelSpectrogramModel(device_id=0)
    audio = _clips(9, 3, 23040)
    full = model(audio)
    for w in range(4):
        part = model(np.ascontiguousarray(audio[:, 1920 * w:1920 * w + 17280]))
        np.testing.assert_array_equal(part, full[:, 12 * w:12 * w + 105])


def test_mel_large_batch_matches_small(cuda_device):
    from heybuddy_b200.spectrogram import MelSpectrogramModel

    model = MelSpectrogramModel(device_id=0)
    audio = _clips(21, 300, 23040)
    got = model(audio)
    np.testing.assert_array_equal(got[17], model(audio[17]))
    _assert_mel_close(got[::37], omel.mel_spectrogram(audio[::37]))
