"""K9: SevenBandParametricEQ + TanhDistortion -- draws and filter design on the CPU, kernels on the GPU."""
import numpy as np
import pytest

from heybuddy_b200 import spec
from heybuddy_b200.dataset.draws import AugmentConfig, DrawTable
from heybuddy_b200.dataset.k9 import EQ_BANDS, biquad_sos
from oracle import k9 as ok9


def _response(sos5, f, sr=16000):
    z = np.exp(-2j * np.pi * f / sr)
    return (sos5[0] + sos5[1] * z + sos5[2] * z * z) / (1.0 + sos5[3] * z + sos5[4] * z * z)


def test_biquad_design_hits_its_gain():
    """RBJ forms: a peaking filter has gain_db at its centre and 0 dB far away; shelves have gain_db on their side and 0 dB on the other."""
    for gain_db in (-6.0, 3.5, 6.0):
        pk = biquad_sos("peaking", np.array([1000.0]), np.array([gain_db]), np.array([2.0]))[0]
        assert abs(20 * np.log10(abs(_response(pk, 1000.0))) - gain_db) < 1e-9
        assert abs(20 * np.log10(abs(_response(pk, 10.0)))) < 0.05 and abs(20 * np.log10(abs(_response(pk, 7990.0)))) < 0.05
        lo = biquad_sos("low_shelf", np.array([60.0]), np.array([gain_db]), np.array([0.7]))[0]
        assert abs(20 * np.log10(abs(_response(lo, 1e-3))) - gain_db) < 1e-3 and abs(20 * np.log10(abs(_response(lo, 7999.0)))) < 1e-3
        hi = biquad_sos("high_shelf", np.array([5000.0]), np.array([gain_db]), np.array([0.7]))[0]
        assert abs(20 * np.log10(abs(_response(hi, 7999.99))) - gain_db) < 1e-2 and abs(20 * np.log10(abs(_response(hi, 1.0)))) < 1e-3
    # a centre frequency above Nyquist is pulled below it and the filter stays stable (poles inside the unit circle)
    hs = biquad_sos("high_shelf", np.array([9486.0]), np.array([6.0]), np.array([0.1]))[0]
    assert np.all(np.abs(np.roots([1.0, hs[3], hs[4]])) < 1.0)


def test_k9_draws_are_counter_based_and_sliceable():
    cfg = AugmentConfig(batch_size=16, seven_band_prob=0.25, tanh_distortion_prob=0.25)
    lengths = np.random.default_rng(0).integers(6400, 22400, size=16 * 12 + 5)
    t = DrawTable.build(lengths, cfg, 99, np.full(8, 16 * spec.CLIP_SAMPLES), 3)
    k = t.k9
    assert k is not None and k.eq_apply.shape == (len(lengths),) and 0.1 < k.eq_apply.mean() < 0.45 and 0.1 < k.tanh_apply.mean() < 0.45
    assert k.eq_sos.shape == (int(k.eq_apply.sum()), len(EQ_BANDS), 5)
    assert ((k.tanh_amount[k.tanh_apply] >= 1e-4) & (k.tanh_amount[k.tanh_apply] <= 0.1)).all() and (k.tanh_amount[~k.tanh_apply] == 0).all()
    # a table built for batches [5, 9) alone holds the same draws as the slice (world-size independence)
    part = DrawTable.build(lengths[5 * 16:9 * 16], cfg, 99, np.full(8, 16 * spec.CLIP_SAMPLES), 3, first_batch=5)
    sl = t.slice(5, 9)
    for a, b in ((part.k9.eq_apply, sl.k9.eq_apply), (part.k9.eq_sos, sl.k9.eq_sos), (part.k9.tanh_amount, sl.k9.tanh_amount)):
        np.testing.assert_array_equal(a, b)
    with pytest.raises(NotImplementedError):
        DrawTable.build(lengths, AugmentConfig(batch_size=16, pitch_shift_prob=0.25), 99)


def test_oracle_tanh_distortion_properties():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(spec.CLIP_SAMPLES) * 0.1).astype(np.float32)
    for amount in (1e-4, 0.01, 0.1):
        y = ok9.tanh_distortion(x, amount)
        assert abs(np.sqrt(np.mean(y.astype(np.float64) ** 2)) / np.sqrt(np.mean(x.astype(np.float64) ** 2)) - 1) < 1e-6   # loudness matched
        assert np.all(np.sign(y) == np.sign(x))
    # more distortion = more compression of the peaks relative to the RMS
    crest = lambda v: np.abs(v).max() / np.sqrt(np.mean(v.astype(np.float64) ** 2))
    assert crest(ok9.tanh_distortion(x, 0.1)) < crest(ok9.tanh_distortion(x, 1e-4)) <= crest(x) + 1e-6
    assert np.array_equal(ok9.tanh_distortion(np.zeros(100, np.float32), 0.05), np.zeros(100, np.float32))


@pytest.mark.gpu
def test_k9_kernels_match_oracle(cuda_device):
    """hb_k9_eq_f32 / hb_k9_tanh_f32 on the selected clips vs scipy.sosfilt / numpy percentile + tanh (restated, parity unpinned)."""
    import torch

    from heybuddy_b200.dataset import k9

    cfg = AugmentConfig(batch_size=8, seven_band_prob=0.5, tanh_distortion_prob=0.5)
    rng = np.random.default_rng(2)
    n = 64
    table = DrawTable.build(rng.integers(6400, 22400, size=n), cfg, 7)
    fixed = np.zeros((n, spec.CLIP_SAMPLES), dtype=np.float32)
    for i in range(n):   # zero-padded bursts like length-fixed clips, plus one all-zero clip and one full-scale clip
        ln = int(rng.integers(6400, 22400))
        fixed[i, 3000:3000 + ln] = (rng.standard_normal(ln) * rng.uniform(0.01, 0.3)).astype(np.float32)[:spec.CLIP_SAMPLES - 3000]
    fixed[5] = 0.0
    fixed[6] = np.sign(fixed[6]) * 1.0
    want = ok9.apply_table(fixed, table)
    got = k9.apply_device(torch.from_numpy(fixed.copy()).cuda(), table).cpu().numpy()
    both = table.k9.eq_apply | table.k9.tanh_apply
    assert both.sum() > 20 and (~both).sum() > 5 and (table.k9.eq_apply & table.k9.tanh_apply).any()
    np.testing.assert_array_equal(got[~both], fixed[~both])                      # untouched clips stay bit-identical
    scale = np.maximum(np.abs(want).max(axis=1, keepdims=True), 1e-12)
    err = np.abs(got - want) / scale
    assert err.max() < 1e-4, (err.max(), int(err.max(axis=1).argmax()))
    assert np.isfinite(got).all()


@pytest.mark.gpu
def test_generator_with_k9_matches_oracle(cuda_device):
    """The public generator with EQ + distortion enabled (staged device path) vs the oracle under the same draw table."""
    import torch

    from heybuddy_b200.dataset.features import SyntheticSpeechSource, TrainingFeaturesGenerator
    from oracle import embed as oembed, mel as omel, pipeline as opipe

    rng = np.random.default_rng(3)
    noise = (rng.standard_normal((24, 40000)) * 0.2).astype(np.float32)
    rirs = []
    for _ in range(4):
        ln = int(rng.integers(3200, 24000))
        r = np.exp(-np.arange(ln) / rng.uniform(300, 3000)) * rng.standard_normal(ln)
        r[int(rng.integers(0, 200))] = 4.0
        rirs.append(r.astype(np.float32))
    n = 48
    gen = TrainingFeaturesGenerator(device_id=0, use_autoconfigure=False, augment_batch_size=8, augment_background_dataset=noise,
                                    augment_impulse_dataset=rirs, augment_seven_band_prob=0.5, augment_tanh_distortion_prob=0.5,
                                    precision="fp32", seed=21, source=SyntheticSpeechSource(5))
    got = gen(n)
    _, aug = gen._pipeline(True)
    clips = SyntheticSpeechSource(5)(n)
    table = DrawTable.build([c.shape[0] for c in clips], aug.cfg, 21, aug.noise_bank.clip_lengths, len(aug.rir_bank))
    assert table.k9 is not None and table.k9.eq_apply.any() and table.k9.tanh_apply.any()
    audio = opipe.augment_table(clips, table, aug.noise_bank.stream.cpu().numpy(), aug.noise_bank.clip_starts, aug.rir_bank.kernels_host)
    weights = spec.init_embedding_weights()
    want = opipe.speech_embeddings([x for x in audio], omel.mel_spectrogram, lambda w: oembed.speech_embedding_model(w, weights, dtype=torch.float64))
    good = opipe.well_conditioned_slots(omel.mel_spectrogram(audio * np.float32(spec.AUDIO_SCALE)))
    err = np.abs(got - want).max(axis=2) / np.abs(want).max()
    assert good.mean() > 0.3 and err[good].max() < 1e-3, err[good].max()
