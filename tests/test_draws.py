"""CPU: the seeded draw table follows the reference's call order and does not depend on sharding."""
import numpy as np

from heybuddy_b200 import spec
from heybuddy_b200.dataset.draws import AugmentConfig, DrawTable, colored_noise_base, draw_batch, pad_before_for
from oracle import augment as oaug


def test_table_is_independent_of_sharding():
    rng = np.random.default_rng(0)
    lengths = rng.integers(6400, 22400, size=1000)
    cfg = AugmentConfig(batch_size=128)
    noise_lengths = np.full(64, 160000)
    full = DrawTable.build(lengths, cfg, seed=2004, noise_clip_lengths=noise_lengths, num_rirs=7)
    assert len(full.batches) == 8 and len(full.batches[-1].pad_before) == 1000 - 7 * 128
    # a "rank" that starts at batch 3 with the cursors the prefix left behind reproduces the same rows
    head = DrawTable.build(lengths[:3 * 128], cfg, seed=2004, noise_clip_lengths=noise_lengths, num_rirs=7)
    tail = DrawTable.build(lengths[3 * 128:], cfg, seed=2004, noise_clip_lengths=noise_lengths, num_rirs=7,
                           first_batch=3, noise_cursor=head.final_noise_cursor, rir_cursor=head.final_rir_cursor)
    for a, b in zip(full.batches[3:], tail.batches):
        assert a.index == b.index and np.array_equal(a.pad_before, b.pad_before)
        assert (a.colored_apply, a.gain_db, a.background_apply, a.reverb_apply) == (b.colored_apply, b.gain_db, b.background_apply, b.reverb_apply)
    assert full.noise_clip_cursor[3:] == tail.noise_clip_cursor and full.rir_index[3:] == tail.rir_index


def test_noise_stream_pulls_whole_clips():
    """augmented.py:246-257: clips are pulled until >= B*T samples; the unused tail is dropped."""
    cfg = AugmentConfig(batch_size=128, background_noise_prob=1.0)
    t = DrawTable.build([20000] * 256, cfg, seed=1, noise_clip_lengths=np.full(2048, 160000), num_rirs=0)
    need = int(np.ceil(128 * spec.CLIP_SAMPLES / 160000))  # 19 clips
    assert t.noise_clip_cursor == [0, need]
    assert t.rir_index == [-1, -1]


def test_ranges_and_order():
    cfg = AugmentConfig(batch_size=16, colored_noise_prob=1.0, background_noise_prob=1.0, reverb_prob=1.0)
    for g in range(20):
        d = draw_batch(7, g, [12000] * 16, cfg)
        assert d.colored_apply and d.gain_apply and d.background_apply and d.reverb_apply
        assert 10.0 <= d.colored_snr_db <= 30.0 and -1.0 <= d.colored_f_decay <= 2.0
        assert -18.0 <= d.gain_db <= 6.0
        assert d.noise_snr_db.shape == (16,) and (-10 <= d.noise_snr_db).all() and (d.noise_snr_db <= 15).all()
        s = spec.CLIP_SAMPLES - 12000
        assert ((d.pad_before >= int(s / 4)) & (d.pad_before < int(3 * s / 4))).all()
        assert abs(np.sqrt(np.mean(d.colored_base.astype(np.float64) ** 2)) - 1) < 1e-6
    assert pad_before_for(spec.CLIP_SAMPLES - 1, spec.CLIP_SAMPLES, np.random.default_rng(0)) == 0
    assert pad_before_for(spec.CLIP_SAMPLES + 5, spec.CLIP_SAMPLES, np.random.default_rng(0)) == 0


def test_colored_base_matches_oracle_restatement():
    g = np.random.default_rng(3).standard_normal(16000)
    for f_decay in (0.0, -1.0, 2.0):
        np.testing.assert_allclose(colored_noise_base(g, f_decay), oaug.colored_noise_base(g, f_decay), atol=1e-6)
    # white noise: f_decay = 0 leaves the pattern unchanged up to the RMS normalisation
    np.testing.assert_allclose(colored_noise_base(g, 0.0), g / np.sqrt(np.mean(g * g)), atol=1e-5)


def test_clip_params_vectorised_matches_per_batch_records():
    """
    AugmentedAudioGenerator.clip_params builds the per-clip hb_clip_aug records of a whole draw table in one vectorised
    pass (it runs on the host once per chunk of the streaming path): byte-identical to the per-batch construction it
    replaced, including a ragged last batch and batches without background noise / reverb / coloured noise.
    """
    import types

    from heybuddy_b200 import _native
    from heybuddy_b200.dataset.augmented import AugmentedAudioGenerator

    cfg = AugmentConfig(batch_size=128, colored_noise_min_f_decay=0.0, colored_noise_max_f_decay=0.0)
    rng = np.random.default_rng(7)
    lengths = rng.integers(6400, 22400, size=128 * 9 + 37)
    table = DrawTable.build(lengths, cfg, 2004, noise_clip_lengths=np.full(64, 160000), num_rirs=271,
                            first_batch=3, noise_cursor=5, rir_cursor=2)
    t = spec.CLIP_SAMPLES

    class Bank:
        class _Stream:
            @staticmethod
            def numel():
                return 64 * 160000 + 128 * t

        stream = _Stream()

        @staticmethod
        def offset_of_clip(c):
            return int(c) * 160000

    fake = types.SimpleNamespace(target_num_samples=t, noise_bank=Bank())
    slots, k = [], 0
    for d in table.batches:
        slots.append(k if d.colored_apply else -1)
        k += int(d.colored_apply)
    got = AugmentedAudioGenerator.clip_params(fake, table.batches, table.noise_clip_cursor, table.rir_index, slots)

    recs = []
    for d, ncur, ridx, cslot in zip(table.batches, table.noise_clip_cursor, table.rir_index, slots):
        b = len(d.pad_before)
        r = np.zeros(b, dtype=_native.CLIP_AUG_DTYPE)
        r["gain"] = d.gain_linear
        r["colored_index"] = cslot if d.colored_apply else -1
        r["colored_snr_db"] = d.colored_snr_db
        r["rir_index"] = ridx if d.reverb_apply else -1
        if d.background_apply and ncur >= 0:
            r["noise_offset"] = Bank.offset_of_clip(ncur) + np.arange(b, dtype=np.int64) * t
            r["noise_snr_db"] = d.noise_snr_db
        else:
            r["noise_offset"] = -1
        recs.append(r)
    want = np.concatenate(recs)
    assert got.dtype == want.dtype and got.shape == want.shape == (len(lengths),)
    assert got.tobytes() == want.tobytes()
    assert any(d.background_apply for d in table.batches) and not all(d.background_apply for d in table.batches)
