"""CPU: the seeded draw table follows the reference's call order and does not depend on sharding."""
import numpy as np

from heybuddy_b200 import spec
from heybuddy_b200.dataset.draws import (AugmentConfig, DrawTable, advance_noise_cursor, colored_noise_base, gaussian_pattern,
                                         pad_before_from_uniform, philox4x32)
from oracle import augment as oaug


def test_philox_known_answers():
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors): the draw table is a standard generator."""
    M = 0xFFFFFFFF
    cases = [((0, 0, 0, 0), 0, (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
             ((M, M, M, M), (M << 32) | M, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
             ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0x299F31D0 << 32) | 0xA4093822, (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, want in cases:
        assert tuple(int(v) for v in philox4x32(*ctr, key)) == want


def test_table_is_independent_of_sharding():
    rng = np.random.default_rng(0)
    lengths = rng.integers(6400, 22400, size=1000)
    cfg = AugmentConfig(batch_size=128)
    noise_lengths = rng.integers(100000, 200000, size=64)
    full = DrawTable.build(lengths, cfg, seed=2004, noise_clip_lengths=noise_lengths, num_rirs=7)
    assert len(full.batches) == 8 and len(full.batches[-1].pad_before) == 1000 - 7 * 128
    # a "rank" that starts at batch 3 with the cursors the prefix left behind reproduces the same rows
    head = DrawTable.build(lengths[:3 * 128], cfg, seed=2004, noise_clip_lengths=noise_lengths, num_rirs=7)
    tail = DrawTable.build(lengths[3 * 128:], cfg, seed=2004, noise_clip_lengths=noise_lengths, num_rirs=7,
                           first_batch=3, noise_cursor=head.final_noise_cursor, rir_cursor=head.final_rir_cursor)
    for a, b in zip(full.batches[3:], tail.batches):
        assert a.index == b.index and np.array_equal(a.pad_before, b.pad_before)
        assert (a.colored_apply, a.gain_db, a.background_apply, a.reverb_apply) == (b.colored_apply, b.gain_db, b.background_apply, b.reverb_apply)
        assert (a.noise_snr_db is None) == (b.noise_snr_db is None) and (a.noise_snr_db is None or np.array_equal(a.noise_snr_db, b.noise_snr_db))
    assert full.noise_clip_cursor[3:] == tail.noise_clip_cursor and full.rir_index[3:] == tail.rir_index
    # slicing a table = building the slice
    sl = full.slice(3, 8)
    assert sl.first_batch == 3 and sl.n_clips == 1000 - 3 * 128 and np.array_equal(sl.pad_before, tail.pad_before)
    assert sl.noise_clip_cursor == tail.noise_clip_cursor and np.array_equal(sl.colored_snr_db, tail.colored_snr_db)


def test_noise_stream_pulls_whole_clips():
    """augmented.py:246-257: clips are pulled until >= B*T samples; the unused tail is dropped."""
    cfg = AugmentConfig(batch_size=128, background_noise_prob=1.0)
    t = DrawTable.build([20000] * 256, cfg, seed=1, noise_clip_lengths=np.full(2048, 160000), num_rirs=0)
    need = int(np.ceil(128 * spec.CLIP_SAMPLES / 160000))  # 19 clips
    assert t.noise_clip_cursor == [0, need]
    assert t.rir_index == [-1, -1]


def test_advance_noise_cursor_matches_the_reference_loop():
    """The vectorised cursor advance against the literal `while got < need: got += len(next clip)` loop, ragged clips, wrapping."""
    rng = np.random.default_rng(5)
    lengths = rng.integers(1, 50, size=37)
    starts = np.concatenate(([0], np.cumsum(lengths)))
    for need in [1, 2, 17, int(lengths.sum()) - 1, int(lengths.sum()), int(lengths.sum()) + 1, 5 * int(lengths.sum()) + 13]:
        for cursor in [0, 1, 17, 36]:
            c, got = cursor, 0
            while got < need:
                got += int(lengths[c % 37])
                c += 1
            assert advance_noise_cursor(cursor, need, starts) == c % 37, (need, cursor)


def test_ranges_and_order():
    cfg = AugmentConfig(batch_size=16, colored_noise_prob=1.0, background_noise_prob=1.0, reverb_prob=1.0)
    table = DrawTable.build([12000] * (16 * 20), cfg, 7, noise_clip_lengths=np.full(8, 16 * spec.CLIP_SAMPLES), num_rirs=3)
    for d in table.batches:
        assert d.colored_apply and d.gain_apply and d.background_apply and d.reverb_apply
        assert 10.0 <= d.colored_snr_db <= 30.0 and -1.0 <= d.colored_f_decay <= 2.0
        assert -18.0 <= d.gain_db <= 6.0
        assert d.noise_snr_db.shape == (16,) and (-10 <= d.noise_snr_db).all() and (d.noise_snr_db <= 15).all()
        s = spec.CLIP_SAMPLES - 12000
        assert ((d.pad_before >= int(s / 4)) & (d.pad_before < int(3 * s / 4))).all()
    assert abs(np.sqrt(np.mean(table.batches[3].colored_base.astype(np.float64) ** 2)) - 1) < 1e-6
    assert len(set(table.pad_before.tolist())) > 100 and len(set(np.round(table.gain_db, 6).tolist())) == 20
    assert table.rir_index == [g % 3 for g in range(20)]
    # the reference's pad rule (augmented.py:216-226): s == 1 pads right, s <= 0 no pad, randint(int(s/4), int(3s/4)) otherwise
    t = spec.CLIP_SAMPLES
    u = np.array([0.0, 0.5, 0.999999])
    assert pad_before_from_uniform([t - 1] * 3, t, u).tolist() == [0, 0, 0]
    assert pad_before_from_uniform([t + 5] * 3, t, u).tolist() == [0, 0, 0]
    assert pad_before_from_uniform([t - 2] * 3, t, u).tolist() == [0, 0, 0]          # randint(0, 1)
    assert pad_before_from_uniform([t - 100] * 3, t, u).tolist() == [25, 50, 74]     # randint(25, 75)


def test_gaussian_pattern_and_colored_base():
    g = gaussian_pattern(2004, 11)
    assert g.shape == (16000,) and abs(g.mean()) < 0.03 and abs(g.std() - 1) < 0.03 and 3.0 < np.abs(g).max() < 6.0
    assert not np.array_equal(g, gaussian_pattern(2004, 12)) and np.array_equal(g, gaussian_pattern(2004, 11))
    for f_decay in (0.0, -1.0, 2.0):
        np.testing.assert_allclose(colored_noise_base(g, f_decay), oaug.colored_noise_base(g, f_decay), atol=1e-6)
    # white noise: f_decay = 0 leaves the pattern unchanged up to the RMS normalisation
    np.testing.assert_allclose(colored_noise_base(g, 0.0), g / np.sqrt(np.mean(g * g)), atol=1e-5)


def test_colored_noise_base_independent_cross_check():
    """
    `colored_noise_base` restates torch_audiomentations' `_gen_noise` with numpy FFTs.  An independent route to the same definition
    (SURVEY.md A.3 item 2): torch.fft instead of pocketfft, and the shaping applied as an explicit circular convolution with the
    impulse response of the 1/f^decay mask instead of a spectral product.
    """
    import torch

    g = gaussian_pattern(7, 3)
    for f_decay in (-1.0, 0.5, 2.0):
        mask = 1.0 / torch.linspace(1.0, (16000 / 2) ** 0.5, 8001, dtype=torch.float64) ** f_decay
        h = torch.fft.irfft(mask.to(torch.complex128), n=16000)                   # impulse response of the mask (real, even)
        gt = torch.from_numpy(g)
        idx = (torch.arange(16000)[:, None] - torch.arange(16000)[None, :]) % 16000
        c = (h[idx] @ gt)                                                         # direct circular convolution, no FFT of g
        c = (c / torch.sqrt(torch.mean(c * c))).numpy()
        np.testing.assert_allclose(colored_noise_base(g, f_decay), c, atol=2e-6)


def test_clip_records_vectorised_matches_per_batch_records():
    """
    DrawTable.clip_records builds the per-clip hb_clip_aug records of a whole draw table in one vectorised pass (it runs on the
    host once per chunk of the streaming path): byte-identical to a per-batch construction, including a ragged last batch and
    batches without background noise / reverb / coloured noise.
    """
    from heybuddy_b200 import _native

    cfg = AugmentConfig(batch_size=128, colored_noise_min_f_decay=0.0, colored_noise_max_f_decay=0.0)
    rng = np.random.default_rng(7)
    lengths = rng.integers(6400, 22400, size=128 * 9 + 37)
    table = DrawTable.build(lengths, cfg, 2004, noise_clip_lengths=np.full(64, 160000), num_rirs=271,
                            first_batch=3, noise_cursor=5, rir_cursor=2)
    t = spec.CLIP_SAMPLES
    starts = np.arange(65, dtype=np.int64) * 160000
    got = table.clip_records(_native.CLIP_AUG_DTYPE, starts, 64 * 160000 + 128 * t)

    recs, slot = [], 0
    for d, ncur, ridx in zip(table.batches, table.noise_clip_cursor, table.rir_index):
        b = len(d.pad_before)
        r = np.zeros(b, dtype=_native.CLIP_AUG_DTYPE)
        r["gain"] = d.gain_linear
        r["colored_index"] = slot if d.colored_apply else -1
        slot += int(d.colored_apply)
        r["colored_snr_db"] = d.colored_snr_db
        r["rir_index"] = ridx if d.reverb_apply else -1
        if d.background_apply and ncur >= 0:
            r["noise_offset"] = int(ncur) * 160000 + np.arange(b, dtype=np.int64) * t
            r["noise_snr_db"] = d.noise_snr_db
        else:
            r["noise_offset"] = -1
        recs.append(r)
    want = np.concatenate(recs)
    assert got.dtype == want.dtype and got.shape == want.shape == (len(lengths),)
    assert got.tobytes() == want.tobytes()
    assert any(d.background_apply for d in table.batches) and not all(d.background_apply for d in table.batches)
    slots, ids, fd = table.colored_slots()
    assert ids.tolist() == [3 + k for k, d in enumerate(table.batches) if d.colored_apply] and (fd == 0).all()
