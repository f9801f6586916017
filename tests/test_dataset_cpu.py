"""CPU: .npy store, batch iterator, file naming and rank sharding (host logic of SURVEY.md 8 rows a12-a15, 8e)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from heybuddy_b200 import spec
from heybuddy_b200.dataset.features import SyntheticSpeechSource, TrainingFeaturesGenerator, shard_batches
from heybuddy_b200.dataset.precalculated import PrecalculatedDatasetIterator, open_shared_memmap
from heybuddy_b200.dataset.training import WakeWordTrainingDatasetIterator

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_npy_store_matches_reference_fixture(tmp_path, golden_dir):
    """Same header bytes, same ordered take() with wrap-around as the reference's own iterator (golden/precalculated.npz)."""
    g = np.load(os.path.join(golden_dir, "precalculated.npz"))
    arr = np.arange(7 * 16 * 96, dtype=np.float32).reshape(7, 16, 96)
    it = PrecalculatedDatasetIterator.from_array(arr, "golden_tmp", directory=str(tmp_path), ordered=True)
    with open(tmp_path / "golden_tmp.npy", "rb") as fh:
        assert np.array_equal(np.frombuffer(fh.read(128), dtype=np.uint8), g["header"])
    for k in ("take0", "take1", "take2"):
        got = it.take(3)
        assert got.shape == (3, 16, 96) and got.dtype == np.float32
        np.testing.assert_array_equal(got[:, 0, 0], g[k])
    assert len(it) == 7 and it.total_taken == 9 and isinstance(it.precalculated, np.memmap)
    # batch layout: positives (label 1) then negatives (label 0), x f32 / y i64 (training.py:254-262)
    neg = PrecalculatedDatasetIterator("golden_tmp", directory=str(tmp_path), ordered=True)
    pos = PrecalculatedDatasetIterator("golden_tmp", directory=str(tmp_path), ordered=True)
    tr = WakeWordTrainingDatasetIterator(positive=[(pos, 2)], negative=[(neg, 3)], num_batch_threads=1, start=False)
    x, y = tr.make_batch()
    assert tuple(x.shape) == tuple(g["batch_x_shape"]) and str(x.dtype) == str(g["batch_x_dtype"]) and str(y.dtype) == str(g["batch_y_dtype"])
    np.testing.assert_array_equal(y.numpy(), g["batch_y"])
    tr.start()
    batches = [b for _, b in zip(range(5), tr)]
    tr.stop()
    assert len(batches) == 5 and all(tuple(b[0].shape) == (5, 16, 96) for b in batches)
    tr.multiply_batch_size(2)
    assert [n for _, n in tr.positive + tr.negative] == [4, 6]
    with pytest.raises(FileNotFoundError):
        PrecalculatedDatasetIterator("missing", directory=str(tmp_path))


def test_shuffled_take_covers_everything(tmp_path):
    arr = np.arange(10, dtype=np.float32)[:, None, None] * np.ones((1, 16, 96), np.float32)
    it = PrecalculatedDatasetIterator.from_array(arr, "s", directory=str(tmp_path), seed=0)
    seen = np.concatenate([it.take(5)[:, 0, 0], it.take(5)[:, 0, 0]])
    assert sorted(seen.tolist()) == list(range(10)) and seen.tolist() != list(range(10))
    assert it.take(4).shape == (4, 16, 96)  # wraps with a reshuffle


def test_file_naming_and_stale_kwargs():
    assert TrainingFeaturesGenerator.get_wake_phrase_file_name("Hey, Buddy!") == "hey_buddy"
    assert TrainingFeaturesGenerator.get_wake_phrase_file_name("hello_world", testing=True) == "hello_world_tst"
    # stale spellings of the reference's own tests are accepted (tests/test_feature_generator.py:17-24)
    g = TrainingFeaturesGenerator(device="cuda", sample_batch_size=5000, tts_batch_size=64, tts_num_threads=2,
                                  augment_batch_size=128, augment_num_threads=2)
    assert g.device_id == 0 and g.augment_batch_size == 128
    from heybuddy_b200.dataset import TrainingDatasetGenerator  # stale class name
    assert TrainingDatasetGenerator is WakeWordTrainingDatasetIterator


def test_synthetic_source_is_row_addressable():
    s = SyntheticSpeechSource(seed=5)
    a = s(4, start=10)
    b = s(2, start=12)
    assert all(c.dtype == np.int16 and 6400 <= c.shape[0] < 22400 and np.abs(c).max() == 32767 for c in a)
    np.testing.assert_array_equal(a[2], b[0])


def test_shard_batches_partition():
    for n, w in ((782, 8), (10, 4), (3, 8), (0, 2)):
        ranges = [shard_batches(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
        assert max(hi - lo for lo, hi in ranges) - min(hi - lo for lo, hi in ranges) <= 1


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, sys.argv[1])
    from heybuddy_b200.dataset.features import shard_batches
    from heybuddy_b200.dataset.precalculated import open_shared_memmap
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, b = 1000, 128
    path = sys.argv[2]
    lo_b, hi_b = shard_batches((n + b - 1) // b, rank, world)
    lo, hi = lo_b * b, min(hi_b * b, n)
    mm = open_shared_memmap(path, (n, 16, 96), rank, dist.barrier)
    rows = np.arange(lo, hi, dtype=np.float32)
    mm[lo:hi] = rows[:, None, None] + np.arange(16, dtype=np.float32)[None, :, None] * 0.001
    mm.flush()
    dist.barrier()
    if rank == 0:
        full = np.load(path, mmap_mode="r")
        assert full.shape == (n, 16, 96)
        assert np.array_equal(full[:, 0, 0], np.arange(n, dtype=np.float32))
        print("OK", lo, hi)
    dist.destroy_process_group()
""")


def test_two_rank_sharded_memmap_gloo(tmp_path):
    """world_size 2 over gloo: each rank writes its own row range of one .npy; no data-path collective."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "shared.npy"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", str(script), ROOT, str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "OK 0 512" in res.stdout
    it = PrecalculatedDatasetIterator("shared", directory=str(tmp_path), ordered=True)
    assert len(it) == 1000 and it.take(2)[1, 0, 0] == 1.0
