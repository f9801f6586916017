"""CPU: appendable .npy writer + shard merge vs bytes produced by the reference's own classes (tests/golden/appendable_npy.npz)."""
import os

import numpy as np
import pytest

from heybuddy_b200.util.npy_append import AppendableNumpyArrayFile, AppendableNumpyHeaderInfo, combine_precalculated

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "appendable_npy.npz"))
PARTS = [GOLD["part_0"], GOLD["part_1"], GOLD["part_2"]]


def _bytes(path):
    with open(path, "rb") as fh:
        return np.frombuffer(fh.read(), dtype=np.uint8)


def test_append_sequence_is_byte_identical_to_reference(tmp_path):
    path = str(tmp_path / "a.npy")
    with AppendableNumpyArrayFile(path) as f:
        for i, p in enumerate(PARTS):
            f.append(p)
            f.fp.flush()
            assert np.array_equal(_bytes(path), GOLD[f"file_after_{i}"]), f"after append {i}"
            assert np.load(path).shape[0] == sum(q.shape[0] for q in PARTS[:i + 1])     # a valid .npy at every step
    with AppendableNumpyArrayFile(path) as f:       # reopen: header rewritten in place at the same length
        f.append(PARTS[0])
    assert np.array_equal(_bytes(path), GOLD["file_reopened"])
    got = np.load(path, mmap_mode="r")
    assert got.shape == (15, 16, 96) and np.array_equal(got[:12], np.concatenate(PARTS)) and np.array_equal(got[12:], PARTS[0])


def test_dtypes_orders_and_1d(tmp_path):
    p16 = str(tmp_path / "h.npy")
    with AppendableNumpyArrayFile(p16) as f:
        f.append(PARTS[0].astype(np.float16))
        f.append(PARTS[1].astype(np.float16))
    assert np.array_equal(_bytes(p16), GOLD["file_f16"])
    p1d = str(tmp_path / "v.npy")
    with AppendableNumpyArrayFile(p1d) as f:
        f.append(np.arange(5, dtype=np.int64))
        f.append(np.arange(4, dtype=np.int64))
    assert np.array_equal(_bytes(p1d), GOLD["file_1d"])
    # Fortran-order files grow along the LAST axis
    pf = str(tmp_path / "f.npy")
    gold = GOLD["file_fortran"]
    import io
    want = np.load(io.BytesIO(gold.tobytes()))
    with AppendableNumpyArrayFile(pf) as f:
        f.append(np.asfortranarray(want[:, :, :4]))
        f.append(np.asfortranarray(want[:, :, 4:]))
    assert np.array_equal(_bytes(pf), gold)
    assert np.array_equal(np.load(pf), want) and want.shape == (3, 16, 9)


def test_shape_and_object_errors(tmp_path):
    path = str(tmp_path / "e.npy")
    with AppendableNumpyArrayFile(path) as f:
        f.append(PARTS[0])
        with pytest.raises(ValueError, match="do not match"):
            f.append(np.zeros((2, 16, 95), np.float32))
        f.append(PARTS[1].astype(np.float64))        # cast to the file's dtype, like the reference
    assert np.load(path).dtype == np.float32 and np.load(path).shape == (5, 16, 96)
    with pytest.raises(ValueError, match="Object arrays"):
        AppendableNumpyArrayFile(str(tmp_path / "o.npy")).append(np.array([{}], dtype=object))
    # empty existing file = start over; delete_if_exists = start over
    empty = str(tmp_path / "z.npy")
    open(empty, "wb").close()
    with AppendableNumpyArrayFile(empty) as f:
        f.append(PARTS[1])
    assert np.load(empty).shape == (2, 16, 96)
    with AppendableNumpyArrayFile(empty, delete_if_exists=True) as f:
        f.append(PARTS[0])
    assert np.load(empty).shape == (3, 16, 96)


def test_deferred_header(tmp_path):
    path = str(tmp_path / "d.npy")
    f = AppendableNumpyArrayFile(path, rewrite_header_on_append=False)
    f.append(PARTS[0])
    f.append(PARTS[1])
    f.fp.flush()
    assert AppendableNumpyHeaderInfo.file_needs_recovery(path)      # header still says 3 rows
    f.close()
    assert not AppendableNumpyHeaderInfo.file_needs_recovery(path) and np.load(path).shape == (5, 16, 96)


@pytest.mark.parametrize("in_place", [True, False])
def test_ensure_appendable_on_plain_save(tmp_path, in_place):
    path = str(tmp_path / "p.npy")
    np.save(path, PARTS[2])
    assert np.array_equal(_bytes(path), GOLD["plain_file"])
    assert AppendableNumpyHeaderInfo.file_is_appendable(path) == bool(GOLD["plain_is_appendable"])
    AppendableNumpyHeaderInfo.ensure_appendable(path, in_place=in_place)
    assert np.array_equal(_bytes(path), GOLD["plain_made_appendable"])
    with AppendableNumpyArrayFile(path) as f:
        f.append(PARTS[0])
    assert np.array_equal(np.load(path), np.concatenate([PARTS[2], PARTS[0]]))


def test_ensure_appendable_moves_data_when_header_grows(tmp_path):
    # a header that ends exactly on the alignment boundary has no room for the spare digits: the data has to move
    path = str(tmp_path / "g.npy")
    arr = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    for in_place in (True, False):
        with open(path, "wb") as fh:
            head = b"{'descr': '<f4', 'fortran_order': False, 'shape': (2, 3, 4), }"
            pad = 64 - (10 + len(head) + 1) % 64
            fh.write(b"\x93NUMPY\x01\x00" + (len(head) + pad % 64 + 1).to_bytes(2, "little") + head + b" " * (pad % 64) + b"\n")
            arr.tofile(fh)
        assert np.array_equal(np.load(path), arr)
        if AppendableNumpyHeaderInfo.file_is_appendable(path):
            pytest.skip("header already roomy on this numpy")
        AppendableNumpyHeaderInfo.ensure_appendable(path, in_place=in_place)
        assert AppendableNumpyHeaderInfo.file_is_appendable(path) and np.array_equal(np.load(path), arr)


@pytest.mark.parametrize("mode", ["truncate", "zerofill"])
def test_recover_torn_file(tmp_path, mode):
    path = str(tmp_path / "t.npy")
    with AppendableNumpyArrayFile(path) as f:
        f.append(PARTS[0])
    with open(path, "ab") as fh:
        fh.write(PARTS[1].tobytes()[: 16 * 96 * 4 + 1000])
    assert AppendableNumpyHeaderInfo.file_needs_recovery(path) == bool(GOLD[f"torn_{mode}_needs_recovery"])
    with pytest.raises(ValueError, match="needs recovery"):
        AppendableNumpyArrayFile(path)
    AppendableNumpyHeaderInfo.recover(path, zerofill_incomplete=(mode == "zerofill"))
    assert np.array_equal(_bytes(path), GOLD[f"torn_{mode}_recovered"])
    assert np.load(path).shape[0] == (4 if mode == "truncate" else 5)


def test_combine_shards(tmp_path):
    rng = np.random.default_rng(5)
    shards = {}
    for name, n_files in (("pos", 3), ("neg", 2)):
        os.makedirs(tmp_path / name)
        for i in range(n_files):
            a = rng.standard_normal((int(rng.integers(1, 6)), 16, 96)).astype(np.float32)
            np.save(tmp_path / name / f"{i:03d}.npy", a)
            shards[str(tmp_path / name / f"{i:03d}.npy")] = a
        (tmp_path / name / "notes.txt").write_text("ignored")
    want = np.concatenate([shards[k] for k in sorted(shards)])
    shape = combine_precalculated(["pos", "neg"], "all.npy", str(tmp_path), batch_size=2)
    assert shape == want.shape and np.array_equal(np.load(tmp_path / "all.npy"), want)
    # half precision + delete: shards and (now empty) directories go away
    for name in ("pos", "neg"):
        os.remove(tmp_path / name / "notes.txt")
    shape = combine_precalculated(["pos", "neg"], "all16.npy", str(tmp_path), half=True, delete=True, batch_size=10)
    got = np.load(tmp_path / "all16.npy")
    assert got.dtype == np.float16 and np.array_equal(got, want.astype(np.float16))
    assert not (tmp_path / "pos").exists() and not (tmp_path / "neg").exists()
    # reset=False appends to the existing target
    os.makedirs(tmp_path / "more")
    np.save(tmp_path / "more" / "0.npy", want[:3])
    combine_precalculated(["more"], "all.npy", str(tmp_path), reset=False)
    assert np.load(tmp_path / "all.npy").shape[0] == want.shape[0] + 3
