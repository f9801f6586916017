"""
Golden fixture for the appendable ``.npy`` writer, produced by the REFERENCE's own
``AppendableNumpyArrayFile`` / ``AppendableNumpyHeaderInfo`` (``util/numpy_util.py:225-564``).

Run in the build container only:   python tests/golden/make_golden_npy.py

The reference imports ``numpy.compat`` (removed from NumPy 2.3); the stub installed here provides the two names it uses
with their historical meaning: ``isfileobj`` (is this a real OS-level file object?) and ``pickle``; three private
``numpy.lib.format`` helpers it calls are re-exposed from ``numpy.lib._format_impl`` (same functions, new home).

Writes ``appendable_npy.npz``: for a fixed sequence of appends the complete file bytes after every step, the header
bytes ``ensure_appendable`` produces for a plain ``np.save`` file, and the result of ``recover`` on a torn file.
"""
from __future__ import annotations

import io
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main() -> None:
    import_reference()
    compat = sys.modules["numpy.compat"]
    compat.isfileobj = lambda f: isinstance(f, (io.FileIO, io.BufferedReader, io.BufferedWriter))
    compat.pickle = pickle
    # NumPy >= 2.3 keeps the private format helpers the reference calls in numpy.lib._format_impl: re-expose them unchanged
    import numpy.lib._format_impl as impl
    import numpy.lib.format as fmt
    for name in ("_check_version", "_header_size_info", "_read_array_header"):
        if not hasattr(fmt, name):
            setattr(fmt, name, getattr(impl, name))
    from heybuddy.util.numpy_util import AppendableNumpyArrayFile, AppendableNumpyHeaderInfo  # reference

    rng = np.random.default_rng(6001)
    parts = [rng.standard_normal((3, 16, 96)).astype(np.float32), rng.standard_normal((2, 16, 96)).astype(np.float32),
             rng.standard_normal((7, 16, 96)).astype(np.float32)]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "a.npy")
        with AppendableNumpyArrayFile(path) as f:
            for i, p in enumerate(parts):
                f.append(p)
                f.fp.flush()
                out[f"file_after_{i}"] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8).copy()
        # reopen + append (header rewritten in place at the same length)
        with AppendableNumpyArrayFile(path) as f:
            f.append(parts[0])
        out["file_reopened"] = np.frombuffer(open(path, "rb").read(), dtype=np.uint8).copy()
        assert np.load(path).shape == (15, 16, 96)

        # float16, 1-D and Fortran-order growth
        p16 = os.path.join(d, "h.npy")
        with AppendableNumpyArrayFile(p16) as f:
            f.append(parts[0].astype(np.float16))
            f.append(parts[1].astype(np.float16))
        out["file_f16"] = np.frombuffer(open(p16, "rb").read(), dtype=np.uint8).copy()
        p1d = os.path.join(d, "v.npy")
        with AppendableNumpyArrayFile(p1d) as f:
            f.append(np.arange(5, dtype=np.int64))
            f.append(np.arange(4, dtype=np.int64))
        out["file_1d"] = np.frombuffer(open(p1d, "rb").read(), dtype=np.uint8).copy()
        pf = os.path.join(d, "f.npy")
        with AppendableNumpyArrayFile(pf) as f:
            f.append(np.asfortranarray(parts[0][:, :, :4]))
            f.append(np.asfortranarray(parts[1][:3, :, :5].repeat(2, axis=0)[:3]))
        out["file_fortran"] = np.frombuffer(open(pf, "rb").read(), dtype=np.uint8).copy()

        # ensure_appendable on a plain np.save file (both strategies give the same bytes)
        plain = os.path.join(d, "p.npy")
        np.save(plain, parts[2])
        out["plain_file"] = np.frombuffer(open(plain, "rb").read(), dtype=np.uint8).copy()
        out["plain_is_appendable"] = np.array(AppendableNumpyHeaderInfo.file_is_appendable(plain))
        AppendableNumpyHeaderInfo.ensure_appendable(plain, in_place=True)
        out["plain_made_appendable"] = np.frombuffer(open(plain, "rb").read(), dtype=np.uint8).copy()

        # recover: a torn tail (half a row) is truncated / zero-filled and the header corrected
        for mode, zf in (("truncate", False), ("zerofill", True)):
            torn = os.path.join(d, f"t_{mode}.npy")
            with AppendableNumpyArrayFile(torn) as f:
                f.append(parts[0])
            with open(torn, "ab") as fh:
                fh.write(parts[1].tobytes()[: 16 * 96 * 4 + 1000])   # one full row + part of another, header not updated
            out[f"torn_{mode}_needs_recovery"] = np.array(AppendableNumpyHeaderInfo.file_needs_recovery(torn))
            AppendableNumpyHeaderInfo.recover(torn, zerofill_incomplete=zf)
            out[f"torn_{mode}_recovered"] = np.frombuffer(open(torn, "rb").read(), dtype=np.uint8).copy()
    for i, p in enumerate(parts):
        out[f"part_{i}"] = p
    np.savez_compressed(os.path.join(HERE, "appendable_npy.npz"), **out)
    print("wrote appendable_npy.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
