"""
Appendable ``.npy`` files and the shard merge built on them.

Replaces ``AppendableNumpyArrayFile`` / ``AppendableNumpyHeaderInfo``
(reference ``util/numpy_util.py:225-564``) and the ``heybuddy combine`` command
(``__main__.py:121-169``) without ``numpy.compat`` or NumPy-private helpers.

File format = NumPy ``.npy`` v1.0 with one twist the reference introduced: after the header dict the
writer leaves ``21 - len(str(rows))`` spare spaces, so the growth-axis length can later be rewritten
**in place** (same header length, data untouched) every time rows are appended.  Such a file stays a
valid ``.npy`` at every step: ``np.load`` / ``np.load(mmap_mode="r")`` / ``PrecalculatedDatasetIterator``
read it unchanged.  Bytes produced here are identical to the reference's (``tests/golden/appendable_npy.npz``).

Why it matters on this path: the featurizer produces ``[n, 16, 96]`` chunks on the GPU; appending them
as they drain replaces the reference's whole-array ``np.concatenate`` + ``np.save`` rewrite, and
``combine_precalculated`` merges per-rank / per-chunk shards into one file the same way.
"""
from __future__ import annotations

import ast
import gc
import os
import struct
import tempfile
import threading
from math import prod
from typing import Any, BinaryIO, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["AppendableNumpyArrayFile", "AppendableNumpyHeaderInfo", "combine_precalculated", "GROWTH_AXIS_MAX_DIGITS"]

GROWTH_AXIS_MAX_DIGITS = 21          # reference numpy_util.py:22
_MAGIC = b"\x93NUMPY"
_ALIGN = 64                          # numpy.lib.format.ARRAY_ALIGN
_COPY_CHUNK = 16 * 1024 ** 2


def _descr(dtype: np.dtype) -> Any:
    return np.lib.format.dtype_to_descr(dtype)


def _header_bytes(shape: Tuple[int, ...], fortran_order: bool, dtype: np.dtype, total_len: Optional[int] = None) -> bytes:
    """
    Magic + length field + header dict (keys sorted, as NumPy writes them) + the spare digits for the growth axis
    (axis 0, or the last axis for Fortran order) + space padding to a 64-byte boundary + ``\\n``.
    ``total_len``: pad to exactly this many bytes instead (in-place rewrite); ValueError if it does not fit.
    Version 1.0 (u16 length) unless the header is too long for it, then 2.0 (u32 length), like ``np.save``.
    """
    body = "{'descr': %r, 'fortran_order': %r, 'shape': %r, }" % (_descr(dtype), bool(fortran_order), tuple(int(s) for s in shape))
    if len(shape) > 0:
        growth = shape[-1] if fortran_order else shape[0]
        body += " " * (GROWTH_AXIS_MAX_DIGITS - len(repr(int(growth))))
    text = body.encode("latin1")
    for version, fmt in (((1, 0), "<H"), ((2, 0), "<I")):
        prefix_len = len(_MAGIC) + 2 + struct.calcsize(fmt)
        hlen = len(text) + 1                                       # + the final newline
        pad = _ALIGN - ((prefix_len + hlen) % _ALIGN)              # a full 64 when already aligned, as NumPy does
        try:
            prefix = _MAGIC + bytes(version) + struct.pack(fmt, hlen + pad)
        except struct.error:
            continue
        out = prefix + text + b" " * pad
        if total_len is not None:
            if len(out) + 1 > total_len:
                raise ValueError(f"header needs {len(out) + 1} bytes, only {total_len} available")
            out += b" " * (total_len - len(out) - 1)
        return out + b"\n"
    raise ValueError("header too long for the .npy format")


def _read_header(fp: BinaryIO) -> Tuple[Tuple[int, ...], bool, np.dtype, int]:
    """(shape, fortran_order, dtype, header size in bytes); the file position ends right after the header."""
    fp.seek(0)
    head = fp.read(len(_MAGIC) + 2)
    if len(head) < 8 or head[:6] != _MAGIC:
        raise ValueError("not a .npy file")
    major = head[6]
    fmt = "<H" if major == 1 else "<I"
    (hlen,) = struct.unpack(fmt, fp.read(struct.calcsize(fmt)))
    text = fp.read(hlen).decode("latin1" if major < 3 else "utf8")
    d = ast.literal_eval(text)
    if not isinstance(d, dict) or set(d) != {"descr", "fortran_order", "shape"}:
        raise ValueError("malformed .npy header")
    return tuple(int(s) for s in d["shape"]), bool(d["fortran_order"]), np.lib.format.descr_to_dtype(d["descr"]), fp.tell()


class AppendableNumpyHeaderInfo:
    """What the header of an existing ``.npy`` says, and whether rows can be appended to the file in place."""

    def __init__(self, fp: BinaryIO) -> None:
        self.shape, self.fortran_order, self.dtype, self.header_size = _read_header(fp)
        self.new_header = _header_bytes(self.shape, self.fortran_order, self.dtype)
        fp.seek(0, os.SEEK_END)
        self.data_length = fp.tell() - self.header_size
        self.is_appendable = len(self.new_header) <= self.header_size
        self.needs_recovery = not (self.dtype.hasobject or self.data_length == prod(self.shape) * self.dtype.itemsize)

    @classmethod
    def file_is_appendable(cls, filename: str) -> bool:
        with open(filename, "rb") as fp:
            return cls(fp).is_appendable

    @classmethod
    def file_needs_recovery(cls, filename: str) -> bool:
        with open(filename, "rb") as fp:
            return cls(fp).needs_recovery

    @classmethod
    def ensure_appendable(cls, filename: str, in_place: bool = False) -> None:
        """Give a plain ``np.save`` file a header with spare digits (moves the data if the header has to grow)."""
        with open(filename, "rb+") as fp:
            info = cls(fp)
            if info.is_appendable:
                return
            new_size, old_size, n = len(info.new_header), info.header_size, info.data_length
            chunk = max(1, min(_COPY_CHUNK, n))
            if in_place:
                # shift the data towards the end, last block first, so nothing is overwritten before it is read
                for i in reversed(range((n + chunk - 1) // chunk)):
                    fp.seek(old_size + i * chunk)
                    block = fp.read(chunk)
                    fp.seek(new_size + i * chunk)
                    fp.write(block)
                fp.seek(0)
                fp.write(info.new_header)
                return
            directory, base = os.path.split(os.path.abspath(filename))
            tmp = tempfile.NamedTemporaryFile(prefix=base, dir=directory, delete=False)
            with tmp:
                tmp.write(info.new_header)
                fp.seek(old_size)
                while True:
                    block = fp.read(chunk)
                    if not block:
                        break
                    tmp.write(block)
        os.replace(tmp.name, filename)

    @classmethod
    def recover(cls, filename: str, zerofill_incomplete: bool = False) -> None:
        """
        After a crash between a data append and its header rewrite: make header and data agree again.  A torn last row is
        truncated (default) or zero-filled.
        """
        with open(filename, "rb+") as fp:
            info = cls(fp)
            if not info.needs_recovery:
                return
            if not info.is_appendable:
                raise ValueError("header not appendable, call ensure_appendable first")
            inner = info.shape[:-1] if info.fortran_order else info.shape[1:]
            row_bytes = prod(inner) * info.dtype.itemsize
            n = info.data_length
            torn = n % row_bytes
            if torn:
                if zerofill_incomplete:
                    fp.seek(0, os.SEEK_END)
                    fp.write(b"\0" * (row_bytes - torn))
                    n += row_bytes - torn
                else:
                    fp.truncate(info.header_size + n - torn)
                    n -= torn
            rows = n // row_bytes
            shape = (*inner, rows) if info.fortran_order else (rows, *inner)
            fp.seek(0)
            fp.write(_header_bytes(shape, info.fortran_order, info.dtype, info.header_size))


class AppendableNumpyArrayFile:
    """
    Append arrays along the growth axis of a ``.npy`` file.

    >>> with AppendableNumpyArrayFile(path) as f:
    ...     f.append(np.array([1, 2, 3]))
    ...     f.append(np.array([4, 5, 6]))
    >>> np.load(path)
    array([1, 2, 3, 4, 5, 6])

    Same constructor and error behaviour as the reference class: an existing non-empty file is opened for appending (it must
    be appendable and consistent), an empty one or ``delete_if_exists`` starts over; the first ``append`` fixes dtype, memory
    order and the non-growth dimensions.  ``rewrite_header_on_append=False`` defers the header rewrite to ``close()``.
    """

    def __init__(self, filename: str, delete_if_exists: bool = False, rewrite_header_on_append: bool = True) -> None:
        self.filename = filename
        self.rewrite_header_on_append = rewrite_header_on_append
        self.lock = threading.Lock()
        self.initialized = False
        self.fp: Optional[BinaryIO] = None
        self.header_length: Optional[int] = None
        if os.path.exists(filename):
            if os.path.getsize(filename) == 0 or delete_if_exists:
                os.unlink(filename)
            else:
                self.initialize_file()

    def initialize_file(self) -> None:
        self.fp = open(self.filename, "rb+")
        info = AppendableNumpyHeaderInfo(self.fp)
        self.shape, self.fortran_order, self.dtype, self.header_length = info.shape, info.fortran_order, info.dtype, info.header_size
        if self.dtype.hasobject:
            raise ValueError("Object arrays cannot be appended to")
        if not info.is_appendable:
            raise ValueError(f"Header of {self.filename} not appendable. Call `AppendableNumpyHeaderInfo.ensure_appendable`")
        if info.needs_recovery:
            raise ValueError(f"Cannot append to {self.filename}, needs recovery. Call `AppendableNumpyHeaderInfo.recover`")
        self.initialized = True

    def _write_array_header(self) -> None:
        if self.fp is None:
            return
        self.fp.seek(0)
        self.fp.write(_header_bytes(self.shape, self.fortran_order, self.dtype, self.header_length))

    def update_header(self) -> None:
        with self.lock:
            self._write_array_header()

    def append(self, arr: np.ndarray) -> None:
        arr = np.asanyarray(arr)
        with self.lock:
            if not self.initialized:
                if arr.dtype.hasobject:
                    raise ValueError("Object arrays cannot be appended to")
                fortran = bool(arr.flags.f_contiguous and not arr.flags.c_contiguous)
                with open(self.filename, "wb") as fp:
                    fp.write(_header_bytes(arr.shape, fortran, arr.dtype))
                    (arr.T if fortran else np.ascontiguousarray(arr)).tofile(fp)
                return self.initialize_file()
            inner_now = self.shape[:-1] if self.fortran_order else self.shape[1:]
            inner_new = arr.shape[:-1] if self.fortran_order else arr.shape[1:]
            if inner_now != inner_new:
                raise ValueError(f"Shapes {inner_now} and {inner_new} do not match")
            assert self.fp is not None
            self.fp.seek(0, os.SEEK_END)
            arr.astype(self.dtype, copy=False).flatten(order="F" if self.fortran_order else "C").tofile(self.fp)
            if self.fortran_order:
                self.shape = (*self.shape[:-1], self.shape[-1] + arr.shape[-1])
            else:
                self.shape = (self.shape[0] + arr.shape[0], *self.shape[1:])
            if self.rewrite_header_on_append:
                self._write_array_header()

    def close(self) -> None:
        with self.lock:
            if self.initialized:
                if not self.rewrite_header_on_append:
                    self._write_array_header()
                assert self.fp is not None
                self.fp.close()
                self.initialized = False

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "AppendableNumpyArrayFile":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


def combine_precalculated(source: Sequence[str], target: str, directory: str, reset: bool = True, half: bool = False,
                          delete: bool = False, batch_size: int = 10) -> Tuple[int, ...]:
    """
    ``heybuddy combine`` (reference ``__main__.py:121-169``): every ``*.npy`` under ``directory/<name>`` for ``name`` in
    ``source``, in sorted path order, appended along axis 0 into ``directory/target`` -- ``batch_size`` files per append,
    optionally cast to float16, optionally deleting the shards (and their then-empty directories).  This is also how
    per-rank feature shards (``<name>/<rank>.npy``) become one file.  Returns the final shape.
    """
    target_path = os.path.join(directory, target)
    if os.path.exists(target_path) and reset:
        os.remove(target_path)
    directories = [os.path.join(directory, name) for name in source]
    files: List[str] = []
    for d in directories:
        files.extend(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".npy"))
    with AppendableNumpyArrayFile(target_path) as out:
        batch: List[np.ndarray] = []
        for filename in sorted(files):
            data = np.load(filename)
            batch.append(data.astype(np.float16) if half else data)
            if len(batch) % batch_size == 0:
                out.append(np.concatenate(batch, axis=0))
                batch = []
                gc.collect()
            if delete:
                os.remove(filename)
        if batch:
            out.append(np.concatenate(batch, axis=0))
        shape = tuple(out.shape) if out.initialized else ()
    if delete:
        for d in directories:
            os.rmdir(d)
    return shape
