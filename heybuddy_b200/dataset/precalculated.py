"""
``PrecalculatedDatasetIterator`` -- the ``.npy`` memmap store of precomputed ``[N, 16, 96]`` embeddings
(reference ``heybuddy/dataset/precalculated.py:365-574``).

Same on-disk format (``np.save`` v1.0, C order, ``<f4``; ``np.load(mmap_mode="r")``), same ``take(n)`` with a
shuffled index and wrap-around reshuffle, ``from_array``, ``metadata``, ``__len__``.  Fixed on purpose:
``from_array(..., directory=X)`` re-opens from X (the reference re-opens from the default directory,
precalculated.py:482-491), index bookkeeping is guarded by the lock the reference declares but never takes
(``take`` is called from many batcher threads), shuffling is seedable.  The hosted multi-GB negative sets and
the BERT-token ``exclude_phrase`` filter need network downloads and are outside the hot path.
"""
from __future__ import annotations

import os
from threading import Lock
from typing import Any, Dict, Iterator, Optional

import numpy as np

__all__ = ["PrecalculatedDatasetIterator", "LOCAL_DIR", "open_shared_memmap", "NpyRowWriter"]

LOCAL_DIR = os.environ.get(
    "HEYBUDDY_B200_PRECALCULATED_DIR",
    os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "precalculated")),
)


def open_shared_memmap(path: str, shape, rank: int = 0, barrier=None, dtype=np.float32) -> np.memmap:
    """
    A pre-sized ``.npy`` every rank writes its own row range into: rank 0 creates it
    (``np.lib.format.open_memmap(mode="w+")``), everybody else opens it ``r+`` after the barrier.
    """
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        mm = np.lib.format.open_memmap(path, mode="w+", dtype=dtype, shape=tuple(shape))
        mm.flush()
    if barrier is not None:
        barrier()
    if rank != 0:
        mm = np.lib.format.open_memmap(path, mode="r+")
        assert tuple(mm.shape) == tuple(shape), (mm.shape, shape)
    return mm


def _fs_type(path: str) -> str:
    """File-system type of the mount that holds ``path`` (longest matching mount point in /proc/mounts)."""
    try:
        real = os.path.realpath(path)
        best, kind = "", ""
        with open("/proc/mounts") as fh:
            for line in fh:
                parts = line.split()
                if len(parts) >= 3 and (real == parts[1] or real.startswith(parts[1].rstrip("/") + "/")) and len(parts[1]) > len(best):
                    best, kind = parts[1], parts[2]
        return kind
    except Exception:
        return ""


class NpyRowWriter:
    """
    Row-range writer of one pre-sized ``.npy`` shared by all ranks: the creating rank writes the exact header ``np.save`` would
    write (v1.0, C order -- the file is byte-identical to ``np.save(array)``, precalculated.py:486) and sizes the file; every
    rank then ``pwrite``s its own rows at their byte offset.  ``pwrite`` copies straight from the caller's buffer (the
    pipeline's pinned D2H slots) into the page cache with the GIL released, so several worker threads write in parallel and no
    page of the file is ever faulted into this process the way a memmap store would.
    """

    def __init__(self, path: str, shape, create: bool = True, barrier=None, dtype=np.float32, shared: bool = False) -> None:
        self.path, self.shape, self.dtype = path, tuple(int(x) for x in shape), np.dtype(dtype)
        self.row_bytes = int(np.prod(self.shape[1:])) * self.dtype.itemsize
        if create:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            tmp = path + ".partial"
            with open(tmp, "wb") as fh:
                np.lib.format.write_array_header_1_0(fh, {"descr": np.lib.format.dtype_to_descr(self.dtype), "fortran_order": False,
                                                          "shape": self.shape})
                header = fh.tell()
                fh.truncate(header + self.shape[0] * self.row_bytes)
            os.replace(tmp, path)
        if barrier is not None:
            barrier()
        self.fd = os.open(path, os.O_RDWR)
        with open(path, "rb") as fh:
            version = np.lib.format.read_magic(fh)
            got_shape, fortran, got_dtype = np.lib.format.read_array_header_1_0(fh) if version == (1, 0) else np.lib.format.read_array_header_2_0(fh)
            self.data_offset = fh.tell()
        assert tuple(got_shape) == self.shape and not fortran and got_dtype == self.dtype, (got_shape, self.shape, got_dtype)
        self._mapped = []
        # one writer process: positional writes were the faster way into the page cache (10.8 vs 16.0 ms per 8192-clip step);
        # several ranks sharing the file: pwrite serialises on the inode lock (27.8 ms at two ranks), stores through a shared
        # mapping do not (21.2 ms) -- measured on the round-2 bench boxes, profiles/README.md "e2e sink"
        self.mode = os.environ.get("HEYBUDDY_B200_SINK") or ("mmap" if shared else "pwrite")
        assert self.mode in ("mmap", "pwrite"), self.mode
        self._mm = None
        self._mm_lock = Lock()

    def write(self, row: int, rows: np.ndarray) -> None:
        """
        Rows ``[row, row + len(rows))`` <- ``rows`` (C-contiguous, the file's dtype).  Thread-safe.  Two ways into the page cache
        (``HEYBUDDY_B200_SINK`` overrides the choice made in ``__init__``): ``pwrite`` issues positional writes, which Linux file
        systems serialise on the inode's write lock; ``mmap`` stores through a shared mapping -- a page fault per new page, no
        per-file lock, so the writer threads of several ranks fill one file side by side.
        """
        assert rows.dtype == self.dtype and rows.flags.c_contiguous and rows.shape[1:] == self.shape[1:]
        assert 0 <= row and row + rows.shape[0] <= self.shape[0]
        if self.mode == "mmap":
            np.copyto(self._memmap()[row:row + rows.shape[0]], rows)
            return
        buf = memoryview(rows).cast("B")
        at, off = 0, self.data_offset + row * self.row_bytes
        while at < len(buf):
            at += os.pwrite(self.fd, buf[at:], off + at)

    def _memmap(self) -> np.ndarray:
        if self._mm is None:
            with self._mm_lock:
                if self._mm is None:
                    self._mm = np.memmap(self.path, dtype=self.dtype, mode="r+", offset=self.data_offset, shape=self.shape)
        return self._mm

    def map_pinned(self, row_lo: int, row_hi: int, populate_threads: int = 4):
        """
        Rows ``[row_lo, row_hi)`` of the file as a PINNED f32 torch tensor, or None when the file system does not allow it: the
        file is mapped shared, the range's pages are faulted in by ``populate_threads`` threads and the mapping is registered
        with the CUDA driver (``cudaHostRegister``), so a device -> host copy of finished rows lands IN the file's page cache
        with no staging buffer and no CPU copy.  Works on memory-backed file systems (tmpfs, /dev/shm); on disk-backed ones the
        driver (or the kernel's long-term-pin rule for file mappings) refuses and the caller falls back to :meth:`write`.

        OPT-IN (``HEYBUDDY_B200_PINNED_FILE=1``): creating the file's pages is what bounds either sink, and on the round-2 bench
        boxes (16-vCPU VMs) faulting them in ahead of the stream ran at ~2 GB/s -- 34-47 ms per 8192-clip step against 11-13 ms
        for ``pwrite`` from worker threads that overlaps the kernels (profiles/README.md, "e2e sink").
        """
        import mmap

        import torch

        if os.environ.get("HEYBUDDY_B200_PINNED_FILE", "") not in ("1", "true") or row_hi <= row_lo:
            return None
        if _fs_type(self.path) not in ("tmpfs", "shm", "ramfs", "devtmpfs"):
            return None      # DMA into disk-backed page-cache pages would bypass the kernel's dirty tracking
        try:
            size = os.fstat(self.fd).st_size
            mm = mmap.mmap(self.fd, size, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ | mmap.PROT_WRITE)
            whole = np.frombuffer(mm, dtype=np.uint8)
            page = mmap.PAGESIZE
            b0 = (self.data_offset + row_lo * self.row_bytes) // page * page
            b1 = min(size, -(-(self.data_offset + row_hi * self.row_bytes) // page) * page)
            # fault the pages in from several threads (a page fault on a memory-backed file takes no inode lock)
            cuts = np.linspace(b0, b1, populate_threads + 1).astype(np.int64) // page * page
            cuts[0], cuts[-1] = b0, b1

            def touch(i: int) -> None:
                lo, hi = int(cuts[i]), int(cuts[i + 1])
                if hi > lo:
                    try:
                        mm.madvise(23, lo, hi - lo)          # MADV_POPULATE_WRITE (Linux >= 5.14)
                    except (OSError, ValueError):
                        view = whole[lo:hi:page]
                        np.bitwise_or(view, 0, out=view)      # one read-modify-write per page

            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(populate_threads) as ex:
                list(ex.map(touch, range(populate_threads)))
            addr = whole.ctypes.data + b0
            rc = torch.cuda.cudart().cudaHostRegister(addr, b1 - b0, 0)
            if int(rc) != 0:
                del whole
                mm.close()
                return None
            at = self.data_offset + row_lo * self.row_bytes
            rows = torch.frombuffer(mm, dtype=torch.float32, count=(row_hi - row_lo) * (self.row_bytes // 4), offset=at)
            rows = rows.view(row_hi - row_lo, *self.shape[1:])
            if not rows.is_pinned():
                torch.cuda.cudart().cudaHostUnregister(addr)
                del rows, whole
                mm.close()
                return None
            self._mapped.append((mm, whole, addr))
            return rows
        except Exception:
            return None

    def close(self) -> None:
        import gc

        for mm, whole, addr in self._mapped:
            try:
                import torch

                torch.cuda.cudart().cudaHostUnregister(addr)
            except Exception:
                pass
            del whole
            gc.collect()
            try:
                mm.close()
            except BufferError:
                pass          # a tensor view is still alive: the mapping goes away with it
        self._mapped = []
        if self._mm is not None:
            del self._mm           # dirty pages stay in the page cache like after a write(); no msync, np.save does none either
            self._mm = None
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None

    def __enter__(self) -> "NpyRowWriter":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


class PrecalculatedDatasetIterator:
    """An extensible dataset iterator over precalculated features."""

    def __init__(self, name: str, directory: Optional[str] = None, exclude_phrase: Optional[str] = None, ordered: bool = False,
                 labeled: bool = False, use_mem_map: bool = True, shuffle: bool = True, data: Optional[np.ndarray] = None,
                 seed: Optional[int] = None) -> None:
        self.lock = Lock()
        self.directory = directory or LOCAL_DIR
        self.name = name
        if exclude_phrase is not None:
            raise NotImplementedError("exclude_phrase needs the hub BERT tokenizer (outside the hot path)")
        self.exclude_phrase = exclude_phrase
        self.index = 0
        self.total_taken = 0
        self.ordered = ordered
        self.labeled = labeled
        self.use_mem_map = use_mem_map
        self._rng = np.random.default_rng(seed)
        if data is not None:
            self._precalculated = data
        if not os.path.exists(self.precalculated_path):
            raise FileNotFoundError(f"Could not find precalculated features at {self.precalculated_path}.")
        if shuffle and not ordered:
            self.shuffle()

    @property
    def precalculated_path(self) -> str:
        return os.path.join(self.directory, f"{self.name}.npy")

    @property
    def precalculated(self) -> np.ndarray:
        if not hasattr(self, "_precalculated"):
            self._precalculated = np.load(self.precalculated_path, **({"mmap_mode": "r"} if self.use_mem_map else {}))
        return self._precalculated

    @property
    def indexes(self) -> np.ndarray:
        if not hasattr(self, "_indexes"):
            self._indexes = np.arange(len(self.precalculated))
        return self._indexes

    @classmethod
    def from_array(cls, array: np.ndarray, name: str, directory: Optional[str] = None, ordered: bool = False,
                   keep_in_memory: bool = False, **kwargs: Any) -> "PrecalculatedDatasetIterator":
        """Saves the features to ``<directory>/<name>.npy`` and opens them."""
        directory = directory or LOCAL_DIR
        os.makedirs(directory, exist_ok=True)
        np.save(os.path.join(directory, f"{name}.npy"), array)
        return cls(name, directory=directory, data=array if keep_in_memory else None, ordered=ordered, **kwargs)

    def shuffle(self) -> "PrecalculatedDatasetIterator":
        if not self.ordered:
            self._rng.shuffle(self.indexes)
        return self

    def take(self, n: int) -> np.ndarray:
        """The next ``n`` samples -> ``[n, 16, 96]``; wraps around with a reshuffle (precalculated.py:501-536)."""
        with self.lock:
            batch = self.precalculated[self.indexes[self.index:self.index + n]]
            if batch.shape[0] < n:
                self.index = n - batch.shape[0]
                self.shuffle()
                batch = np.concatenate([batch, self.precalculated[self.indexes[:self.index]]])
            else:
                self.index += n
            if self.labeled:
                batch = batch[:, :-1]
            self.total_taken += n
        return np.asarray(batch)

    def iterate(self) -> Iterator[np.ndarray]:
        while True:
            yield self.take(1)

    def metadata(self) -> Dict[str, Any]:
        return {"name": self.name, "path": self.precalculated_path, "shape": self.precalculated.shape,
                "ordered": self.ordered, "labeled": self.labeled, "use_mem_map": self.use_mem_map}

    def __len__(self) -> int:
        return int(self.precalculated.shape[0])

    def __iter__(self) -> Iterator[np.ndarray]:
        return self.iterate()

    def __repr__(self) -> str:
        return f"{type(self).__name__}(num_samples={len(self)})"
