"""Host-side sink bandwidth: pwrite / memmap stores of [rows,16,96] f32 blocks from pinned memory, 1..8 threads, per directory."""
import os, sys, time, tempfile, shutil
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import torch

rows, row_bytes = 4096, 16 * 96 * 4
block = torch.empty(rows * 16 * 96, dtype=torch.float32).pin_memory().numpy().reshape(rows, 16, 96)
block[:] = 1.0
n_blocks = 40
for d in [tempfile.gettempdir(), "/dev/shm", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")]:
    if not os.path.isdir(d):
        continue
    print(d, "free GB", shutil.disk_usage(d).free / 1e9, flush=True)
    for threads in (1, 2, 4, 8):
        path = os.path.join(d, "hb_io_bw.bin")
        with open(path, "wb") as fh:
            fh.truncate(n_blocks * rows * row_bytes)
        fd = os.open(path, os.O_RDWR)
        buf = memoryview(block).cast("B")
        def w(i):
            at = 0
            while at < len(buf):
                at += os.pwrite(fd, buf[at:], i * rows * row_bytes + at)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(w, range(n_blocks)))
        dt = time.perf_counter() - t0
        os.close(fd)
        print(f"  pwrite  threads={threads}: {n_blocks * rows * row_bytes / dt / 1e9:.2f} GB/s", flush=True)
        os.remove(path)
        with open(path, "wb") as fh:
            fh.truncate(n_blocks * rows * row_bytes)
        mm = np.memmap(path, dtype=np.float32, mode="r+", shape=(n_blocks * rows, 16, 96))
        def m(i):
            np.copyto(mm[i * rows:(i + 1) * rows], block)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(m, range(n_blocks)))
        dt = time.perf_counter() - t0
        print(f"  memmap  threads={threads}: {n_blocks * rows * row_bytes / dt / 1e9:.2f} GB/s", flush=True)
        del mm
        os.remove(path)
print("cpus", os.cpu_count())
