"""Host probe: O_DIRECT writes (page cache bypassed) into a fresh file vs buffered pwrite, 4 threads, 1.2 GB (the e2e sink question)."""
import mmap, os, sys, threading, time

SIZE, CHUNK, THREADS = 1200 << 20, 8 << 20, 4
buf = mmap.mmap(-1, CHUNK)            # page-aligned
buf.write(b"\x01" * CHUNK)


def run(path, flags):
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC | flags, 0o644)
    os.ftruncate(fd, SIZE)
    n = SIZE // CHUNK
    def work(t):
        for i in range(t, n, THREADS):
            os.pwrite(fd, buf, i * CHUNK)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(THREADS)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    os.close(fd); os.unlink(path)
    return SIZE / dt / 1e9


for d in sys.argv[1:] or ["/tmp"]:
    for name, flags in (("buffered", 0), ("O_DIRECT", os.O_DIRECT)):
        try:
            print(f"{d} {name}: {run(os.path.join(d, 'io_probe.bin'), flags):.2f} GB/s", flush=True)
        except OSError as e:
            print(f"{d} {name}: {e}", flush=True)
