"""Is DMA straight out of the page cache cheaper than a copy into pinned memory (tools only)?  Per file: mmap +
cudaHostRegister + H2D + cudaHostUnregister, against preadv into a pinned buffer + H2D."""
import ctypes
import mmap
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

size = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 105_000_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = tempfile.mkdtemp(dir="/dev/shm")
paths = []
for i in range(6):
    p = os.path.join(d, f"t{i}.bin")
    np.random.default_rng(i).integers(0, 255, size, dtype=np.uint8).tofile(p)
    paths.append(p)
rt = torch.cuda.cudart()
dev = torch.empty(size, dtype=torch.uint8, device="cuda")
pinned = torch.empty(size, dtype=torch.uint8, pin_memory=True)
stream = torch.cuda.current_stream()
libc = ctypes.CDLL(None)
cudart = ctypes.CDLL("libcudart.so.12") if False else None


def via_copy(p, pool):
    fd = os.open(p, os.O_RDONLY)
    mem = memoryview(pinned.numpy())
    chunk = 4 << 20

    def part(off):
        want = min(chunk, size - off)
        got = 0
        while got < want:
            got += os.preadv(fd, [mem[off + got : off + want]], off + got)

    list(pool.map(part, range(0, size, chunk)))
    os.close(fd)
    dev.copy_(pinned, non_blocking=True)
    stream.synchronize()


def via_register(p, flags, prot_write):
    fd = os.open(p, os.O_RDWR if prot_write else os.O_RDONLY)
    mm = mmap.mmap(fd, size, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ | (mmap.PROT_WRITE if prot_write else 0))
    arr = np.frombuffer(mm, dtype=np.uint8)
    ptr = arr.ctypes.data
    t0 = time.perf_counter()
    err = rt.cudaHostRegister(ptr, size, flags)
    t1 = time.perf_counter()
    if int(err) != 0:
        del arr
        mm.close()
        os.close(fd)
        raise RuntimeError(f"cudaHostRegister -> {err}")
    src = torch.from_numpy(arr) if prot_write else torch.frombuffer(mm, dtype=torch.uint8)
    dev.copy_(src, non_blocking=True)
    stream.synchronize()
    t2 = time.perf_counter()
    rt.cudaHostUnregister(ptr)
    t3 = time.perf_counter()
    del src, arr
    mm.close()
    os.close(fd)
    return t1 - t0, t2 - t1, t3 - t2


pool = ThreadPoolExecutor(max_workers=threads)
for rep in range(2):
    for p in paths:
        via_copy(p, pool)
t = time.perf_counter()
for p in paths:
    via_copy(p, pool)
dt = (time.perf_counter() - t) / len(paths)
print(f"preadv x{threads} into pinned + H2D: {dt * 1e3:.2f} ms per {size / 1e6:.0f} MB file = {size / dt / 1e9:.1f} GB/s")
for name, flags, pw in (("register (rw mapping)", 0, True), ("register read-only flag", 8, False)):
    try:
        for p in paths:
            via_register(p, flags, pw)
        acc = [0.0, 0.0, 0.0]
        t = time.perf_counter()
        for p in paths:
            for i, v in enumerate(via_register(p, flags, pw)):
                acc[i] += v
        dt = (time.perf_counter() - t) / len(paths)
        print(f"{name}: {dt * 1e3:.2f} ms per file = {size / dt / 1e9:.1f} GB/s   register {acc[0] / len(paths) * 1e3:.2f} ms, "
              f"H2D {acc[1] / len(paths) * 1e3:.2f} ms, unregister {acc[2] / len(paths) * 1e3:.2f} ms")
    except Exception as exc:  # noqa: BLE001
        print(f"{name}: {type(exc).__name__}: {exc}")
for p in paths:
    os.remove(p)
os.rmdir(d)
