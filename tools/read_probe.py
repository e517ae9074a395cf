"""How fast can this box move a file from tmpfs into pinned memory (tools only): preadv with k threads, or a fresh
mmap of the file per read + copy (page faults included), with and without MAP_POPULATE."""
import mmap
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

size = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 105_000_000
d = tempfile.mkdtemp(dir="/dev/shm")
paths = []
for i in range(4):
    p = os.path.join(d, f"t{i}.bin")
    np.random.default_rng(i).integers(0, 255, size, dtype=np.uint8).tofile(p)
    paths.append(p)
pinned = torch.empty(size, dtype=torch.uint8, pin_memory=True)
view = pinned.numpy()
mem = memoryview(view)
print("cpus", len(os.sched_getaffinity(0)))


def run(k, chunk, mode):
    pool = ThreadPoolExecutor(max_workers=k)
    best = 1e9
    for rep in range(8):
        p = paths[rep % len(paths)]
        t = time.perf_counter()
        fd = os.open(p, os.O_RDONLY)
        if mode.startswith("mmap"):
            flags = mmap.MAP_SHARED | (mmap.MAP_POPULATE if mode == "mmap+populate" else 0)
            mm = mmap.mmap(fd, size, flags=flags, prot=mmap.PROT_READ)
            src = np.frombuffer(mm, dtype=np.uint8)

        def part(off):
            want = min(chunk, size - off)
            if mode.startswith("mmap"):
                np.copyto(view[off : off + want], src[off : off + want])
            else:
                got = 0
                while got < want:
                    got += os.preadv(fd, [mem[off + got : off + want]], off + got)
            return want

        list(pool.map(part, range(0, size, chunk)))
        if mode.startswith("mmap"):
            del src
            mm.close()
        os.close(fd)
        if rep >= 2:
            best = min(best, time.perf_counter() - t)
    pool.shutdown()
    return size / best / 1e9


for mode in ("preadv", "mmap", "mmap+populate"):
    for k in (3, 8, 15):
        for chunk in (1 << 20, 4 << 20):
            print(f"{mode:14s} threads {k:2d} chunk {chunk >> 20:2d} MB: {run(k, chunk, mode):6.1f} GB/s", flush=True)
for p in paths:
    os.remove(p)
os.rmdir(d)
