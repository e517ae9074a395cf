"""preadv into pinned memory vs a fresh mmap of the file + a copy with non-temporal stores (tools only)."""
import ctypes
import mmap
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if not os.path.exists(os.path.join(HERE, "libcopy_nt_probe.so")):
    import subprocess

    subprocess.check_call(["gcc", "-O2", "-mavx2", "-shared", "-fPIC", os.path.join(HERE, "copy_nt_probe.c"), "-o",
                           os.path.join(HERE, "libcopy_nt_probe.so")])
nt = ctypes.CDLL(os.path.join(HERE, "libcopy_nt_probe.so"))
nt.copy_nt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
size = 105_000_000
d = tempfile.mkdtemp(dir="/dev/shm")
paths = []
for i in range(6):
    p = os.path.join(d, f"t{i}.bin")
    np.random.default_rng(i).integers(0, 255, size, dtype=np.uint8).tofile(p)
    paths.append(p)
pinned = torch.empty(size + 64, dtype=torch.uint8, pin_memory=True)
view = pinned.numpy()
mem = memoryview(view)
dst0 = view.ctypes.data
print("cpus", len(os.sched_getaffinity(0)))


def run(k, chunk, mode):
    pool = ThreadPoolExecutor(max_workers=k)
    best = 1e9
    for rep in range(9):
        p = paths[rep % len(paths)]
        t = time.perf_counter()
        fd = os.open(p, os.O_RDONLY)
        if mode != "preadv":
            mm = mmap.mmap(fd, size, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ)
            src = np.frombuffer(mm, dtype=np.uint8)
            src0 = src.ctypes.data

        def part(off):
            want = min(chunk, size - off)
            if mode == "mmap+nt":
                nt.copy_nt(dst0 + off, src0 + off, want)
            elif mode == "mmap+copyto":
                np.copyto(view[off : off + want], src[off : off + want])
            else:
                got = 0
                while got < want:
                    got += os.preadv(fd, [mem[off + got : off + want]], off + got)
            return want

        list(pool.map(part, range(0, size, chunk)))
        if mode != "preadv":
            del src
            mm.close()
        os.close(fd)
        if rep >= 3:
            best = min(best, time.perf_counter() - t)
    pool.shutdown()
    return size / best / 1e9


for k in (3, 8, 15):
    for mode in ("preadv", "mmap+copyto", "mmap+nt"):
        print(f"threads {k:2d} {mode:12s}: {run(k, 4 << 20, mode):6.1f} GB/s", flush=True)
for p in paths:
    os.remove(p)
os.rmdir(d)
