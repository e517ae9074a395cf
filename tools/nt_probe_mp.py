"""Several processes (one per GPU rank in the real thing) reading files at once, 3 reader threads each: preadv into
pinned memory vs fresh mmap + non-temporal copy (tools only).  python tools/nt_probe_mp.py PROCS [THREADS]"""
import ctypes
import mmap
import multiprocessing as mp
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if not os.path.exists(os.path.join(HERE, "libcopy_nt_probe.so")):
    import subprocess

    subprocess.check_call(["gcc", "-O2", "-mavx2", "-shared", "-fPIC", os.path.join(HERE, "copy_nt_probe.c"), "-o",
                           os.path.join(HERE, "libcopy_nt_probe.so")])
size = 105_000_000
FILES = 6


def worker(rank, procs, threads, mode, barrier, out, root):
    import torch

    nt = ctypes.CDLL(os.path.join(HERE, "libcopy_nt_probe.so"))
    nt.copy_nt.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    cpus = sorted(os.sched_getaffinity(0))
    per = max(1, len(cpus) // procs)
    os.sched_setaffinity(0, cpus[rank * per : (rank + 1) * per])
    paths = [os.path.join(root, f"r{rank}_{i}.bin") for i in range(FILES)]
    for i, p in enumerate(paths):
        np.random.default_rng(rank * 100 + i).integers(0, 255, size, dtype=np.uint8).tofile(p)
    pinned = torch.empty(size + 64, dtype=torch.uint8, pin_memory=True)
    view = pinned.numpy()
    mem = memoryview(view)
    dst0 = view.ctypes.data
    pool = ThreadPoolExecutor(max_workers=threads)
    chunk = 4 << 20

    def read(p):
        fd = os.open(p, os.O_RDONLY)
        if mode == "mmap+nt":
            mm = mmap.mmap(fd, size, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ)
            src = np.frombuffer(mm, dtype=np.uint8)
            src0 = src.ctypes.data

        def part(off):
            want = min(chunk, size - off)
            if mode == "mmap+nt":
                nt.copy_nt(dst0 + off, src0 + off, want)
            else:
                got = 0
                while got < want:
                    got += os.preadv(fd, [mem[off + got : off + want]], off + got)

        list(pool.map(part, range(0, size, chunk)))
        if mode == "mmap+nt":
            del src
            mm.close()
        os.close(fd)

    for p in paths:
        read(p)
    barrier.wait()
    t = time.perf_counter()
    for _ in range(3):
        for p in paths:
            read(p)
    out[rank] = time.perf_counter() - t
    for p in paths:
        os.remove(p)


if __name__ == "__main__":
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    root = tempfile.mkdtemp(dir="/dev/shm")
    ctx = mp.get_context("spawn")
    for mode in ("preadv", "mmap+nt", "preadv", "mmap+nt"):
        barrier = ctx.Barrier(procs)
        out = ctx.Array("d", procs)
        ps = [ctx.Process(target=worker, args=(r, procs, threads, mode, barrier, out, root)) for r in range(procs)]
        for p in ps:
            p.start()
        for p in ps:
            p.join()
        wall = max(out)
        print(f"{procs} processes x {threads} threads, {mode:8s}: {procs * 3 * FILES * size / wall / 1e9:6.1f} GB/s aggregate", flush=True)
    os.rmdir(root)
