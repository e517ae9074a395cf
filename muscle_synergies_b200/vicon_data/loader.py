"""load_vicon_file on B200: host glue around the CUDA scan / parse kernels.

Drop-in for the reference's `load_vicon_file(csv_filename) -> ViconNexusData`
(src/muscle_synergies/vicon_data/load_csv.py:96-135).  Flow per file:

  1. file bytes -> pinned host buffer -> HBM (one async copy);
  2. ms_scan (CUDA): row terminators, blank (section separator) rows, quote count;
  3. host: the 2 x 5 header lines (a few KB) through `HeaderMachine`, which applies the
     reference's header rules (header.py);
  4. ms_parse (CUDA): every data row of both sections -> channel-major float64 blocks in HBM,
     bit-identical to float() per field;
  5. host: wrap the blocks in ViconNexusData / DeviceData (DataFrames are built lazily).

Errors: the kernels report the byte offset of the first field float() would reject; the host
replays the reference's rule on that single row to raise the same
`RuntimeError("error parsing line i of file f: ...")` chained to the same cause.

There is no CPU data path: without a CUDA device or without libms_b200.so every entry
point raises.
"""
import contextlib
import ctypes
import os
import re
import threading
import time
from typing import Callable, List, Optional, Tuple

import numpy as np

from .. import _native as nat
from .data_model import (
    DeviceData,
    ForcesEMGFrameTracker,
    SectionBlock,
    TrajFrameTracker,
    ViconNexusData,
)
from .definitions import DeviceType, SamplingFreq, SectionType
from .header import (
    HeaderMachine,
    SectionLayout,
    file_encoding,
    replay_data_row,
    rows_from_bytes,
    wrap_error,
)

_TERMINATOR = re.compile(rb"\r\n|\r|\n")
_HEADER_LINES = 5
_PEEK = 1 << 13  # bytes of header text fetched with the scan summary; take_lines' first window
_META_INFO = 256  # offset of the peek's {offset, count} behind the scan summary in the loader's meta buffer
_header_cache = {}  # (header bytes, section, lines) -> parsed SectionLayout


_READ_CHUNK = 32 << 20
_READ_CHUNK_MIN = 2 << 20
_READ_THREADS_MAX = int(os.environ.get("MS_B200_READ_THREADS", "16"))
_READ_MODE = os.environ.get("MS_B200_READ_MODE", "auto")  # auto | preadv | mmap (tools / tests)
_read_pool = None

# Which kernels load a file: None = the single-pass kernel (ms_load_fused) with the two-pass path
# (ms_scan -> host -> ms_parse) behind it for everything it declines; "two_pass" = the two-pass path only.
# FORCE_TILE: CSV bytes per thread block of the single-pass kernel (tests sweep it to move tile boundaries).
FORCE_PATH = os.environ.get("MS_B200_LOADER") or None
FORCE_TILE = None
TIMELINE = None  # tools/step_timeline.py: a list that receives (label, perf_counter) marks of the single-pass path
TUNE_TILE = os.environ.get("MS_B200_TUNE_TILE") == "1"  # size tiles to just under a multiple of 32 rows
_FMETA_PEEK = 256  # offset of the two header peeks behind ms_load_result in the single-pass meta buffer
_FMETA_BYTES = _FMETA_PEEK + 2 * nat.MS_LOAD_PEEK


def read_file_into(name, view: np.ndarray, size: int, on_chunk: Optional[Callable[[int, int], None]] = None) -> None:
    """Reads the first `size` bytes of file `name` into the uint8 array `view` with several threads (one
    thread copies out of the page cache at ~5 GB/s, a tenth of what the PCIe link then moves); `on_chunk(offset,
    length)` is called in file order as the chunks complete."""
    global _read_pool
    if _read_pool is None:
        from concurrent.futures import ThreadPoolExecutor

        try:
            cpus = len(os.sched_getaffinity(0))  # this process's share of the box (one rank per GPU may bind a subset)
        except AttributeError:
            cpus = os.cpu_count() or 2
        # one process per GPU (torchrun): the ranks of a box share its cores - eight pools of 15 threads on 32 cores
        # only get in each other's way (measured: 41 GB/s aggregate, against 62 for the same files read alone)
        try:
            cpus = max(1, cpus // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
        except ValueError:
            pass
        # preadv out of the page cache into pinned memory: 5 GB/s per thread, 34 GB/s with 8 and 45-53 GB/s with 15 threads
        # on the 16-core share of a 1-GPU box (tools/read_probe.py); the PCIe link takes 55 GB/s
        _read_pool = ThreadPoolExecutor(max_workers=max(2, min(_READ_THREADS_MAX, cpus - 1)), thread_name_prefix="ms-read")
    # two chunks per worker: a 100 MB trial keeps the whole pool busy (in 32 MB chunks it kept four threads busy)
    workers = _read_pool._max_workers
    chunk = min(_READ_CHUNK, max(_READ_CHUNK_MIN, -(-size // (2 * workers * _READ_CHUNK_MIN)) * _READ_CHUNK_MIN))
    fd = os.open(name, os.O_RDONLY)
    mapped = None
    try:
        mem = memoryview(view)
        # Few reader threads (one rank per GPU shares the box's cores with the others): the file is mapped and copied
        # with non-temporal stores - a third less DRAM traffic per byte, which is what several ranks reading at once are
        # bound by (4 processes x 3 threads: 44 -> 58 GB/s).  Many threads in one process: preadv, which does not
        # contend for the address space's lock on page faults (15 threads: 44 GB/s against 39.5 mapped).
        if size > 0 and (_READ_MODE == "mmap" or (_READ_MODE == "auto" and workers <= 8 and size >= (1 << 20))):
            if os.fstat(fd).st_size < size:
                raise IOError(f"short read on {name}: the file has fewer than {size} bytes")
            import mmap

            mapped = mmap.mmap(fd, size, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ)
            src0 = np.frombuffer(mapped, dtype=np.uint8)
            src_ptr, dst_ptr = src0.ctypes.data, view.ctypes.data
            copy = nat.lib().ms_host_copy_stream

            def part(off):
                want = min(chunk, size - off)
                nat.check(copy(dst_ptr + off, src_ptr + off, want), "ms_host_copy_stream")
                return off, want
        else:
            def part(off):
                want = min(chunk, size - off)
                got = 0
                while got < want:
                    k = os.preadv(fd, [mem[off + got : off + want]], off + got)
                    if k <= 0:
                        raise IOError(f"short read on {name}: {off + got} of {size} bytes")
                    got += k
                return off, want

        for off, want in _read_pool.map(part, range(0, size, chunk)):
            if on_chunk is not None:
                on_chunk(off, want)
    finally:
        if mapped is not None:
            src0 = None
            try:
                mapped.close()
            except BufferError:  # a worker's exception left a view alive: the map goes with it
                pass
        os.close(fd)


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise nat.NativeError("muscle_synergies_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16 + 16


class _Source:
    """The CSV bytes: a device tensor and, when there is one, the host copy."""

    def __init__(self, d_bytes, n: int, host: Optional[np.ndarray]):
        self.d_bytes = d_bytes
        self.n = n
        self.host = host
        self.windows = []  # (offset, bytes) pieces that came back with the scan summary

    def fetch(self, offset: int, length: int) -> bytes:
        offset = max(0, offset)
        end = min(self.n, offset + length)
        if end <= offset:
            return b""
        if self.host is not None:
            return self.host[offset:end].tobytes()
        for w_off, w_bytes in self.windows:
            if w_off <= offset and end <= w_off + len(w_bytes):
                return w_bytes[offset - w_off : end - w_off]
        return self.d_bytes[offset:end].cpu().numpy().tobytes()

    def take_lines(self, offset: int, k: int) -> Tuple[bytes, int]:
        """Bytes of the first k physical lines starting at `offset` (fewer at EOF)."""
        window = _PEEK
        while True:
            chunk = self.fetch(offset, window)
            ends = []
            for m in _TERMINATOR.finditer(chunk):
                # a '\r' at the very end of the window may be half of a '\r\n'
                if m.end() == len(chunk) and chunk[-1:] == b"\r" and offset + len(chunk) < self.n:
                    break
                ends.append(m.end())
                if len(ends) == k:
                    return chunk[: ends[-1]], k
            if offset + len(chunk) >= self.n:
                tail = chunk[ends[-1]:] if ends else chunk
                return chunk, len(ends) + (1 if tail else 0)
            window *= 4

    def locate_row(self, ws, pos: int):
        """(csv row index, first byte, one past the terminator) of the row holding byte `pos`,
        from the delimiter masks the scan kernel left in the workspace (error path only)."""
        torch = _torch()
        lib = nat.lib()
        off = int(lib.ms_workspace_masks_offset(self.n))
        n_seg = (self.n + 15) // 16
        masks = ws[off : off + 4 * n_seg].view(torch.int32)
        seg = pos // 16
        term = masks & 0xFFFF
        bits = torch.zeros((), dtype=torch.int64, device=masks.device)
        for k in range(16):  # popcount of the terminator bits before the segment of `pos`
            bits += ((term[:seg] >> k) & 1).sum()
        reach = nat.MS_MAX_ROW_BYTES // 16 + 2
        while True:  # widen until the row's two ends are in the window (rows may be longer than a tile's overhang)
            lo, hi = max(0, seg - reach), min(n_seg, seg + reach)
            window = term[lo:hi].cpu().numpy().astype(np.int64)
            ends = [16 * (lo + i) + b for i, m in enumerate(window) if m for b in range(16) if (m >> b) & 1]
            before = [e for e in ends if e < pos]
            after = [e for e in ends if e >= pos]
            if (before or lo == 0) and (after or hi == n_seg):
                break
            reach *= 8
        row = int(bits.item()) + sum(1 for e in before if e >= 16 * seg)
        start = before[-1] + 1 if before else 0
        stop = after[0] + 1 if after else self.n
        return row, start, stop


class ViconLoader:
    """Reusable loader bound to one CUDA device and stream."""

    def __init__(self, device=None, stream=None):
        torch = _torch()
        self.torch = torch
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.stream = stream
        self.lib = nat.lib()
        self._pinned = None
        # what every scan sends back: [summary | {offset, count} | the PEEK bytes after the first blank row]
        # (ms_peek_after_blank) in one copy, and the first PEEK bytes of the file; buffers live with the loader
        self._d_meta = torch.empty(_META_INFO + 16 + _PEEK, dtype=torch.uint8, device=self.device)
        self._pinned_meta = torch.empty(_META_INFO + 16 + _PEEK, dtype=torch.uint8, pin_memory=True)
        self._pinned_head = torch.empty(_PEEK, dtype=torch.uint8, pin_memory=True)
        self._pinned_status = torch.empty(1, dtype=torch.int64, pin_memory=True)
        self._side_streams = None  # copy-in / copy-out streams of load_many, created on first use
        self._pipe_stream = None  # load_device_many: the two streams its single-pass kernels are queued on in turn
        self._work_stream = None  # high-priority stream offered to the caller of load_device_many
        # single-pass path: what the previous file looked like (sizes the next file's arena and tile), pinned
        # buffers on loan to results (data_model.HostLease), meta buffers, counters
        self._history = None
        self._host_pool = []
        self._fmeta_pool = []
        self.stats = {"fused": 0, "two_pass": 0, "declined": {}}
        self._fused_skip = self._fused_backoff = 0
        self._lock = threading.RLock()  # one load at a time per loader: its staging and meta buffers are shared

    # ---- public -----------------------------------------------------------------------------
    def load_file(self, csv_filename) -> ViconNexusData:
        with self._lock, self.torch.cuda.device(self.device):
            return self._load_file(csv_filename)

    def _load_file(self, csv_filename) -> ViconNexusData:
        torch = self.torch
        size = os.path.getsize(csv_filename)  # FileNotFoundError propagates unwrapped, like open()
        staging = self._staging(size)
        view = staging.numpy()
        stream, _ = self._stream_ptr()
        with self._on_stream(stream):
            d_bytes = torch.empty(_pad16(size), dtype=torch.uint8, device=self.device)

            def uploaded(off, length):  # each chunk goes to the GPU while the next ones are still being read
                d_bytes[off : off + length].copy_(staging[off : off + length], non_blocking=True)

            read_file_into(csv_filename, view, size, uploaded)
        return self._run(_Source(d_bytes, size, view[:size]), str(csv_filename))

    def load_bytes(self, data, name: str = "<bytes>") -> ViconNexusData:
        """`data`: bytes / bytearray / uint8 numpy array / uint8 CPU tensor holding the CSV."""
        with self._lock, self.torch.cuda.device(self.device):
            return self._load_bytes(data, name)

    def _load_bytes(self, data, name: str) -> ViconNexusData:
        torch = self.torch
        if isinstance(data, torch.Tensor):
            if data.is_cuda:
                return self.load_device(data, name=name)
            host = data.numpy()
        else:
            host = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        n = host.shape[0]
        if isinstance(data, torch.Tensor) and data.is_pinned():
            staging = data
        else:
            staging = self._staging(n)
            staging.numpy()[:n] = host
        return self._load_host(host, staging, name)

    def load_device(self, d_bytes, n: Optional[int] = None, name: str = "<device bytes>", host=None,
                    defer_check: bool = False) -> ViconNexusData:
        """CSV bytes already in HBM.  The tensor must be readable 16 bytes past `n` rounded up
        to 16 (allocate it with `padded_size(n)`).

        defer_check (extension): return as soon as the parse kernel is queued; whatever the data rows
        would have raised is raised by `data.check()` instead - which `Segmenter(data)` calls after its
        own wait - so that a load followed by device-side work costs one host wait less."""
        n = int(d_bytes.numel() if n is None else n)
        if d_bytes.numel() < _pad16(n):
            torch = self.torch
            padded = torch.empty(_pad16(n), dtype=torch.uint8, device=self.device)
            padded[:n].copy_(d_bytes[:n])
            d_bytes = padded
        return self._run(_Source(d_bytes, n, host), name, defer_check)

    @property
    def work_stream(self):
        """A high-priority stream of this loader's device for the caller's own work on the trials that
        `load_device_many` yields (Segmenter, window cuts): its small kernels then run beside the next trial's
        loader kernel instead of behind it."""
        if self._work_stream is None:
            _, highest = self.torch.cuda.Stream.priority_range()
            self._work_stream = self.torch.cuda.Stream(self.device, priority=highest)
        return self._work_stream

    def load_device_many(self, sources, names=None, depth: int = 2, defer_check: bool = True, stream=None):
        """Trials whose CSV bytes are already in HBM, `depth` of them in flight (extension): yields one
        ViconNexusData per source, in order.

        sources: an iterable of CUDA uint8 tensors or (tensor, n) pairs, each readable to `padded_size(n)`.
        The single-pass kernels of the next `depth` trials are queued on the loader's own pipeline streams (two,
        taken in turn: consecutive trials are independent, so one kernel's last blocks and the next one's first
        share the GPU) before a trial is handed to the caller, so the GPU parses trial i + 1 while the host finishes
        the objects of trial i and the caller works on them.  stream: the stream the caller uses the results on
        (default: the current one; `loader.work_stream` is a high-priority stream made for it).  What a trial's
        data rows would have raised is raised when the trial is reached - by `data.check()` / `Segmenter(data)`
        with defer_check, as for `load_device`.  A trial the single-pass kernel declines runs through the two-pass
        path when it is reached; results are the same either way."""
        import collections

        torch = self.torch
        pipes = self._pipes()
        it = iter(sources)
        name_iter = iter(names) if names is not None else None
        pending = collections.deque()
        counter = [0]

        def submit():
            try:
                item = next(it)
            except StopIteration:
                return False
            d_bytes, n = item if isinstance(item, (tuple, list)) else (item, None)
            n = int(d_bytes.numel() if n is None else n)
            name = next(name_iter) if name_iter is not None else f"<device bytes {counter[0]}>"
            counter[0] += 1
            with self._lock, torch.cuda.device(self.device):
                if d_bytes.numel() < _pad16(n):
                    padded = torch.empty(_pad16(n), dtype=torch.uint8, device=self.device)
                    padded[:n].copy_(d_bytes[:n])
                    d_bytes = padded
                src = _Source(d_bytes, n, None)
                ticket = None
                if self._want_fused(n):
                    # the bytes were written on the caller's stream: the pipeline stream reads them after that
                    pipe = pipes[counter[0] & 1]
                    pipe.wait_stream(torch.cuda.current_stream(self.device))
                    ticket = self._submit_fused(src, name, stream=pipe)
            pending.append((src, name, ticket))
            return True

        while len(pending) < max(1, depth) and submit():
            pass
        while pending:
            src, name, ticket = pending.popleft()
            with self._lock, torch.cuda.device(self.device):
                data = self._finish_fused(ticket) if ticket is not None else None
                if data is not None:
                    self.stats["fused"] += 1
                    self._fused_backoff = 0
                    # allocated on a pipeline stream, used (and let go) on the caller's
                    cur = stream if stream is not None else torch.cuda.current_stream(self.device)
                    for blk in data.blocks:
                        if blk.tensor is not None:
                            blk.tensor.record_stream(cur)
                else:
                    self.stats["two_pass"] += 1
                    data = self._run_two_pass(src, name, defer_check)
            submit()  # keep `depth` trials queued while the caller works on this one
            yield data

    @staticmethod
    def padded_size(n: int) -> int:
        return _pad16(n)

    # ---- internals -----------------------------------------------------------------------------
    def _staging(self, n: int):
        torch = self.torch
        need = _pad16(n)
        if self._pinned is None or self._pinned.numel() < need:
            self._pinned = torch.empty(max(need, 1 << 20), dtype=torch.uint8, pin_memory=True)
        return self._pinned

    def _stream_ptr(self):
        s = self.stream if self.stream is not None else self.torch.cuda.current_stream(self.device)
        return s, ctypes.c_void_p(s.cuda_stream)

    def _load_host(self, host: np.ndarray, staging, name: str) -> ViconNexusData:
        torch = self.torch
        n = int(host.shape[0])
        stream, _ = self._stream_ptr()
        with torch.cuda.stream(stream):
            d_bytes = torch.empty(_pad16(n), dtype=torch.uint8, device=self.device)
            if n:
                d_bytes[:n].copy_(staging[:n], non_blocking=True)
        return self._run(_Source(d_bytes, n, host), name)

    def _on_stream(self, stream):
        """Context that makes `stream` current - nothing to do when the loader uses the current stream."""
        return self.torch.cuda.stream(stream) if self.stream is not None else contextlib.nullcontext()

    def _scan(self, src: _Source, ws=None):
        """ms_scan (or, when `ws` holds a previous scan, the quote-aware ms_scan_quoted)."""
        torch = self.torch
        stream, sptr = self._stream_ptr()
        ws_bytes = int(self.lib.ms_workspace_bytes(src.n))
        entry, what = (self.lib.ms_scan, "ms_scan") if ws is None else (self.lib.ms_scan_quoted, "ms_scan_quoted")
        # device side of what travels back: [summary | {offset, count} of the peek | peek bytes], one copy
        meta = self._d_meta
        d_summary, d_info, d_peek = meta.data_ptr(), meta.data_ptr() + _META_INFO, meta.data_ptr() + _META_INFO + 16
        with self._on_stream(stream):
            if ws is None:
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            nat.check(entry(src.d_bytes.data_ptr(), src.n, ws.data_ptr(), ws_bytes, d_summary, sptr), what)
            peek = src.host is None and src.n > 0
            if peek:
                # both sections' header lines come back with the summary (one wait instead of three)
                k0 = min(src.n, _PEEK)
                self._pinned_head[:k0].copy_(src.d_bytes[:k0], non_blocking=True)
                nat.check(self.lib.ms_peek_after_blank(src.d_bytes.data_ptr(), src.n, d_summary, 0, d_peek, _PEEK, d_info,
                                                       sptr), "ms_peek_after_blank")
                self._pinned_meta.copy_(meta, non_blocking=True)
            else:
                self._pinned_meta[:_META_INFO].copy_(meta[:_META_INFO], non_blocking=True)
        stream.synchronize()
        host = self._pinned_meta.numpy()
        summary = nat.ScanSummary.from_buffer_copy(host[: ctypes.sizeof(nat.ScanSummary)].tobytes())
        if peek:
            src.windows = [(0, self._pinned_head.numpy()[:k0].tobytes())]
            off2, cnt2 = (int(v) for v in host[_META_INFO : _META_INFO + 16].view(np.int64))
            if off2 >= 0 and cnt2 > 0:
                src.windows.append((off2, host[_META_INFO + 16 : _META_INFO + 16 + cnt2].tobytes()))
        return summary, ws

    def _want_fused(self, n: int) -> bool:
        """Whether the single-pass kernel gets this file (call once per file: it counts down the back-off)."""
        if FORCE_PATH == "two_pass" or n <= 0:
            return False
        if self._fused_skip > 0 and FORCE_PATH != "fused":
            self._fused_skip -= 1  # the last files were not for the single-pass kernel: do not run both on every one
            return False
        return True

    def _pipes(self):
        """The two streams single-pass kernels of consecutive files are queued on in turn (batch entry points)."""
        if self._pipe_stream is None:
            self._pipe_stream = (self.torch.cuda.Stream(self.device), self.torch.cuda.Stream(self.device))
        return self._pipe_stream

    def _run(self, src: _Source, name: str, defer_check: bool = False) -> ViconNexusData:
        with self._lock, self.torch.cuda.device(self.device):
            if self._want_fused(src.n):
                data = self._run_fused(src, name)
                if data is not None:
                    self.stats["fused"] += 1
                    self._fused_backoff = 0
                    return data
            self.stats["two_pass"] += 1
            return self._run_two_pass(src, name, defer_check)

    # ---- single pass ---------------------------------------------------------------------------------
    def _decline(self, why: str, back_off: bool = True):
        self.stats["declined"][why] = self.stats["declined"].get(why, 0) + 1
        if back_off:
            self._fused_backoff = min(64, max(2, 2 * self._fused_backoff))
            self._fused_skip = self._fused_backoff
        return None

    def _fused_sizes(self, src: _Source, name: str):
        """(arena doubles, row capacity of each section, tile bytes) for ms_load_fused, or None.  From the previous
        file this loader parsed when there is one (trials of one session share a layout); else from the first
        rows of this file.  A wrong guess costs a second run (the kernel reports MS_LOAD_OVERFLOW), never a
        wrong array."""
        n = src.n
        h = self._history
        if h is not None:
            scale = n / h["n"]
            cap1 = int(h["rows"][0] * scale * 1.02) + 64
            cap2 = int(h["rows"][1] * scale * 1.05) + 64
            arena = h["keep"][0] * cap1 + 2 + h["keep"][1] * cap2
            row_len = h["row_len"]
            # a tile stages this much past its end: room for a row twice the longest average of the last file
            overhang = h.get("overhang") or min(nat.MS_MAX_ROW_BYTES, max(1024, -(-int(2 * max(row_len, h["row_len2"]) + 64) // 512) * 512))
        else:
            m1 = HeaderMachine(SectionType.FORCES_EMG)
            try:
                err, hdr = _feed_header(src, 0, 0, 1 << 62, m1, name)
            except Exception:  # noqa: BLE001 - the two-pass path raises it in the reference's words
                return None
            if err is not None or not m1.done or m1.layout.num_cols < 3:
                return None
            rest = src.fetch(len(hdr), _PEEK)
            ends = [m.end() for m in _TERMINATOR.finditer(rest)]
            if len(ends) < 2:
                return None
            row_len = ends[-1] / len(ends)
            cap1 = int(n / row_len * 1.03) + 64
            cap2 = 0  # the second section takes what is left
            arena = (m1.layout.num_cols - 2) * cap1 + 2 + n // 12 + 4096
            overhang = nat.MS_MAX_ROW_BYTES
        tile = 0  # the largest tile that fits beside the overhang
        groups = int((nat.MS_TILE_BYTES + row_len / 2) // (32 * row_len))
        if groups >= 1 and TUNE_TILE:
            # just under a multiple of 32 rows per tile: the kernel parses rows in groups of 32 lanes
            tile = max(4096, min(nat.MS_TILE_BYTES, int((32 * groups - 2) * row_len) // 16 * 16))
        if FORCE_TILE is not None:
            tile, overhang = int(FORCE_TILE), nat.MS_MAX_ROW_BYTES
        return arena, cap1, cap2, tile, overhang

    def _fmeta_acquire(self):
        if self._fmeta_pool:
            return self._fmeta_pool.pop()
        torch = self.torch
        return (torch.empty(_FMETA_BYTES, dtype=torch.uint8, device=self.device),
                torch.empty(_FMETA_BYTES, dtype=torch.uint8, pin_memory=True))

    def _run_fused(self, src: _Source, name: str) -> Optional[ViconNexusData]:
        """ms_load_fused: bytes -> both sections' blocks in one launch.  Returns None when the kernel (or the host's
        reading of the header text it sends back) says the file is not plain enough; the caller then runs the
        two-pass path, which handles - and words the errors of - everything."""
        ticket = self._submit_fused(src, name)
        return None if ticket is None else self._finish_fused(ticket)

    def _submit_fused(self, src: _Source, name: str, stream=None):
        """Queues ms_load_fused and the copy of its result word on `stream` (default: the loader's), builds the
        objects of a file that looks like the previous one, and returns what `_finish_fused` needs - without
        waiting for the GPU.  None: no size estimate (the caller runs the two-pass path)."""
        torch = self.torch
        mark = (lambda label: TIMELINE.append((label, time.perf_counter()))) if TIMELINE is not None else (lambda label: None)
        mark("enter")
        sizes = self._fused_sizes(src, name)
        if sizes is None:
            return self._decline("no size estimate", back_off=False)
        arena_elems, cap1, cap2, tile, overhang = sizes
        self.last_plan = (tile, overhang)  # what the kernel was asked for (bench.py times the kernel with the same)
        if stream is None:
            stream, sptr = self._stream_ptr()
            on_stream = self._on_stream(stream)
        else:
            sptr, on_stream = ctypes.c_void_p(stream.cuda_stream), torch.cuda.stream(stream)
        meta = self._fmeta_acquire()
        d_meta, h_meta = meta
        with on_stream:
            arena = torch.empty(arena_elems, dtype=torch.float64, device=self.device)
            ws_bytes = int(self.lib.ms_load_workspace_bytes(src.n, tile))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            plan = nat.LoadPlan(arena.data_ptr(), arena_elems, (ctypes.c_int64 * 2)(cap1, cap2), tile, overhang)
            nat.check(self.lib.ms_load_fused(src.d_bytes.data_ptr(), src.n, ctypes.byref(plan), ws.data_ptr(), ws_bytes,
                                             d_meta.data_ptr(), d_meta.data_ptr() + _FMETA_PEEK, sptr), "ms_load_fused")
            h_meta.copy_(d_meta, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(stream)
        # While the kernel runs: the objects of a file that looks like the previous one (same header text -> the same
        # cached layouts).  Whether it does is checked below, once the header text is back; row counts and the
        # arena views are filled in then.
        mark("launched")
        ahead = None
        prev_layouts = (self._history or {}).get("layouts")
        if prev_layouts is not None:
            try:
                guess = _Plan()
                guess.layouts = prev_layouts
                guess_blocks = [SectionBlock(None, 0), SectionBlock(None, 0)]
                ahead = (_build(guess, guess_blocks), guess_blocks)
            except (TypeError, ValueError, KeyError):
                ahead = None  # Builder.build would raise: on the ordinary path below, after the rows are known good
        mark("built ahead")
        # arena / ws ride along: the kernel writes them until `copied`
        return (src, name, arena, ws, meta, copied, ahead, prev_layouts, overhang)

    def _finish_fused(self, ticket) -> Optional[ViconNexusData]:
        """Waits for a submitted ms_load_fused, reads its result word and header text, and finishes the objects.
        None: the file is not for the single-pass kernel (see _run_fused)."""
        mark = (lambda label: TIMELINE.append((label, time.perf_counter()))) if TIMELINE is not None else (lambda label: None)
        src, name, arena, ws, meta, copied, ahead, prev_layouts, overhang = ticket
        d_meta, h_meta = meta
        copied.synchronize()
        mark("kernel done")
        host = h_meta.numpy()
        res = nat.LoadResult.from_buffer_copy(host[: ctypes.sizeof(nat.LoadResult)].tobytes())
        peeks = [host[_FMETA_PEEK + s * nat.MS_LOAD_PEEK : _FMETA_PEEK + s * nat.MS_LOAD_PEEK + max(0, int(res.peek_bytes[s]))].tobytes()
                 if res.have & (nat.MS_LOAD_HAVE_HEADER0 << s) else b"" for s in (0, 1)]
        self._fmeta_pool.append(meta)
        self.last_result = res

        if res.flags:
            # a wrong size guess (alone) is no reason to avoid the kernel: the two-pass run below seeds the next guess
            if res.flags == 1 and overhang < nat.MS_MAX_ROW_BYTES and self._history is not None:
                self._history["overhang"] = nat.MS_MAX_ROW_BYTES  # a row longer than the last file suggested
                return self._decline("flags ROW_TOO_LONG (short overhang)", back_off=False)
            return self._decline("flags " + "|".join(v for k, v in nat.MS_LOAD_FLAG_NAMES.items() if res.flags & k),
                                 back_off=res.flags != 16)
        if res.status != nat.MS_ERR_NONE:
            return self._decline("bad field")
        want = 0
        for s in (0, 1):
            want |= (nat.MS_LOAD_HAVE_HEADER0 | nat.MS_LOAD_HAVE_DESC0 | nat.MS_LOAD_HAVE_ROWS0) << s
        if res.have & want != want:
            return self._decline("not two sections")
        if not (res.n_blank_rows == 1 or (res.n_blank_rows == 2 and res.tail_rows == 0)):
            return self._decline("blank rows")
        if res.data_rows[0] < 0 or res.data_rows[1] < 0:
            return self._decline("short header")
        layouts, header_quotes = [], 0
        for s, kind in ((0, SectionType.FORCES_EMG), (1, SectionType.TRAJECTORIES)):
            at_eof = int(res.header_offset[s]) + len(peeks[s]) >= src.n
            chunk = _first_lines(peeks[s], _HEADER_LINES, at_eof)
            if chunk is None:
                return self._decline("header text")
            header_quotes += chunk.count(b'"')
            lay = _header_cache.get((chunk, kind, _HEADER_LINES))
            if lay is None:
                machine = HeaderMachine(kind)
                try:
                    rows, consumed = rows_from_bytes(chunk, _HEADER_LINES)
                    if consumed != len(rows) or len(rows) != _HEADER_LINES:
                        return self._decline("header text")
                    for row in rows:
                        machine.feed(row)
                except Exception:  # noqa: BLE001 - raised, worded and numbered by the two-pass path
                    return self._decline("header error")
                lay = machine.layout
                if lay.complete and len(_header_cache) < 64:
                    _header_cache[(chunk, kind, _HEADER_LINES)] = lay
            if not lay.complete or lay.num_cols != res.num_cols[s] or lay.n_keep > res.n_keep[s] or lay.num_cols < 3:
                return self._decline("column count")
            layouts.append(lay)
        if int(res.n_quotes) > header_quotes:
            return self._decline("quotes")

        rows = [int(res.data_rows[0]), int(res.data_rows[1])]
        plan_ = _Plan()
        plan_.layouts = layouts
        first2 = _HEADER_LINES + rows[0] + 1 + _HEADER_LINES
        plan_.data_rows = [(_HEADER_LINES, _HEADER_LINES + rows[0]), (first2, first2 + rows[1])]
        reuse = ahead is not None and layouts[0] is prev_layouts[0] and layouts[1] is prev_layouts[1]
        blocks = ahead[1] if reuse else [SectionBlock(None, 0), SectionBlock(None, 0)]
        for s in (0, 1):
            keep, stride, off = int(res.n_keep[s]), int(res.stride[s]), int(res.out_offset[s])
            blocks[s].tensor = arena[off : off + keep * stride].view(keep, stride)[: layouts[s].n_keep]
            blocks[s].n_rows = rows[s]
        sec1_bytes = int(res.blank_end[0]) + 1
        keep_overhang = (self._history or {}).get("overhang")
        self._history = {"n": src.n, "rows": rows, "keep": [int(res.n_keep[0]), int(res.n_keep[1])],
                         "row_len": max(1.0, sec1_bytes / max(1, rows[0] + _HEADER_LINES + 1)),
                         "row_len2": max(1.0, (src.n - sec1_bytes) / max(1, rows[1] + _HEADER_LINES + 1)),
                         "overhang": keep_overhang, "layouts": layouts}
        if reuse:
            data = ahead[0]
            data.emg._frame_tracker._sampling_freq.num_frames = rows[1]  # one SamplingFreq behind every tracker
        else:
            data = _build(plan_, blocks)  # raises what Builder.build raises (user_data.py:310-433): the rows are clean
        data.blocks = blocks
        mark("returned")
        return data

    # ---- two passes ------------------------------------------------------------------------------------
    def _run_two_pass(self, src: _Source, name: str, defer_check: bool = False) -> ViconNexusData:
        torch = self.torch
        summary, ws = self._scan(src)
        try:
            plan = _plan(src, summary, name)
        except _QuotedData:
            # '"' in the data rows: rescan with the csv in-quote state, then plan again
            summary, ws = self._scan(src, ws)
            plan = _plan(src, summary, name, quoted=True)
        stream, sptr = self._stream_ptr()

        sections = (nat.Section * nat.MS_MAX_SECTIONS)()
        blocks: List[Optional[SectionBlock]] = []
        n_sec = 0
        with self._on_stream(stream):
            for lay, (r0, r1) in zip(plan.layouts, plan.data_rows):
                if lay is None or not lay.complete:
                    blocks.append(None)
                    continue
                n_rows = max(0, r1 - r0)
                block = torch.empty((lay.n_keep, n_rows), dtype=torch.float64, device=self.device)
                blocks.append(SectionBlock(block, n_rows))
                if n_rows > 0:
                    s = sections[n_sec]
                    s.row_begin, s.row_end = r0, r1
                    s.num_cols, s.n_keep = lay.num_cols, lay.n_keep
                    s.d_out = block.data_ptr() if lay.n_keep > 0 else None
                    s.stride = n_rows
                    n_sec += 1
            d_status = torch.empty(1, dtype=torch.int64, device=self.device)
            nat.check(
                self.lib.ms_parse(src.d_bytes.data_ptr(), src.n, ws.data_ptr(), sections, n_sec, d_status.data_ptr(), sptr),
                "ms_parse",
            )
            if defer_check:
                h_status = torch.empty(1, dtype=torch.int64, pin_memory=True)  # outlives this call
            else:
                h_status = self._pinned_status
            h_status.copy_(d_status, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(stream)
        # the objects are built while the kernel runs; errors keep the reference's order: a bad data
        # row first (raised while reading), the builder's complaints last (user_data.py:310-433)
        data, build_error = None, None
        try:
            data = _build(plan, blocks)
            data.blocks = [b for b in blocks if b is not None]  # the per-section HBM blocks (extension)
        except (TypeError, ValueError, KeyError) as exc:
            build_error = exc

        n_terminators = int(summary.n_terminators)

        def check():
            copied.synchronize()
            key = int(h_status.item()) & 0xFFFFFFFFFFFFFFFF
            if key != nat.MS_ERR_NONE and key & 7 == nat.MS_ERR_KIND_ROW_TOO_LONG:
                # a row the tiled kernel cannot stage: the whole buffer again with one thread per row (ms_parse_long)
                with self.torch.cuda.device(self.device), self._on_stream(stream):
                    d_rows = torch.empty(int(self.lib.ms_parse_long_workspace_bytes(n_terminators)), dtype=torch.uint8,
                                         device=self.device)
                    nat.check(self.lib.ms_parse_long(src.d_bytes.data_ptr(), src.n, ws.data_ptr(), sections, n_sec,
                                                     n_terminators, d_rows.data_ptr(), d_status.data_ptr(), sptr), "ms_parse_long")
                    key = int(d_status.cpu().item()) & 0xFFFFFFFFFFFFFFFF
            if key != nat.MS_ERR_NONE:
                _raise_device_error(src, plan, key, name, ws)
            if plan.deferred_error is not None:
                raise plan.deferred_error
            if build_error is not None:
                raise build_error

        if defer_check and data is not None:
            data._pending_check = check
            return data
        check()
        self._remember(src, plan)
        return data

    def _remember(self, src: _Source, plan):
        """Shapes of a file the two-pass path parsed: the single-pass kernel's size guess for the next one."""
        lays, rows = plan.layouts, [max(0, b - a) for a, b in plan.data_rows]
        if lays[0] is None or lays[1] is None or not (lays[0].complete and lays[1].complete) or plan.sec1_bytes <= 0:
            return
        self._history = {"n": src.n, "rows": rows, "keep": [lays[0].num_cols - 2, lays[1].num_cols - 2],
                         "row_len": max(1.0, plan.sec1_bytes / max(1, rows[0] + _HEADER_LINES + 1)),
                         "row_len2": max(1.0, (src.n - plan.sec1_bytes) / max(1, rows[1] + _HEADER_LINES + 1)),
                         "overhang": (self._history or {}).get("overhang")}

    def load_many(self, sources, names=None, to_host: bool = True, host_slots: int = 2, return_exceptions: bool = False):
        """Pipelined batch load: yields one ViconNexusData per source, in order.

        sources: an iterable of pinned uint8 CPU tensors (or anything load_bytes accepts;
        non-pinned inputs are staged through a fresh pinned buffer).  While file i is scanned and
        parsed on the compute stream, file i+1 is copied host->device on a copy stream and the
        arrays of file i-1 travel device->host on a third stream (PCIe is full duplex).  With
        to_host the section blocks are copied into pinned buffers borrowed from the loader's pool: a
        buffer belongs to the result it was filled for (its `.df` / `.host()` arrays stay valid for as
        long as anything refers to them) and returns to the pool when the last such array is gone, so a
        caller that lets results go recycles two or three buffers and one that keeps them all gets a
        fresh buffer per file (`host_slots` is accepted for compatibility and ignored).  An error in a file is
        raised when that file is reached - or, with return_exceptions, yielded in its place so
        that the rest of the batch still loads."""
        torch = self.torch
        src_iter = iter(sources)
        name_iter = iter(names) if names is not None else None
        s_comp, _ = self._stream_ptr()
        # the side streams live with the loader: the caching allocator keeps a pool per stream, so fresh
        # streams on every call would strand the previous call's blocks
        if self._side_streams is None:
            self._side_streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        s_copy, s_d2h = self._side_streams
        counter = [0]

        def stage():
            try:
                src = next(src_iter)
            except StopIteration:
                return None
            i = counter[0]
            counter[0] += 1
            name = next(name_iter) if name_iter is not None else f"<bytes {i}>"
            if isinstance(src, Exception):
                return (name, src)
            if isinstance(src, torch.Tensor) and not src.is_cuda and src.is_pinned():
                pinned, n = src, int(src.numel())
            else:
                host = np.frombuffer(src, dtype=np.uint8) if not isinstance(src, (np.ndarray, torch.Tensor)) else src
                host = host.numpy() if isinstance(host, torch.Tensor) else host
                n = int(host.shape[0])
                pinned = torch.empty(_pad16(n), dtype=torch.uint8, pin_memory=True)
                pinned.numpy()[:n] = host
            with torch.cuda.stream(s_copy):
                d_bytes = torch.empty(_pad16(n), dtype=torch.uint8, device=self.device)
                d_bytes[:n].copy_(pinned[:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_copy)
            return (name, d_bytes, n, ev, pinned)

        pipes = self._pipes()

        def submit(staged):
            """The single-pass kernel of a staged file, queued on a pipeline stream behind its host->device copy - not
            behind whatever the caller has queued on the compute stream for the files before it."""
            if staged is None or len(staged) == 2:
                return None
            name, d_bytes, n, ev, pinned = staged
            src = _Source(d_bytes, n, pinned.numpy()[:n])
            ticket = None
            try:
                with self._lock, torch.cuda.device(self.device):
                    if self._want_fused(n):
                        pipe = pipes[counter[0] & 1]
                        pipe.wait_event(ev)
                        d_bytes.record_stream(pipe)
                        ticket = self._submit_fused(src, name, stream=pipe)
            except Exception:  # noqa: BLE001 - the two-pass path below raises what there is to raise
                ticket = None
            return src, ticket

        i = 0
        nxt = stage()
        nxt_sub = submit(nxt)
        while nxt is not None:
            cur, cur_sub = nxt, nxt_sub
            nxt = stage()  # the next file's host->device copy and its kernel are in flight while this one is finished
            nxt_sub = submit(nxt)
            try:
                if len(cur) == 2:
                    raise cur[1]
                name, d_bytes, n, ev, pinned = cur
                src, ticket = cur_sub
                with self._lock, torch.cuda.device(self.device):
                    data = self._finish_fused(ticket) if ticket is not None else None
                    if data is not None:
                        self.stats["fused"] += 1
                        self._fused_backoff = 0
                        for blk in data.blocks:  # allocated on a pipeline stream, used on the compute stream
                            if blk.tensor is not None:
                                blk.tensor.record_stream(s_comp)
                    else:
                        s_comp.wait_event(ev)
                        d_bytes.record_stream(s_comp)
                        self.stats["two_pass"] += 1
                        data = self._run_two_pass(src, name)
            except Exception as exc:  # noqa: BLE001
                if not return_exceptions:
                    raise
                yield exc
                i += 1
                continue
            if to_host:
                done = torch.cuda.Event()
                done.record(s_comp)
                s_d2h.wait_event(done)
                for blk in data.blocks:
                    blk.tensor.record_stream(s_d2h)
                    blk.prefetch_host(s_d2h, pool=self._host_pool)
            yield data
            i += 1
        s_d2h.synchronize()

    def load_files(self, csv_filenames, to_host: bool = True, host_slots: int = 2, read_ahead: int = 2):
        """Batch front-end: yields (filename, ViconNexusData) for every
        file, in order.  A reader thread fills a ring of pinned buffers `read_ahead` files ahead of the
        GPU; the GPU side is `load_many` (H2D, scan/parse and D2H of consecutive files overlap).
        A file that fails yields (filename, exception) instead of stopping the batch."""
        import queue
        import threading

        torch = self.torch
        names = [str(f) for f in csv_filenames]
        # files in use at once: one being parsed, one staged, `read_ahead` queued, one being read
        ring = [None] * (read_ahead + 3)
        q = queue.Queue(maxsize=read_ahead)

        stop = threading.Event()  # set when the consumer goes away before the last file

        def hand_over(item) -> bool:
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.2)
                    return True
                except queue.Full:
                    continue
            return False

        def reader():
            for i, name in enumerate(names):
                if stop.is_set():
                    return
                try:
                    size = os.path.getsize(name)
                    slot = i % len(ring)
                    buf = ring[slot]
                    if buf is None or buf.numel() < _pad16(size):
                        buf = torch.empty(max(_pad16(size), 1 << 20), dtype=torch.uint8, pin_memory=True)
                        ring[slot] = buf
                    read_file_into(name, buf.numpy(), size)
                    item = (name, buf[:size])
                except Exception as exc:  # noqa: BLE001 - reported with the file it belongs to
                    item = (name, exc)
                if not hand_over(item):
                    return
            hand_over(None)

        threading.Thread(target=reader, daemon=True, name="ms-load-files").start()

        def sources():
            while True:
                item = q.get()
                if item is None:
                    return
                yield item

        order = []

        def feed():
            for name, src in sources():
                order.append(name)
                yield src

        k = 0
        try:
            for result in self.load_many(feed(), names=_LazyNames(order), to_host=to_host, host_slots=host_slots,
                                         return_exceptions=True):
                yield order[k], result
                k += 1
        finally:
            stop.set()  # a consumer that stops early must not leave the reader blocked on a full queue


class _LazyNames:
    """Iterator over a list that is still being appended to (names follow the sources)."""

    def __init__(self, items):
        self.items = items
        self.k = 0

    def __iter__(self):
        return self

    def __next__(self):
        name = self.items[self.k]
        self.k += 1
        return name


# ---- planning: which rows are what ------------------------------------------------------------------
class _Plan:
    def __init__(self):
        self.layouts: List[Optional[SectionLayout]] = [None, None]
        self.data_rows: List[Tuple[int, int]] = [(0, 0), (0, 0)]
        self.deferred_error: Optional[Exception] = None  # raised if the data rows before it are clean
        self.sec1_bytes = 0  # bytes up to and including the blank row that closes the first section


def _feed_header(src: _Source, offset: int, first_row: int, n_rows_total: int, machine: HeaderMachine, name: str):
    """Feeds up to five header rows starting at csv row `first_row` / byte `offset`.
    Returns (error or None, header_bytes)."""
    want = min(_HEADER_LINES, max(0, n_rows_total - first_row))
    if want == 0:
        return None, b""
    chunk, n_lines = src.take_lines(offset, want)
    cached = _header_cache.get((chunk, machine.expected, want)) if want == _HEADER_LINES else None
    if cached is not None:
        # identical header bytes (the usual case in a batch of trials): reuse the parsed layout
        machine.layout = cached  # layouts are read-only once parsed
        machine.next_line = None
        return None, chunk
    rows, consumed = rows_from_bytes(chunk, want)
    if consumed != len(rows) or len(rows) != min(want, n_lines):
        raise NotImplementedError(
            f"{name}: a quoted field spans several lines in the header of row {first_row + 1}; "
            "not supported by the CUDA loader"
        )
    for i, row in enumerate(rows):
        try:
            machine.feed(row)
        except Exception as exc:  # noqa: BLE001 - the reference wraps everything (load_csv.py:131)
            return wrap_error(first_row + i + 1, name, exc), chunk
    if machine.done and len(_header_cache) < 64:
        _header_cache[(chunk, machine.expected, want)] = machine.layout
    return None, chunk


def _first_lines(buf: bytes, k: int, at_eof: bool) -> Optional[bytes]:
    """The first k physical lines of `buf` (which ends at the end of the file when at_eof), or None when buf
    does not hold them completely."""
    ends = []
    for m in _TERMINATOR.finditer(buf):
        if m.end() == len(buf) and buf[-1:] == b"\r" and not at_eof:
            break  # may be half of a "\r\n"
        ends.append(m.end())
        if len(ends) == k:
            return buf[: ends[-1]]
    if at_eof and len(ends) == k - 1 and len(buf) > (ends[-1] if ends else 0):
        return buf  # the last line is unterminated
    return None


class _QuotedData(Exception):
    """The data rows contain quote characters: the plain scan's delimiters cannot be trusted."""


def _plan(src: _Source, summary, name: str, quoted: bool = False) -> _Plan:
    plan = _Plan()
    n_rows = int(summary.n_rows)
    if summary.flags & nat.MS_SCAN_BLANK_OVERFLOW:
        raise NotImplementedError(f"{name}: more blank rows than the CUDA loader can order ({summary.n_blank_rows})")
    blanks = [(int(summary.blank_row[i]), int(summary.blank_end[i])) for i in range(int(summary.n_reported))]
    all_reported = int(summary.n_reported) == int(summary.n_blank_rows)

    def first_blank_from(row: int):
        for b in blanks:
            if b[0] >= row:
                return b
        if not all_reported:
            raise NotImplementedError(f"{name}: too many blank rows before the section separator")
        return None

    if summary.flags & nat.MS_SCAN_HAS_HIGH_BYTES and src.host is not None:
        # open(filename) decodes the whole file; undecodable bytes raise before any parsing
        # of the chunk they are in (UnicodeDecodeError, never wrapped: load_csv.py:128)
        src.host.tobytes().decode(file_encoding())

    # ---- section 1 header
    m1 = HeaderMachine(SectionType.FORCES_EMG)
    err, hdr1 = _feed_header(src, 0, 0, n_rows, m1, name)
    header_quotes = hdr1.count(b'"')
    if err is not None:
        raise err  # lines 1-5: nothing can precede it
    plan.layouts[0] = m1.layout
    if not m1.done:
        return plan  # file ends inside the first header
    _check_supported(m1.layout, name)
    b1 = first_blank_from(_HEADER_LINES)
    end1 = b1[0] if b1 is not None else n_rows
    plan.data_rows[0] = (_HEADER_LINES, end1)
    plan.sec1_bytes = b1[1] + 1 if b1 is not None else 0
    if b1 is None:
        _check_quotes(summary, header_quotes, quoted)
        return plan

    # ---- section 2 header
    m2 = HeaderMachine(SectionType.TRAJECTORIES)
    err, hdr2 = _feed_header(src, b1[1] + 1, b1[0] + 1, n_rows, m2, name)
    header_quotes += hdr2.count(b'"')
    _check_quotes(summary, header_quotes, quoted)
    if err is not None:
        plan.deferred_error = err
        return plan
    plan.layouts[1] = m2.layout
    if not m2.done:
        return plan
    _check_supported(m2.layout, name)
    first2 = b1[0] + 1 + _HEADER_LINES
    b2 = first_blank_from(first2)
    end2 = b2[0] if b2 is not None else n_rows
    plan.data_rows[1] = (first2, end2)

    # ---- anything after the second blank row is an error in the reference (Appendix C)
    if b2 is not None and b2[0] + 1 < n_rows:
        m3 = HeaderMachine(None)
        err, _ = _feed_header(src, b2[1] + 1, b2[0] + 1, min(n_rows, b2[0] + 2), m3, name)
        if err is None:  # pragma: no cover - every row raises in this state
            raise AssertionError("row after the second blank row did not raise")
        plan.deferred_error = err
    return plan


def _check_supported(lay: SectionLayout, name: str):
    if lay.num_cols < 3:
        raise NotImplementedError(f"{name}: a section header with fewer than 3 columns is not supported")


def _check_quotes(summary, header_quotes: int, quoted: bool):
    if not quoted and int(summary.n_quotes) > header_quotes:
        raise _QuotedData()


def _raise_device_error(src: _Source, plan: _Plan, key: int, name: str, ws):
    pos, kind = key >> 3, key & 7
    if kind == nat.MS_ERR_KIND_ROW_TOO_LONG:
        raise NotImplementedError(
            f"{name}: a CSV row longer than {nat.MS_MAX_ROW_BYTES} bytes near byte {pos} is not supported"
        )
    row_index, start, stop = src.locate_row(ws, pos)
    rows, _ = rows_from_bytes(src.fetch(start, stop - start), 1)
    num_cols = 0
    for lay, (r0, r1) in zip(plan.layouts, plan.data_rows):
        if lay is not None and r0 <= row_index < r1:
            num_cols = lay.num_cols
    try:
        replay_data_row(rows[0] if rows else [], num_cols)
    except Exception as exc:  # noqa: BLE001
        raise wrap_error(row_index + 1, name, exc) from exc
    if kind == nat.MS_ERR_KIND_NON_ASCII:
        raise NotImplementedError(
            f"{name}: line {row_index + 1} has a numeric field of more than 128 characters with non-ASCII ones among them; "
            "CPython accepts it but the CUDA loader does not"
        )
    raise AssertionError(f"{name}: device rejected a field on line {row_index + 1} that float() accepts")


def _build(plan: _Plan, blocks) -> ViconNexusData:
    """Builder.build (user_data.py:310-433)."""
    lay1, lay2 = plan.layouts
    freq1 = lay1.frequency if lay1 is not None else None
    freq2 = lay2.frequency if lay2 is not None else None
    num_frames = max(0, plan.data_rows[1][1] - plan.data_rows[1][0])
    sampling = SamplingFreq(freq1, freq2, num_frames)
    trackers = (ForcesEMGFrameTracker(sampling), TrajFrameTracker(sampling))
    by_type = {}
    for lay, block, tracker in zip((lay1, lay2), blocks, trackers):
        if lay is None:
            continue
        if lay.devices and not lay.complete:
            # the file ended before the units line: DeviceData(units=None) -> tuple(None)
            # (user_data.py:710)
            raise TypeError("'NoneType' object is not iterable")
        for dev in lay.devices:
            data = DeviceData(
                dev.name, dev.device_type, dev.units, tracker, None,
                block=block, first_channel=dev.first_col - 2, coords=dev.coords,
            )
            by_type.setdefault(dev.device_type, []).append(data)
    emgs = by_type.get(DeviceType.EMG, [])
    if len(emgs) != 1:
        raise ValueError(f"found {len(emgs)} EMG devices - expected one")
    for needed in (DeviceType.FORCE_PLATE, DeviceType.TRAJECTORY_MARKER):
        if needed not in by_type:
            raise KeyError(needed)
    return ViconNexusData(
        forcepl=by_type[DeviceType.FORCE_PLATE], emg=emgs[0], traj=by_type[DeviceType.TRAJECTORY_MARKER]
    )


# ---- module-level convenience ---------------------------------------------------------------------------
_default_loaders = {}


def _default_loader() -> ViconLoader:
    torch = _torch()
    dev = torch.cuda.current_device()
    if dev not in _default_loaders:
        _default_loaders[dev] = ViconLoader(f"cuda:{dev}")
    return _default_loaders[dev]


def load_vicon_file(csv_filename) -> ViconNexusData:
    """Load data from a Vicon Nexus CSV file (drop-in for load_csv.py:96-135)."""
    return _default_loader().load_file(csv_filename)


def load_vicon_bytes(data, name: str = "<bytes>") -> ViconNexusData:
    """Same as load_vicon_file for CSV bytes already in memory (host or CUDA uint8)."""
    return _default_loader().load_bytes(data, name)
