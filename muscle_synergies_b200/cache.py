"""Binary cache of parsed trials (the reference has no equivalent).

A Vicon CSV is text for 48 M doubles; once parsed, the channel-major float64 blocks the loader
produced can be kept next to it and the next analysis starts from a plain read + H2D copy:

    offset 0    b"MSB200TC"  uint32 version  uint32 header_bytes            (little-endian)
    offset 16   JSON header: source, sampling frequencies, devices (name, type, units, coords, section,
                first_channel), sections (n_channels, n_rows, offset, crc32)
    4 KiB aligned: one block per section, float64 little-endian, [channel][row] - the layout of
                SectionBlock.tensor and of the reference's DataFrame blocks

`save_trial` / `load_trial` convert between that file and ViconNexusData; `load_vicon_file_cached`
is `load_vicon_file` that consults / fills a cache file keyed on the CSV's size and mtime.  Values
round-trip bit-exactly (NaN payloads included: raw bytes are stored).
"""
import json
import os
import struct
import zlib
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat
from .vicon_data.data_model import (
    DeviceData,
    ForcesEMGFrameTracker,
    SectionBlock,
    TrajFrameTracker,
    ViconNexusData,
)
from .vicon_data.definitions import DeviceType, SamplingFreq

MAGIC = b"MSB200TC"
VERSION = 1
ALIGN = 4096


class CacheError(RuntimeError):
    """The file is not a trial cache, is truncated, or fails its checksum."""


def _aligned(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


def _raw(block: np.ndarray) -> np.ndarray:
    return block.reshape(-1).view(np.uint8)  # works for empty blocks too, unlike memoryview.cast


# ---- file level (host only) ---------------------------------------------------------------------------
def write_trial_file(path, header: dict, blocks: Sequence[np.ndarray]) -> None:
    """header: everything but the `sections` list, which is derived from `blocks`
    ((n_channels, n_rows) float64 arrays, C order).  Written to a temporary name and renamed."""
    blocks = [np.ascontiguousarray(b, dtype="<f8") for b in blocks]
    sections = [{"n_channels": int(b.shape[0]), "n_rows": int(b.shape[1]), "offset": 0,
                 "crc32": zlib.crc32(_raw(b)) & 0xFFFFFFFF} for b in blocks]
    head = dict(header, version=VERSION, sections=sections)
    # offsets depend on the header length, which depends on the offsets' digits: fix the width first
    for _ in range(3):
        raw = json.dumps(head, sort_keys=True).encode("utf-8")
        pos = _aligned(16 + len(raw))
        changed = False
        for sec, b in zip(sections, blocks):
            changed |= sec["offset"] != pos
            sec["offset"] = pos
            pos = _aligned(pos + b.nbytes)
        if not changed:
            break
    raw = json.dumps(head, sort_keys=True).encode("utf-8")
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(MAGIC + struct.pack("<II", VERSION, len(raw)) + raw)
        for sec, b in zip(sections, blocks):
            f.seek(sec["offset"])
            f.write(_raw(b))
        f.truncate(_aligned(f.tell()))
    os.replace(tmp, path)


def read_trial_header(path) -> dict:
    with open(path, "rb") as f:
        return _read_header(f, path)


def _read_header(f, path) -> dict:
    fixed = f.read(16)
    if len(fixed) < 16 or fixed[:8] != MAGIC:
        raise CacheError(f"{path}: not a muscle_synergies_b200 trial cache")
    version, n = struct.unpack("<II", fixed[8:])
    if version != VERSION:
        raise CacheError(f"{path}: cache format version {version}, this build reads {VERSION}")
    raw = f.read(n)
    if len(raw) < n:
        raise CacheError(f"{path}: truncated header")
    try:
        return json.loads(raw.decode("utf-8"))
    except ValueError as exc:
        raise CacheError(f"{path}: corrupt header") from exc


def read_trial_file(path, verify: bool = True, into=None) -> Tuple[dict, List[np.ndarray]]:
    """(header, blocks).  `into(section_index, n_bytes)` may supply the destination buffers
    (e.g. pinned memory); the default is fresh numpy arrays."""
    with open(path, "rb") as f:
        head = _read_header(f, path)
        blocks = []
        for i, sec in enumerate(head["sections"]):
            shape = (sec["n_channels"], sec["n_rows"])
            nbytes = 8 * shape[0] * shape[1]
            buf = into(i, nbytes) if into is not None else np.empty(nbytes, dtype=np.uint8)
            f.seek(sec["offset"])
            got = f.readinto(memoryview(buf)[:nbytes]) if nbytes else 0
            if got != nbytes:
                raise CacheError(f"{path}: truncated section {i}")
            if verify and (zlib.crc32(memoryview(buf)[:nbytes]) & 0xFFFFFFFF) != sec["crc32"]:
                raise CacheError(f"{path}: checksum mismatch in section {i}")
            blocks.append(np.frombuffer(buf, dtype="<f8", count=shape[0] * shape[1]).reshape(shape))
    return head, blocks


# ---- ViconNexusData level ---------------------------------------------------------------------------------
def _describe(data: ViconNexusData, source: Optional[str]) -> Tuple[dict, list]:
    devices = list(data.forcepl) + [data.emg] + list(data.traj)
    section_blocks = []
    for dev in devices:
        if dev._block is None:
            raise ValueError("only trials produced by the CUDA loader can be cached")
        if not any(dev._block is b for b in section_blocks):
            section_blocks.append(dev._block)
    head = {
        "source": source,
        "sampling": {"forces_emg": data.emg.sampling_frequency, "traj": data.traj[0].sampling_frequency,
                     "num_frames": data.traj[0]._frame_tracker.num_frames},
        "devices": [{"name": d.name, "type": d.dev_type.name, "units": list(d.units), "coords": d.columns,
                     "section": next(i for i, b in enumerate(section_blocks) if b is d._block),
                     "first_channel": d._first_channel} for d in devices],
    }
    return head, section_blocks


def save_trial(data: ViconNexusData, path, source: Optional[dict] = None) -> None:
    """Writes the parsed trial to `path` (one device->host copy per section unless already made)."""
    data.check()  # a trial loaded with defer_check=True must not be cached before its rows are known good
    head, section_blocks = _describe(data, source)
    write_trial_file(path, head, [b.host() for b in section_blocks])


def load_trial(path, device=None, verify: bool = False) -> ViconNexusData:
    """ViconNexusData from a cache file: blocks are read into pinned memory and copied to the GPU;
    the pinned copy doubles as the host side of `.df`, so no device->host copy ever happens."""
    import torch

    if not torch.cuda.is_available():
        raise nat.NativeError("load_trial needs a CUDA device: ViconNexusData lives in HBM; there is no CPU fallback")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    pinned = {}

    def into(i, nbytes):
        pinned[i] = torch.empty(max(nbytes, 8), dtype=torch.uint8, pin_memory=True)
        return pinned[i].numpy()

    head, host_blocks = read_trial_file(path, verify=verify, into=into)
    blocks = []
    for i, hb in enumerate(host_blocks):
        t = pinned[i][: hb.nbytes].view(torch.float64).view(hb.shape)
        blk = SectionBlock(t.to(dev, non_blocking=True), hb.shape[1])
        blk._host = hb
        blk._pinned = pinned[i]  # keeps the host copy alive
        blocks.append(blk)
    torch.cuda.current_stream(dev).synchronize()
    samp = head["sampling"]
    sampling = SamplingFreq(samp["forces_emg"], samp["traj"], samp["num_frames"])
    trackers = (ForcesEMGFrameTracker(sampling), TrajFrameTracker(sampling))
    by_type = {}
    for d in head["devices"]:
        dtype = DeviceType[d["type"]]
        tracker = trackers[1] if dtype is DeviceType.TRAJECTORY_MARKER else trackers[0]
        dd = DeviceData(d["name"], dtype, d["units"], tracker, None, block=blocks[d["section"]],
                        first_channel=d["first_channel"], coords=d["coords"])
        by_type.setdefault(dtype, []).append(dd)
    data = ViconNexusData(forcepl=by_type[DeviceType.FORCE_PLATE], emg=by_type[DeviceType.EMG][0],
                          traj=by_type[DeviceType.TRAJECTORY_MARKER])
    data.blocks = blocks
    return data


def cache_path_for(csv_filename, cache_dir=None) -> str:
    csv_filename = os.fspath(csv_filename)
    if cache_dir is None:
        return csv_filename + ".msb200"
    # same-named trials of different directories must not share a cache file: the absolute path is part of the name
    import hashlib

    tag = hashlib.sha1(os.path.abspath(csv_filename).encode()).hexdigest()[:12]
    return os.path.join(os.fspath(cache_dir), f"{os.path.basename(csv_filename)}.{tag}.msb200")


def load_vicon_file_cached(csv_filename, cache_dir=None, loader=None, verify: bool = True) -> ViconNexusData:
    """`load_vicon_file` with a binary cache beside the CSV (or in `cache_dir`): the cache is used when it
    records this CSV's absolute path, current size and mtime, otherwise the CSV is parsed and the cache rewritten.
    `verify` checks the CRC32 of every block read from the cache (a damaged cache is rebuilt from the CSV).  A
    cache that cannot be written (read-only directory) is not an error: the parsed data is returned all the same."""
    from .vicon_data.loader import _default_loader

    st = os.stat(csv_filename)
    stamp = {"path": os.path.abspath(os.fspath(csv_filename)), "size": st.st_size, "mtime_ns": st.st_mtime_ns}
    cpath = cache_path_for(csv_filename, cache_dir)
    if os.path.exists(cpath):
        try:
            src = read_trial_header(cpath).get("source") or {}
            if (src.get("path") == stamp["path"] and src.get("size") == stamp["size"]
                    and src.get("mtime_ns") == stamp["mtime_ns"]):
                return load_trial(cpath, device=loader.device if loader is not None else None, verify=verify)
        except CacheError:
            pass  # stale or damaged cache: fall through to the CSV, then rewrite it
    data = (loader if loader is not None else _default_loader()).load_file(csv_filename)
    try:
        save_trial(data, cpath, source=stamp)
    except OSError:
        pass  # the cache is an optimisation; the data is what was asked for
    return data
