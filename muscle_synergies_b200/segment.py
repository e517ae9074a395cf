"""Trial windowing on the GPU: drop-in for the compute part of the reference's project/segment.py.

    Phase / Trecho / Cycle          segment.py:21-87
    reactions                       segment.py:118-121
    Segmenter                       segment.py:124-298   (same methods and argument forms)
    _transition_indices             segment.py:667-755   -> CUDA (ms_find_transitions)
    _organize_transitions           segment.py:787-917   -> 40 ints -> 32 (frame, subframe) slices, host

plus `Segmenter.cut`, which gathers any set of windows from a device's channel-major block
with one kernel launch (ms_cut_windows) instead of one `df.iloc[...]` per window.

The plotting half of the reference file (SegmentPlotter, :301-664) is out of scope.

Behavioural note: when the signal holds fewer than `num_segments`
alternations the reference silently re-appends a stale index (:745-752); this implementation
raises ValueError instead.  Parity is claimed only when all transitions exist.
"""
import ctypes
import os
import threading
from collections import OrderedDict
from enum import Enum, auto
from typing import List, Mapping, Optional, Sequence, Tuple, Union

from . import _native as nat
from .vicon_data.data_model import DeviceData, FrameSubfr, ViconNexusData
from .vicon_data.definitions import DeviceType  # noqa: F401  (re-exported like the reference)


class Phase(Enum):
    DAA = "DAA"
    AS = "AS"
    DAE = "DAE"
    BL = "BL"

    @staticmethod
    def from_str(phase: str) -> "Phase":
        return {"DAA": Phase.DAA, "DAE": Phase.DAE, "AS": Phase.AS, "BL": Phase.BL}[phase.upper()]


class Trecho(Enum):
    FIRST = auto()
    SECOND = auto()
    THIRD = auto()
    FOURTH = auto()


class Cycle(Enum):
    FIRST = auto()
    SECOND = auto()


Segments = Mapping[Trecho, Mapping[Cycle, Mapping[Phase, slice]]]
PhaseRef = Union[Phase, int, str]


def reactions(vicon_nexus_data: ViconNexusData):
    """Vertical ground reaction of the two force plates as pandas Series (segment.py:118-121)."""
    left_fp, right_fp = vicon_nexus_data.forcepl
    return left_fp.df["Fz"], right_fp.df["Fz"]


def _fz_tensor(dev: DeviceData):
    """The Fz channel of a force plate as a CUDA view (what df["Fz"] selects)."""
    coords = list(dev._coords)
    if coords.count("Fz") != 1:
        raise KeyError("Fz")
    return dev.tensor[coords.index("Fz")]


_pinned_pool = threading.local()  # per thread: numel -> a few pinned buffers, handed out round-robin


def _pinned_like(t):
    """A pinned host buffer shaped like `t`; its contents are read right after the copy, so a small ring
    per thread is enough."""
    import torch

    n = int(t.numel())
    if not hasattr(_pinned_pool, "rings"):
        _pinned_pool.rings = {}
    ring = _pinned_pool.rings.setdefault((n, t.dtype), [[], 0])
    if len(ring[0]) < 4:
        ring[0].append(torch.empty(n, dtype=t.dtype, pin_memory=True))
        return ring[0][-1]
    ring[1] = (ring[1] + 1) % len(ring[0])
    return ring[0][ring[1]]


class _TransitionSearch:
    """ms_find_transitions queued on the current stream; `finish()` waits and reads the result.
    `extra_words` int64 slots follow the results in the same buffer (`extra`), so that work queued
    behind the search can return its own small outputs in the same device->host copy."""

    @nat.on_tensor_device
    def __init__(self, left_fz, right_fz, min_phase_size: int, num_segments: int, extra_words: int = 0, plans=()):
        """plans: (divisor, n_rows, n_channels) per device whose 32 phase windows the same launch plans; plan i's
        starts / stops / offsets land in extra[i * _PLAN_WORDS : (i + 1) * _PLAN_WORDS]."""
        import torch

        lib = nat.lib()
        if not (left_fz.is_cuda and right_fz.is_cuda):
            raise nat.NativeError("transition_indices needs CUDA tensors; there is no CPU fallback")
        left_fz = left_fz.contiguous()
        right_fz = right_fz.contiguous()
        if left_fz.dtype != torch.float64 or right_fz.dtype != torch.float64 or left_fz.shape != right_fz.shape:
            raise ValueError("left/right reactions must be float64 vectors of equal length")
        n = int(left_fz.numel())
        dev = left_fz.device
        self.min_phase_size, self.num_segments = min_phase_size, num_segments
        self.want = want = num_segments if num_segments > 0 else max(1, n)
        work = torch.empty(int(lib.ms_transitions_workspace_bytes(n)), dtype=torch.uint8, device=dev)
        # one result buffer -> one device->host copy: [indices (int64) | loaded flags, count (int32) | extra]
        self.tail_words = (want + 2 + 1) // 2
        self.res = torch.empty(want + self.tail_words + extra_words, dtype=torch.int64, device=dev)
        self.d_transitions = self.res[:want]
        tail = self.res[want : want + self.tail_words].view(torch.int32)
        loaded, self.d_found = tail[:want], tail[want : want + 1]
        self.extra = self.res[want + self.tail_words :]
        self.stream = torch.cuda.current_stream(dev)
        plans = list(plans)
        c_plans = (nat.WindowPlan * max(1, len(plans)))()
        for i, (divisor, n_rows, n_ch) in enumerate(plans):
            base = self.extra.data_ptr() + 8 * i * _PLAN_WORDS
            c_plans[i] = nat.WindowPlan(int(divisor), int(n_rows), int(n_ch), 0, base, base + 8 * 32, base + 8 * 64)
        nat.check(
            lib.ms_segment_trial(
                left_fz.data_ptr(), right_fz.data_ptr(), n, int(min_phase_size), int(want), work.data_ptr(),
                self.d_transitions.data_ptr(), loaded.data_ptr(), self.d_found.data_ptr(), c_plans, len(plans),
                ctypes.c_void_p(self.stream.cuda_stream),
            ),
            "ms_segment_trial",
        )
        self._keep = (left_fz, right_fz, work)

    def finish(self):
        """(indices, loaded flags, extra words as a host int64 tensor); raises like the reference when
        fewer than num_segments transitions exist."""
        import torch

        want = self.want
        # into pinned memory: a pageable destination goes through the driver's staging buffer and costs
        # a few tens of microseconds on the critical path of every trial
        host = _pinned_like(self.res)
        host.copy_(self.res, non_blocking=True)
        self.stream.synchronize()
        tail = host[want : want + self.tail_words].view(torch.int32)
        k = int(tail[want].item())
        if self.num_segments > 0 and k < self.num_segments:
            legs = 1 if k % 2 == 0 else 2
            raise ValueError(
                f"no phase found with {self.min_phase_size} adjacent measurements with {legs} leg(s) with a nonzero reaction"
                f" (found {k} of {self.num_segments} transitions)"
            )
        return host[:k].tolist(), tail[:k].tolist(), host[want + self.tail_words :]


def transition_indices(left_fz, right_fz, min_phase_size: int = 10, num_segments: int = 40, with_loaded=False):
    """_transition_indices (segment.py:667-755) on two CUDA float64 vectors.

    Returns a list of python ints (and, with_loaded, which plates are loaded at each index:
    bit 0 left, bit 1 right).  Raises ValueError when fewer than num_segments exist
    (num_segments=0 keeps the reference meaning "as many as there are").
    """
    idx, loaded_host, _ = _TransitionSearch(left_fz, right_fz, min_phase_size, num_segments).finish()
    if with_loaded:
        return idx, loaded_host
    return idx


def organize_transitions(to_framesubfr, transitions: Sequence[int], loaded: Sequence[int], to_framesubfr_many=None) -> Segments:
    """_organize_transitions (segment.py:787-917): 40 indices -> 4 trechos x 2 cycles x 4 phases.
    `to_framesubfr_many`, when given, converts all 64 bounds in one call (same checks, same results)."""

    def single_leg_phase_type(pos: int) -> Phase:
        flags = loaded[pos]
        if flags == 3 or flags == 0:
            raise ValueError(
                "expected index corresponding to a phase in which there is ground reaction for exactly one leg."
            )
        return Phase.BL if flags & 1 else Phase.AS

    def phase_seq(second_phase: Phase, trecho: Trecho) -> List[Phase]:
        if trecho in (Trecho.FIRST, Trecho.THIRD):
            if second_phase is Phase.BL:
                return [Phase.DAA, Phase.BL, Phase.DAE, Phase.AS]
            return [Phase.DAE, Phase.AS, Phase.DAA, Phase.BL]
        if second_phase is Phase.BL:
            return [Phase.DAE, Phase.BL, Phase.DAA, Phase.AS]
        return [Phase.DAA, Phase.AS, Phase.DAE, Phase.BL]

    # phase i of a trecho runs from bounds[i] to bounds[i + 1] - 1 (both converted to (frame, subframe))
    all_bounds = [list(transitions[10 * n + 1 : 10 * n + 10]) for n in range(len(Trecho))]
    wanted = [idx for bounds in all_bounds for i in range(8) for idx in (bounds[i], bounds[i + 1] - 1)]
    if to_framesubfr_many is not None:
        converted = iter(to_framesubfr_many(wanted))
    else:
        converted = iter([to_framesubfr(idx) for idx in wanted])

    segments = {}
    for n, trecho in enumerate(Trecho):
        names = phase_seq(single_leg_phase_type(10 * n + 2), trecho)
        slices = [slice(next(converted), next(converted)) for _ in range(8)]
        segments[trecho] = {
            Cycle.FIRST: OrderedDict(zip(names, slices[:4])),
            Cycle.SECOND: OrderedDict(zip(names, slices[4:])),
        }
    return segments


class _WindowViews(Sequence):
    """The gathered windows as a read-only sequence of (n_columns, n_rows_w) CUDA tensors: views of one
    buffer, made when asked for (32 tensor objects per trial are host time the GPU would wait for)."""

    def __init__(self, buffer, n_channels: int, starts, stops, offsets):
        self._buffer, self._n_ch = buffer, n_channels
        self._starts, self._stops, self._offsets = starts, stops, offsets
        self._views = {}

    def __len__(self):
        return len(self._starts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        if i not in self._views:
            self._views[i] = self._buffer[self._offsets[i] : self._offsets[i + 1]].view(
                self._n_ch, self._stops[i] - self._starts[i])
        return self._views[i]


_PLAN_WORDS = 32 + 32 + 33  # starts, stops, offsets of the 32 phase windows of one device
# Cross-check every device-planned window against the host's index arithmetic (the tests switch it on; it costs
# 64 conversions per device and trial on the critical path of a step, for an identity the tests establish)
VERIFY_DEVICE_PLAN = os.environ.get("MS_B200_VERIFY_PLAN") == "1"


class PendingSegmenter:
    """A Segmenter whose GPU work is queued (`Segmenter.begin`): `finish()` waits for it and completes the object."""

    def __init__(self, seg: "Segmenter", begun):
        self._seg, self._begun = seg, begun

    def finish(self) -> "Segmenter":
        if self._begun is not None:
            begun, self._begun = self._begun, None
            self._seg._complete(*begun)
        return self._seg


class Segmenter:
    """Segments a trial into trechos, cycles and phases from the two force plates."""

    def __init__(self, data: ViconNexusData, min_phase_size: int = 10, num_segments: int = 40, cut_phases_of=()):
        """`cut_phases_of` (extension): devices of `data` whose 32 phase windows are gathered on the GPU
        in the same submission as the transition search (no host round trip in between); read them
        with `phase_cuts(device)`.  A trial loaded with `defer_check=True` is checked here."""
        self._complete(*self._begin(data, min_phase_size, num_segments, cut_phases_of))

    @classmethod
    def begin(cls, data: ViconNexusData, min_phase_size: int = 10, num_segments: int = 40, cut_phases_of=()) -> "PendingSegmenter":
        """Queues the transition search and the window gathers of `Segmenter(data, ...)` on the current stream and
        returns at once (extension): `.finish()` waits for them and gives the Segmenter - what the constructor would
        have raised is raised there.  A caller that works through a stream of trials starts trial i + 1 before it
        finishes trial i, and never stands waiting for a search."""
        seg = cls.__new__(cls)
        return PendingSegmenter(seg, seg._begin(data, min_phase_size, num_segments, cut_phases_of))

    def _begin(self, data: ViconNexusData, min_phase_size: int, num_segments: int, cut_phases_of):
        left_fp, right_fp = data.forcepl  # exactly two plates, like reactions()
        self._data = data
        precut = list(cut_phases_of)
        if precut and num_segments < 40:
            raise ValueError("cut_phases_of needs the full 40 transitions")
        in_launch = precut[: nat.MS_MAX_WINDOW_PLANS]  # their window plans come out of the search launch itself
        search = _TransitionSearch(_fz_tensor(left_fp), _fz_tensor(right_fp), min_phase_size, num_segments,
                                   extra_words=_PLAN_WORDS * len(precut),
                                   plans=[self._plan_args(dev, left_fp) for dev in in_launch])
        queued = [self._queue_phase_cuts(search, dev, i, left_fp, planned=i < len(in_launch)) for i, dev in enumerate(precut)]
        return search, queued

    def _complete(self, search: "_TransitionSearch", queued):
        data = self._data
        left_fp = data.forcepl[0]
        try:
            self.transitions, self._loaded, extra = search.finish()
        finally:
            check = getattr(data, "check", None)
            if check is not None:
                check()  # a parse error of a deferred load comes first, as it would have at load time
        # The (frame, subframe) slices of all 32 phases are made when first asked for (get_times_of ...): 64
        # conversions that the device-side window plan does not need.  What building them could raise is raised
        # here, where the reference raises it (segment.py:852-879, user_data.py:575-597).
        self._plate = left_fp
        self._segments_made = None
        if len(self.transitions) >= 40:
            for n in range(len(Trecho)):
                if self._loaded[10 * n + 2] in (0, 3):
                    raise ValueError(
                        "expected index corresponding to a phase in which there is ground reaction for exactly one leg."
                    )
            bounds = [self.transitions[j] for n in range(len(Trecho)) for j in (10 * n + 1, 10 * n + 9)]
            if min(bounds) - 0 < 0 or max(bounds) > left_fp._frame_tracker.final_index:
                self._segments  # out of the device's rows: the conversion raises the reference's IndexError
        else:
            self._segments
        self._phase_cuts = {}
        for i, (dev, out) in enumerate(queued):
            self._phase_cuts[id(dev)] = (dev, self._finish_phase_cuts(dev, out, extra[i * _PLAN_WORDS : (i + 1) * _PLAN_WORDS]))

    @property
    def _segments(self):
        if self._segments_made is None:
            self._segments_made = organize_transitions(self._plate.to_framesubfr, self.transitions, self._loaded,
                                                       getattr(self._plate, "to_framesubfr_many", None))
        return self._segments_made

    @staticmethod
    def _plan_args(dev: DeviceData, plate: DeviceData):
        """(divisor, n_rows, n_channels) of ms_window_plan: force-plate sample index -> row of `dev`."""
        src = dev.tensor
        same_section = dev._frame_tracker.per_frame == plate._frame_tracker.per_frame
        return (1 if same_section else plate._frame_tracker.num_subframes), int(src.shape[1]), int(src.shape[0])

    def _queue_phase_cuts(self, search: "_TransitionSearch", dev: DeviceData, slot: int, plate: DeviceData, planned: bool = False):
        import torch

        with torch.cuda.device(dev.tensor.device):
            return self._queue_phase_cuts_on_device(search, dev, slot, plate, planned)

    def _queue_phase_cuts_on_device(self, search: "_TransitionSearch", dev: DeviceData, slot: int, plate: DeviceData, planned: bool):
        """Plans (on the device) and gathers the 32 phase windows of `dev` behind the search."""
        import torch

        lib = nat.lib()
        src = dev.tensor
        n_ch, n_rows = int(src.shape[0]), int(src.shape[1])
        same_section = dev._frame_tracker.per_frame == plate._frame_tracker.per_frame
        divisor = 1 if same_section else plate._frame_tracker.num_subframes
        meta = search.extra[slot * _PLAN_WORDS : (slot + 1) * _PLAN_WORDS]
        starts, stops, offsets = meta[:32], meta[32:64], meta[64:97]
        out = torch.empty(max(1, n_ch * n_rows), dtype=torch.float64, device=src.device)  # upper bound: every row once
        sptr = ctypes.c_void_p(search.stream.cuda_stream)
        if not planned:
            nat.check(lib.ms_plan_phase_windows(search.d_transitions.data_ptr(), search.d_found.data_ptr(), search.want, 0,
                                                divisor, n_rows, n_ch, starts.data_ptr(), stops.data_ptr(),
                                                offsets.data_ptr(), sptr), "ms_plan_phase_windows")
        if n_ch:
            nat.check(lib.ms_cut_windows(src.data_ptr(), int(src.stride(0)), n_ch, starts.data_ptr(), stops.data_ptr(),
                                         offsets.data_ptr(), 32, out.data_ptr(), n_rows, sptr), "ms_cut_windows")
        return dev, out

    def _finish_phase_cuts(self, dev: DeviceData, out, meta):
        meta = meta.tolist()
        starts, stops, offsets = meta[:32], meta[32:64], meta[64:97]
        if VERIFY_DEVICE_PLAN:
            # the device planned the row ranges; the host path (get_times_of -> DeviceData.to_index) must agree
            windows = [w[3] for w in self.all_phase_windows()]
            flat = dev.to_index_many([b for w in windows for b in (w.start, w.stop)])
            n_rows = int(dev.tensor.shape[1])
            for i in range(32):
                a = min(flat[2 * i], n_rows)
                if (starts[i], stops[i]) != (a, max(a, min(flat[2 * i + 1], n_rows))):
                    raise AssertionError("device-planned phase window differs from the host's")
        return _WindowViews(out, int(dev.tensor.shape[0]), starts, stops, offsets)

    def phase_cuts(self, device: DeviceData):
        """The 32 phase windows of `device` (order of `all_phase_windows`) as (n_columns, n_rows_w) CUDA
        tensors: the ones gathered at construction (`cut_phases_of`), else gathered now."""
        hit = self._phase_cuts.get(id(device))
        if hit is not None:
            return hit[1]
        return Segmenter.cut(device, [w[3] for w in self.all_phase_windows()])

    # ---- reference API ------------------------------------------------------------------------
    def ith_phase(self, trecho: Union[Trecho, int], i: int) -> Phase:
        if i not in range(1, 5):
            raise IndexError("i should be a number between 1 and 4")
        trecho = self._parse_trecho(trecho)
        return tuple(self._segments[trecho][Cycle.FIRST].keys())[(i - 1) % 4]

    def get_times_of(self, trecho, cycle=None, phase=None) -> slice:
        trecho, cycle, phase = self._parse_segment_args(trecho, cycle, phase)
        if phase is not None:
            return self._segments[trecho][cycle][phase]
        if cycle is not None:
            return self._times_of_cycle(trecho, cycle)
        first = self._times_of_cycle(trecho, Cycle.FIRST)
        second = self._times_of_cycle(trecho, Cycle.SECOND)
        return slice(first.start, second.stop)

    def _times_of_cycle(self, trecho: Trecho, cycle: Cycle) -> slice:
        phases = tuple(self._segments[trecho][cycle].values())
        return slice(phases[0].start, phases[3].stop)

    def _parse_segment_args(self, trecho, cycle, phase_ref):
        def must_be_omitted(given: bool):
            if given:
                raise ValueError(
                    "the optional arguments should be ommitted if a (trecho, cycle, phase_ref) triple is given"
                )

        if phase_ref is not None and cycle is None:
            raise ValueError("if a phase is given, a cycle should also be")
        given = cycle is not None or phase_ref is not None
        try:
            trecho, cycle, phase_ref = trecho
        except TypeError:
            pass
        except ValueError:
            trecho, cycle = trecho
            must_be_omitted(given)
        else:
            must_be_omitted(given)
        trecho = self._parse_trecho(trecho)
        cycle = self._parse_cycle(cycle)
        return trecho, cycle, self._parse_phase(trecho, phase_ref)

    @staticmethod
    def _parse_trecho(trecho: Union[Trecho, int]) -> Trecho:
        # ints are 1-based positions (the reference's intent; its `trecho in Trecho` test
        # only behaves that way on Python <= 3.7)
        if isinstance(trecho, Trecho):
            return trecho
        return tuple(Trecho)[trecho - 1]

    @staticmethod
    def _parse_cycle(cycle: Optional[Union[Cycle, int]] = None) -> Optional[Cycle]:
        if cycle is None or isinstance(cycle, Cycle):
            return cycle
        return tuple(Cycle)[cycle - 1]

    def _parse_phase(self, trecho: Trecho, phase_ref: Optional[PhaseRef]) -> Optional[Phase]:
        if phase_ref is None or isinstance(phase_ref, Phase):
            return phase_ref
        try:
            return Phase.from_str(phase_ref)
        except (KeyError, AttributeError):
            pass
        return self.ith_phase(trecho, phase_ref)

    # ---- GPU window gather ----------------------------------------------------------------------
    def all_phase_windows(self) -> List[Tuple[Trecho, Cycle, Phase, slice]]:
        out = []
        for trecho, cycles in self._segments.items():
            for cycle, phases in cycles.items():
                for phase, sl in phases.items():
                    out.append((trecho, cycle, phase, sl))
        return out

    @staticmethod
    def cut(device: DeviceData, windows: Sequence[slice]):
        """Rows `device[w]` for every (frame, subframe) slice w, gathered on the GPU.

        Returns a list of (n_columns, n_rows_w) float64 CUDA tensors (channel-major), the
        same rows `device.df.iloc[device.to_index(w)]` selects (exclusive stop)."""
        if all(w.step is None and w.start is not None and w.stop is not None for w in windows):
            flat = device.to_index_many([b for w in windows for b in (w.start, w.stop)])
            return cut_windows(device, [slice(flat[2 * i], flat[2 * i + 1]) for i in range(len(windows))])
        return cut_windows(device, [device.to_index(w) for w in windows])


@nat.on_tensor_device
def cut_windows(device: DeviceData, index_slices: Sequence[slice]):
    """Batched `df.iloc[a:b]` on the device-resident block of one DeviceData."""
    import torch

    lib = nat.lib()
    src = device.tensor
    n_rows = int(src.shape[1])
    n_ch = int(src.shape[0])
    starts, stops = [], []
    for sl in index_slices:
        if sl.step not in (None, 1):
            raise ValueError("cut_windows supports unit-step slices only")
        a, b, _ = sl.indices(n_rows)
        starts.append(a)
        stops.append(max(a, b))
    lens = [b - a for a, b in zip(starts, stops)]
    offsets = [0]
    for ln in lens:
        offsets.append(offsets[-1] + ln * n_ch)
    dev = src.device
    meta = torch.tensor([starts, stops, offsets[:-1]], dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
    out = torch.empty(max(1, offsets[-1]), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    if index_slices and n_ch:
        nat.check(
            lib.ms_cut_windows(
                src.data_ptr(), int(src.stride(0)), n_ch, meta[0].data_ptr(), meta[1].data_ptr(), meta[2].data_ptr(),
                len(lens), out.data_ptr(), max(lens) if lens else 0, ctypes.c_void_p(stream.cuda_stream),
            ),
            "ms_cut_windows",
        )
    return [out[offsets[i] : offsets[i + 1]].view(n_ch, lens[i]) for i in range(len(lens))]
