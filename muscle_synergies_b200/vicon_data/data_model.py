"""What load_vicon_file returns: the reference's user-facing data model over GPU-resident arrays.

API mirror of src/muscle_synergies/vicon_data/user_data.py: ViconNexusData (:42-301),
frame trackers (:483-661), DeviceData (:664-772) - same attribute names, index math and
exceptions.  Differences, all additive:

  * the arrays live in HBM as one channel-major float64 block per CSV section
    (`SectionBlock`); `DeviceData.tensor` is a zero-copy (channels, rows) view of it;
  * `DeviceData.df` is built lazily from a pinned host copy of the block, which has the
    memory layout pandas itself ends up with for the reference (user_data.py:396).
"""
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import pandas as pd

from .definitions import DeviceType, SamplingFreq

FrameSubfr = Tuple[int, int]


class HostLease:
    """A pinned host buffer on loan from a loader's pool.  numpy arrays made from it (and every view of them, the
    blocks of a DataFrame included) keep it alive; when the last one is gone the buffer goes back to the pool.
    Nothing a caller still holds is ever overwritten by a later file."""

    def __init__(self, pinned, pool, shape):
        self.pinned = pinned
        self.pool = pool
        self.event = None
        self.__array_interface__ = {
            "data": (pinned.data_ptr(), False), "shape": tuple(shape), "typestr": "<f8", "strides": None, "version": 3,
        }

    def __del__(self):
        try:
            if self.event is not None:
                self.event.synchronize()  # a copy may still be writing into it
            if self.pool is not None and len(self.pool) < 8:
                self.pool.append(self.pinned)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class SectionBlock:
    """Channel-major float64 block of one CSV section: tensor[channel, row]."""

    def __init__(self, tensor, n_rows: int):
        self.tensor = tensor  # torch.float64, (n_keep, stride) on the GPU, stride >= n_rows
        self.n_rows = n_rows
        self._host = None
        self._pending = None  # (HostLease or pinned tensor, event) of an asynchronous device->host copy

    def _copy_rows(self, dst_ptr: int, stream):
        """The (n_keep, n_rows) part of the block to contiguous host memory: one 2-D DMA copy, whatever the stride."""
        import ctypes

        from .. import _native as nat

        n_keep = int(self.tensor.shape[0])
        stride = int(self.tensor.stride(0)) if n_keep > 1 else self.n_rows
        if n_keep == 0 or self.n_rows == 0:
            return
        nat.check(
            nat.lib().ms_copy_rows_to_host(dst_ptr, self.n_rows * 8, self.tensor.data_ptr(), max(stride, self.n_rows) * 8,
                                           self.n_rows * 8, n_keep, ctypes.c_void_p(stream.cuda_stream)),
            "ms_copy_rows_to_host",
        )

    def prefetch_host(self, stream, pinned=None, pool=None):
        """Starts the device->host copy on `stream`; `host()` waits for it.  `pinned`: a caller-owned pinned
        buffer of at least n_keep * n_rows doubles (valid for as long as the caller says); `pool`: a list of
        pinned tensors to borrow from - the buffer returns to it when the last host array of this block is gone."""
        import torch

        shape = (int(self.tensor.shape[0]), self.n_rows)
        need = max(1, shape[0] * shape[1])
        lease = None
        if pinned is None:
            buf = None
            if pool is not None:
                for i, cand in enumerate(pool):
                    if cand.numel() >= need:
                        buf = pool.pop(i)
                        break
            if buf is None:
                buf = torch.empty(need, dtype=torch.float64, pin_memory=True)
            lease = HostLease(buf, pool, shape)
            pinned = buf
        with torch.cuda.device(self.tensor.device):
            self._copy_rows(pinned.data_ptr(), stream)
            event = torch.cuda.Event()
            event.record(stream)
        if lease is not None:
            lease.event = event
        self._pending = (lease if lease is not None else pinned[: shape[0] * shape[1]].view(shape), event)

    def host(self) -> np.ndarray:
        """(n_keep, n_rows) float64 numpy array (one device->host copy, cached)."""
        if self._host is None:
            if self._pending is not None:
                dst, event = self._pending
                event.synchronize()
                self._host = np.asarray(dst) if isinstance(dst, HostLease) else dst.numpy()
                self._pending = None
            else:
                import torch

                # lazily, into ordinary memory: page-locking a fresh buffer of this size costs several times the
                # copy itself; the pipelined loaders (load_many) bring pooled pinned buffers instead
                host = np.empty((int(self.tensor.shape[0]), self.n_rows), dtype=np.float64)
                with torch.cuda.device(self.tensor.device):
                    stream = torch.cuda.current_stream(self.tensor.device)
                    self._copy_rows(host.ctypes.data, stream)
                    stream.synchronize()
                self._host = host
        return self._host


class ViconNexusData:
    """The data of one Vicon Nexus CSV file, grouped by device type."""

    def __init__(self, forcepl: Sequence["DeviceData"], emg: "DeviceData", traj: Sequence["DeviceData"]):
        self.forcepl = forcepl
        self.emg = emg
        self.traj = traj

    def check(self) -> None:
        """Raises what loading this trial raised on the device, if it was loaded with
        `defer_check=True` and not checked yet (extension; a no-op otherwise)."""
        pending, self._pending_check = getattr(self, "_pending_check", None), None
        if pending is not None:
            pending()

    def __getitem__(self, device_type: Union[DeviceType, str]):
        device_type = self._parse_device_type(device_type)
        if device_type is DeviceType.FORCE_PLATE:
            return self.forcepl
        if device_type is DeviceType.EMG:
            return self.emg
        if device_type is DeviceType.TRAJECTORY_MARKER:
            return self.traj
        raise KeyError(f"device type not understood: {device_type}")

    def get_cols(self, device_type, device_inds: Optional[Sequence[int]] = None, time=None, cols=None):
        def one(dev: "DeviceData"):
            frame = dev.df if time is None else dev[time]
            return frame[cols]

        device_type = self._parse_device_type(device_type)
        if device_type is DeviceType.EMG:
            return one(self.emg)
        devices = self[device_type]
        if device_inds is not None:
            devices = [devices[i] for i in device_inds]
        return tuple(one(dev) for dev in devices)

    def plot_cols(self, device_type, col, device_inds=None, time=None, labels=None, show=True, **kwargs):
        import matplotlib.pyplot as plt  # plotting is outside the accelerated path

        fig, ax = plt.subplots()
        series = self.get_cols(device_type, device_inds=device_inds, time=time, cols=col)
        if self._parse_device_type(device_type) is DeviceType.EMG:
            series = (series,)
        if labels is None:
            labels = [None] * len(series)
        for label, current in zip(labels, series):
            ax.plot(self.time_seq(device_type), current, label=label, **kwargs)
        if show:
            plt.show()
            return None
        return fig, ax

    def sampling_frequency(self, device_type) -> int:
        return self._get_device_of_type(device_type).sampling_frequency

    def time_seq(self, device_type) -> pd.Series:
        return self._get_device_of_type(self._parse_device_type(device_type)).time_seq()

    def to_framesubfr(self, device_type, index):
        return self._get_device_of_type(device_type).to_framesubfr(index)

    def to_index(self, device_type, frame, subframe: Optional[int] = None):
        return self._get_device_of_type(device_type).to_index(frame, subframe)

    def _get_device_of_type(self, device_type) -> "DeviceData":
        if self._parse_device_type(device_type) is DeviceType.EMG:
            return self.emg
        return self[device_type][0]

    @staticmethod
    def _parse_device_type(device_type):
        try:
            return DeviceType.from_str(device_type)
        except AttributeError:
            return device_type

    def __repr__(self):
        return "ViconNexusData(forcepl=[...], emg=<DeviceData>, traj=[...])"

    def describe(self) -> str:
        def amount(num, noun):
            return f"{num} {noun}{'' if num == 1 else 's'}"

        def members(seq):
            seq = list(seq)
            if len(seq) > 2:
                seq = [seq[0], "...", seq[-1]]
            return ", ".join(map(str, seq))

        return (
            "ViconNexusData:\n"
            f"+ emg: {amount(len(self.emg.df.columns), 'column')}\n"
            f"+ forcepl ({amount(len(self.forcepl), 'device')}): {members(self.forcepl)}\n"
            f"+ traj ({amount(len(self.traj), 'device')}): {members(self.traj)}"
        )


class FrameIndex:
    """(frame, subframe) <-> row index of one CSV section (the contract of user_data.py:483-661).

    One class for both sections, described by two numbers: `per_frame`, the rows a frame occupies in this section
    (the number of subframes for forces / EMG, 1 for trajectories, where every subframe of a frame is the same
    row), and the frame count.  Frames count from 1, subframes from 0, rows from 0.  Scalars, slices and whole
    batches go through the same array arithmetic and the same bounds test."""

    def __init__(self, sampling_freq: SamplingFreq, per_frame_is_subframes: bool):
        self._sampling_freq = sampling_freq
        self._dense = per_frame_is_subframes
        self._times = None

    # ---- geometry
    @property
    def num_frames(self) -> int:
        return self._sampling_freq.num_frames

    @property
    def num_subframes(self) -> int:
        return self._sampling_freq.num_subframes

    @property
    def per_frame(self) -> int:
        return self.num_subframes if self._dense else 1

    @property
    def sampling_frequency(self) -> int:
        return self._sampling_freq.freq_forces_emg if self._dense else self._sampling_freq.freq_traj

    @property
    def final_index(self) -> int:
        return self.num_frames * self.per_frame - 1

    # ---- bounds: the reference's messages (user_data.py:575-597), first offender in argument order
    def _check_rows(self, rows: np.ndarray, given):
        bad = np.flatnonzero((rows < 0) | (rows > self.final_index) | (rows != np.floor(rows)))
        if bad.size:
            raise IndexError(f"index {given[int(bad[0])]} out of bounds (max is self.final_index)")

    def _check_pairs(self, frames: np.ndarray, subs: np.ndarray, given):
        bad_f = (frames < 1) | (frames > self.num_frames) | (frames != np.floor(frames))
        bad_s = (subs < 0) | (subs >= self.num_subframes) | (subs != np.floor(subs))
        bad = np.flatnonzero(bad_f | bad_s)
        if bad.size:
            i = int(bad[0])
            if bad_f[i]:
                raise IndexError(f"frame {given[i][0]} is out of bounds")
            raise IndexError(f"subframe {given[i][1]} out of range")

    # ---- conversions on arrays
    def _rows_of(self, frames: np.ndarray, subs: np.ndarray) -> np.ndarray:
        return (frames - 1) * self.per_frame + (subs if self._dense else 0)

    def _pairs_of(self, rows: np.ndarray):
        return rows // self.per_frame + 1, (rows % self.per_frame if self._dense else np.zeros_like(rows))

    @staticmethod
    def _slice_parts(slice_):
        """The parts of a slice in the order the reference validates them: stop, then start and step if given."""
        return [("stop", slice_.stop)] + [(k, v) for k, v in (("start", slice_.start), ("step", slice_.step)) if v is not None]

    # ---- the reference's two methods
    def to_index(self, frame, subframe=None):
        if subframe is not None:
            return self.to_index_many([(frame, subframe)])[0]
        if not isinstance(frame, slice):
            f, s = frame  # a pair given as one argument is converted unchecked, as in the reference (user_data.py:527-528)
            return int(self._rows_of(np.asarray(f), np.asarray(s)))
        parts = self._slice_parts(frame)
        rows = dict(zip((k for k, _ in parts), self.to_index_many([v for _, v in parts])))
        return slice(rows.get("start"), rows.get("stop"), rows.get("step"))

    def to_framesubfr(self, index):
        if not isinstance(index, slice):
            return self.to_framesubfr_many([index])[0]
        parts = self._slice_parts(index)
        pairs = dict(zip((k for k, _ in parts), self.to_framesubfr_many([v for _, v in parts])))
        return slice(pairs.get("start"), pairs.get("stop"), pairs.get("step"))

    # ---- batches (extension: the windowing stage converts 64 bounds per trial)
    def to_index_many(self, pairs: Sequence[FrameSubfr]) -> List[int]:
        given = [tuple(p) for p in pairs]
        if not given:
            return []
        arr = np.asarray(given)
        self._check_pairs(arr[:, 0], arr[:, 1], given)
        return [int(r) for r in self._rows_of(arr[:, 0].astype(np.int64), arr[:, 1].astype(np.int64))]

    def to_framesubfr_many(self, indices: Sequence[int]) -> List[FrameSubfr]:
        given = list(indices)
        if not given:
            return []
        rows = np.asarray(given)
        self._check_rows(rows, given)
        frames, subs = self._pairs_of(rows.astype(np.int64))
        return [(int(f), int(s)) for f, s in zip(frames, subs)]

    def time_seq(self) -> pd.Series:
        """Sample k (from 0) is at (k + 1) / f (user_data.py:605-608)."""
        if self._times is None or len(self._times) != self.final_index + 1:
            self._times = pd.Series(np.arange(1, self.final_index + 2, 1) * (1 / self.sampling_frequency))
        return self._times


def ForcesEMGFrameTracker(sampling_freq: SamplingFreq) -> FrameIndex:
    """The Devices section: `num_subframes` rows per frame (user_data.py:626-642)."""
    return FrameIndex(sampling_freq, per_frame_is_subframes=True)


def TrajFrameTracker(sampling_freq: SamplingFreq) -> FrameIndex:
    """The Trajectories section: one row per frame, whatever the subframe (user_data.py:645-661)."""
    return FrameIndex(sampling_freq, per_frame_is_subframes=False)


_SectionFrameTracker = FrameIndex


class DeviceData:
    """Data of one measurement device (user_data.py:664-772)."""

    def __init__(
        self,
        device_name: str,
        device_type: DeviceType,
        units,
        frame_tracker: _SectionFrameTracker,
        dataframe: Optional[pd.DataFrame] = None,
        *,
        block: Optional[SectionBlock] = None,
        first_channel: int = 0,
        coords: Optional[Sequence[str]] = None,
    ):
        self.name = device_name
        self.dev_type = device_type
        self.units = tuple(units)
        self._frame_tracker = frame_tracker
        self._df = dataframe
        self._block = block
        self._first_channel = first_channel
        self._coords = list(coords) if coords is not None else (list(dataframe.columns) if dataframe is not None else [])

    # ---- arrays -------------------------------------------------------------------------
    @property
    def tensor(self):
        """(n_columns, n_rows) float64 CUDA tensor: a view of the section block, channel-major."""
        if self._block is None:
            raise AttributeError("this DeviceData was not produced by the CUDA loader")
        c0 = self._first_channel
        return self._block.tensor[c0 : c0 + len(self._coords), : self._block.n_rows]

    @property
    def columns(self):
        """Column names of `df`, without materialising it on the host."""
        return list(self._coords)

    @property
    def values_cm(self) -> np.ndarray:
        """(n_columns, n_rows) host array, channel-major (the block layout of `df`)."""
        c0 = self._first_channel
        return self._block.host()[c0 : c0 + len(self._coords)]

    @property
    def df(self) -> pd.DataFrame:
        if self._df is None:
            # zero-copy: the transposed rows of the channel-major host block ARE the (n_cols, n_rows) C-order
            # block pandas stores; pandas 3 would otherwise copy every device's slice
            self._df = pd.DataFrame(self.values_cm.T, columns=self._coords, dtype=float, copy=False)
        return self._df

    @df.setter
    def df(self, value):
        self._df = value

    # ---- reference API --------------------------------------------------------------------
    @property
    def sampling_frequency(self) -> int:
        return self._frame_tracker.sampling_frequency

    def time_seq(self) -> pd.Series:
        return self._frame_tracker.time_seq()

    def __getitem__(self, indices):
        if isinstance(indices, slice):
            return self.df.iloc[self.to_index(indices)]
        return self.df.iloc[self.to_index(*indices)]

    def to_framesubfr(self, index):
        return self._frame_tracker.to_framesubfr(index)

    def to_framesubfr_many(self, indices):
        return self._frame_tracker.to_framesubfr_many(indices)

    def to_index_many(self, pairs):
        return self._frame_tracker.to_index_many(pairs)

    def to_index(self, frame, subframe: Optional[int] = None):
        return self._frame_tracker.to_index(frame, subframe)

    def __eq__(self, other) -> bool:
        return (
            self.name == other.name
            and self.dev_type == other.dev_type
            and self.units == other.units
            and self.df.equals(other.df)
        )

    def __str__(self):
        return f'DeviceData("{self.name}")'

    def __repr__(self):
        return f"<{str(self)}>"
