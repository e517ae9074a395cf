"""Host-side handling of the 2 x 5 header lines of a Vicon Nexus CSV file and of error text.

The header lines are a few KB; they are parsed on the host with the same rules as the
reference's reader states so that device names, coordinates, units, sampling rates and
every exception message are identical:

    SectionTypeState          reader.py:250-308
    SamplingFrequencyState    reader.py:311-356
    DevicesHeaderFinder       reader.py:381-443
    ForcePlateGrouper         reader.py:446-528
    ForcesEMGDevicesState     reader.py:667-735
    TrajDevicesState          reader.py:738-757
    CoordinatesState          reader.py:760-794
    UnitsState                reader.py:797-835
    DeviceAggregator._my_cols aggregator.py:104-124

The data rows never come through here: they are scanned and parsed by the CUDA kernels.
The only other host work is turning a device-reported error position into the reference's
exception (load_csv.py:128-134) by replaying the reference's per-row rule on that ONE row.
"""
import csv
import io
import locale
from dataclasses import dataclass, field
from typing import List, Optional

from .definitions import DeviceType, SectionType, ViconCSVLines


def file_encoding() -> str:
    """Encoding `open(filename)` uses in text mode (load_csv.py:29)."""
    return locale.getpreferredencoding(False)


def rows_from_bytes(chunk: bytes, max_rows: int):
    """First `max_rows` csv rows of `chunk`, read exactly as load_csv.py:21-31 reads the file
    (text mode, universal newlines, excel dialect).  Returns (rows, physical_lines_consumed).
    """
    text = io.TextIOWrapper(io.BytesIO(chunk), encoding=file_encoding(), newline=None)
    reader = csv.reader(text)
    rows = []
    for row in reader:
        rows.append(row)
        if len(rows) >= max_rows:
            break
    return rows, reader.line_num


def strip_and_trim(row):
    """_ReaderState._preprocess_row (reader.py:116-130)."""
    out = [entry.strip() for entry in row]
    while out and not out[-1]:
        out.pop()
    return out


@dataclass
class DeviceLayout:
    name: str
    device_type: DeviceType
    first_col: int
    last_col: Optional[int]
    coords: List[str] = field(default_factory=list)
    units: List[str] = field(default_factory=list)

    def cut(self, cols):
        # aggregator.py:104-124: an open-ended device (EMG) ends where the first row it is
        # shown (the coordinates line) ends
        if self.last_col is None:
            self.last_col = len(cols) - 1
        return cols[self.first_col : self.last_col + 1]


@dataclass
class SectionLayout:
    kind: SectionType
    frequency: Optional[int] = None
    devices: List[DeviceLayout] = field(default_factory=list)
    num_cols: int = 0
    complete: bool = False  # all five header lines were seen
    _n_keep: Optional[int] = field(default=None, repr=False, compare=False)

    @property
    def n_keep(self) -> int:
        """Channels the device stores: csv columns [2, 2 + n_keep)."""
        if self.complete and self._n_keep is not None:
            return self._n_keep  # a complete layout no longer changes (and is shared through the header cache)
        last = 1
        for dev in self.devices:
            if dev.last_col is not None:
                last = max(last, min(dev.last_col, self.num_cols - 1))
        keep = max(0, last - 1)
        if self.complete:
            self._n_keep = keep
        return keep


class HeaderMachine:
    """Feeds header rows of one section; raises what the reference's states raise."""

    def __init__(self, expected: Optional[SectionType]):
        # expected is None once both sections are over (aggregator.py:292-294)
        self.expected = expected
        self.layout = SectionLayout(kind=expected if expected is not None else SectionType.FORCES_EMG)
        self.next_line = ViconCSVLines.SECTION_TYPE_LINE

    @property
    def done(self) -> bool:
        return self.layout.complete

    def feed(self, row):
        line = self.next_line
        lay = self.layout
        if line is ViconCSVLines.SECTION_TYPE_LINE:
            row = strip_and_trim(row)
            self._single_col(row, line)
            word = row[0]
            if word == "Devices":
                parsed = SectionType.FORCES_EMG
            elif word == "Trajectories":
                parsed = SectionType.TRAJECTORIES
            else:
                raise ValueError(
                    'first row in a section should contain "Devices" or "Trajectories" in its first column'
                )
            if self.expected is None:
                raise AttributeError("'NoneType' object has no attribute 'section_type'")
            if parsed is not self.expected:
                raise ValueError(f"row implies current section is {parsed} but expected {self.expected}")
            self.next_line = ViconCSVLines.SAMPLING_FREQUENCY_LINE
        elif line is ViconCSVLines.SAMPLING_FREQUENCY_LINE:
            row = strip_and_trim(row)
            self._single_col(row, line)
            lay.frequency = int(row[0])
            self.next_line = ViconCSVLines.DEVICE_NAMES_LINE
        elif line is ViconCSVLines.DEVICE_NAMES_LINE:
            row = strip_and_trim(row)
            headers = self._find_headers(row)
            if lay.kind is SectionType.FORCES_EMG:
                plates, emg = headers[:-1], headers[-1]
                for i in range(0, len(plates), 3):
                    col, text = plates[i]
                    plate_name, _ = text.split("-")
                    lay.devices.append(DeviceLayout(plate_name[:-1], DeviceType.FORCE_PLATE, col, col + 9 - 1))
                lay.devices.append(DeviceLayout(emg[1], DeviceType.EMG, emg[0], None))
            else:
                for col, text in headers:
                    lay.devices.append(DeviceLayout(text, DeviceType.TRAJECTORY_MARKER, col, col + 3 - 1))
            self.next_line = ViconCSVLines.COORDINATES_LINE
        elif line is ViconCSVLines.COORDINATES_LINE:
            row = strip_and_trim(row)
            lay.num_cols = len(row)
            for dev in lay.devices:
                dev.coords = dev.cut(row)
            self.next_line = ViconCSVLines.UNITS_LINE
        elif line is ViconCSVLines.UNITS_LINE:
            row = row[0 : lay.num_cols]
            for dev in lay.devices:
                dev.units = dev.cut(row)
            self.next_line = ViconCSVLines.DATA_LINE
            lay.complete = True
        else:  # pragma: no cover - callers stop at DATA_LINE
            raise AssertionError("header machine fed past the units line")

    @staticmethod
    def _single_col(row, line):
        if row[1:]:
            raise ValueError(f"row {line} should contain nothing outside its first column")

    @staticmethod
    def _find_headers(row):
        def error():
            raise ValueError("this line should contain two blank columns then one device name every 3 columns")

        if row[0] or row[1]:
            error()
        for col in range(2, len(row)):
            wants_name = (col - 2) % 3 == 0
            if bool(row[col]) != wants_name:
                error()
        return [(col, row[col]) for col in range(2, len(row), 3)]


def replay_data_row(row, num_cols):
    """DataState.feed_row's parsing of one row (reader.py:927-948); raises what float() raises."""
    return [None if not entry else float(entry) for entry in row[0:num_cols]]


def wrap_error(line_no: int, csv_filename, exc: Exception) -> RuntimeError:
    """load_csv.py:131-134."""
    err = RuntimeError(f"error parsing line {line_no} of file {csv_filename}: " + str(exc))
    err.__cause__ = exc
    return err
