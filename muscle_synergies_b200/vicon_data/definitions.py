"""Vocabulary of the Vicon Nexus CSV format.

Same names, members and values as the reference's
src/muscle_synergies/vicon_data/definitions.py (SectionType :23, ViconCSVLines :45,
DeviceType :89 with from_str :112-126, ForcePlateMeasurement :135, SamplingFreq :163-199)
so that user code and type checks written against the reference keep working.
"""
from dataclasses import dataclass
from enum import Enum
from typing import List, NewType

Row = NewType("Row", List[str])


class SectionType(Enum):
    FORCES_EMG = 1
    TRAJECTORIES = 2


class ViconCSVLines(Enum):
    SECTION_TYPE_LINE = 1
    SAMPLING_FREQUENCY_LINE = 2
    DEVICE_NAMES_LINE = 3
    COORDINATES_LINE = 4
    UNITS_LINE = 5
    DATA_LINE = 6
    BLANK_LINE = 7


class DeviceType(Enum):
    FORCE_PLATE = 1
    EMG = 2
    TRAJECTORY_MARKER = 3

    @staticmethod
    def from_str(device_type: str) -> "DeviceType":
        key = device_type.upper()
        if key == "EMG":
            return DeviceType.EMG
        if key in ("FORCE PLATE", "FP", "FORCEPL"):
            return DeviceType.FORCE_PLATE
        if key in ("TRAJ", "MARKER"):
            return DeviceType.TRAJECTORY_MARKER
        raise ValueError(f"device type not understood: {device_type}")

    def section_type(self) -> SectionType:
        if self in (DeviceType.EMG, DeviceType.FORCE_PLATE):
            return SectionType.FORCES_EMG
        return SectionType.TRAJECTORIES


class ForcePlateMeasurement(Enum):
    FORCE = 1
    MOMENT = 2
    COP = 3


@dataclass
class SamplingFreq:
    freq_forces_emg: int
    freq_traj: int
    num_frames: int

    @property
    def num_subframes(self) -> int:
        ratio = self.freq_forces_emg / self.freq_traj
        assert ratio == int(ratio)
        return int(ratio)
