"""Vicon Nexus CSV loading on B200: same public names as the reference's vicon_data package
(src/muscle_synergies/vicon_data/__init__.py:17-26)."""
from .data_model import DeviceData, ViconNexusData
from .definitions import DeviceType, SamplingFreq, SectionType, ViconCSVLines
from .loader import ViconLoader, load_vicon_bytes, load_vicon_file

__all__ = (
    "load_vicon_file",
    "ViconNexusData",
    "DeviceData",
)
