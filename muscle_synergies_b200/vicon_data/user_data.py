"""Import-compatibility alias: the reference keeps these names in vicon_data/user_data.py."""
from .data_model import (  # noqa: F401
    DeviceData,
    ForcesEMGFrameTracker,
    FrameSubfr,
    SectionBlock,
    TrajFrameTracker,
    ViconNexusData,
)
from .definitions import DeviceType, SamplingFreq  # noqa: F401
