"""Import-compatibility alias: the reference keeps load_vicon_file in vicon_data/load_csv.py."""
from .loader import ViconLoader, load_vicon_bytes, load_vicon_file  # noqa: F401
