"""ctypes wrapper of oracle/vicon_oracle_c.c (C restatement of the reference loader's data
path).  TEST INFRASTRUCTURE ONLY - see the header of the C file for what it restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvicon_oracle.so")
_lib = None


class Info(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_int64 * 2), ("num_cols", ctypes.c_int32 * 2), ("err_line", ctypes.c_int64)]


def build(force=False):
    src = os.path.join(_HERE, "vicon_oracle_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libvicon_oracle.so"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        for fn in (_lib.mso_count, _lib.mso_parse):
            fn.restype = ctypes.c_int
        _lib.mso_count.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(Info)]
        _lib.mso_parse.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(Info), ctypes.c_void_p, ctypes.c_void_p]
    return _lib


def parse(data: np.ndarray):
    """data: uint8 array with the CSV.  Returns (devices, traj): ROW-major float64 arrays of
    shape (n_rows, num_cols - 2) holding every kept column of the two sections."""
    lib = _load()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    info = Info()
    rc = lib.mso_count(data.ctypes.data, data.nbytes, ctypes.byref(info))
    if rc != 0:
        raise ValueError(f"C oracle cannot handle this file (code {rc}, line {info.err_line})")
    outs = [np.empty((info.n_rows[s], max(0, info.num_cols[s] - 2)), dtype=np.float64) for s in range(2)]
    rc = lib.mso_parse(data.ctypes.data, data.nbytes, ctypes.byref(info), outs[0].ctypes.data, outs[1].ctypes.data)
    if rc != 0:
        raise ValueError(f"C oracle: code {rc} at line {info.err_line}")
    return outs[0], outs[1]
