"""ctypes wrapper over tools/synth_vicon.c (synthetic Vicon Nexus CSV generator).

Bench/test infrastructure.  Layouts:
  D    - dynamic_trial.csv shape: 2 plates + 8 EMG @2000 Hz, 40 markers @100 Hz, 62.23 s
  T10  - 10-minute trial: 2 plates + 16 EMG @2000 Hz, 40 markers @100 Hz, 600 s (~473 MB)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libms_synth.so")
_SRC = os.path.join(_HERE, "synth_vicon.c")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", _SRC, "-o", _SO, "-lm"]
        )
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.ms_synth_vicon.restype = ctypes.c_int64
        _lib.ms_synth_vicon.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_int, ctypes.c_double, ctypes.c_int,
        ]
    return _lib


LAYOUTS = {
    "D": dict(seconds=62.23, n_emg=8, n_markers=40),
    "T10": dict(seconds=600.0, n_emg=16, n_markers=40),
    "T127": dict(seconds=127.0, n_emg=16, n_markers=40),  # ~100 MB: configs[2] unit
}


def synth_vicon(seed=0, seconds=1.0, n_plates=2, n_emg=16, n_markers=40, f_emg=2000, f_traj=100,
                crlf=True, blank_marker_frac=0.05, trailing_blank=False, out=None) -> np.ndarray:
    """Returns the CSV bytes as a uint8 numpy array (a view of `out` if given and large enough)."""
    lib = _load()
    n_traj = max(1, int(np.floor(seconds * f_traj + 0.5)))
    n_dev = n_traj * (f_emg // f_traj)
    width = max(2 + 9 * n_plates + n_emg, 2 + 3 * n_markers)
    est = n_dev * (16 + 11 * (9 * n_plates + n_emg) + width) + n_traj * (16 + 10 * 3 * n_markers + width) + (1 << 16)
    buf = out if out is not None and out.nbytes >= est else np.empty(est, dtype=np.uint8)
    args = (ctypes.c_uint64(seed), ctypes.c_double(seconds), n_plates, n_emg, n_markers, f_emg, f_traj,
            int(bool(crlf)), ctypes.c_double(blank_marker_frac), int(bool(trailing_blank)))
    n = lib.ms_synth_vicon(buf.ctypes.data, buf.nbytes, *args)
    if n < 0:
        raise ValueError(f"ms_synth_vicon failed with code {n}")
    if n > buf.nbytes:
        buf = np.empty(n, dtype=np.uint8)
        n = lib.ms_synth_vicon(buf.ctypes.data, buf.nbytes, *args)
    return buf[:n]


def synth_layout(name: str, seed=0, **overrides) -> np.ndarray:
    kw = dict(LAYOUTS[name])
    kw.update(overrides)
    return synth_vicon(seed=seed, **kw)


if __name__ == "__main__":
    import sys
    import time

    name = sys.argv[1] if len(sys.argv) > 1 else "D"
    t = time.time()
    data = synth_layout(name, seed=int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    print(f"{name}: {data.nbytes / 1e6:.1f} MB in {time.time() - t:.2f} s", file=sys.stderr)
    if len(sys.argv) > 2:
        data.tofile(sys.argv[2])
