"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/ms_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "ms_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(ms_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_path():
    names = _declared_functions()
    for needed in ("ms_scan", "ms_parse", "ms_find_transitions", "ms_cut_windows", "ms_workspace_bytes"):
        assert needed in names


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from muscle_synergies_b200 import _native

    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ms_b200.h but not exported"
    assert b"sm_100a" in _native.lib().ms_version()
    assert _native.lib().ms_workspace_bytes(1 << 20) > 0


def test_library_contains_sm_100a_code():
    from muscle_synergies_b200 import _native

    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:  # cuobjdump missing: not a failure of the build
        return
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import muscle_synergies_b200 as ms
    from muscle_synergies_b200._native import NativeError

    with pytest.raises(NativeError):
        ms.load_vicon_file(os.path.join(ROOT, "tests", "golden", "abridged_data.csv"))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the product may import, link or call it."""
    pkg = os.path.join(ROOT, "muscle_synergies_b200")
    for base, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "vicon_oracle" not in text and "libvicon_oracle" not in text, f


def test_nmf_plan_is_host_only_and_lays_the_batch_out():
    """ms_nmf_plan (pure host code): the problem table a repeated sweep keeps on the device - rank, then the offsets
    of each problem's W, H and X in the packed arrays (32 bytes per problem)."""
    import numpy as np

    from muscle_synergies_b200 import _native

    lib = ctypes.CDLL(_native.LIB_PATH)
    lib.ms_nmf_plan.restype = ctypes.c_int32
    i32p = ctypes.POINTER(ctypes.c_int32)
    ranks = np.array([1, 3, 8, 2], dtype=np.int32)
    xi = np.array([0, 0, 1, 2], dtype=np.int32)
    n, m = 200, 16
    table = np.zeros(len(ranks) * 4, dtype=np.int64)
    kmax = lib.ms_nmf_plan(n, m, ranks.ctypes.data_as(i32p), xi.ctypes.data_as(i32p), len(ranks), ctypes.c_void_p(table.ctypes.data))
    assert kmax == 8
    rows = table.reshape(len(ranks), 4)
    assert (rows[:, 0] & 0xFFFFFFFF).tolist() == ranks.tolist()
    assert rows[:, 1].tolist() == [0, 200, 800, 2400]          # W offsets: n * sum of the ranks before
    assert rows[:, 2].tolist() == [0, 16, 64, 192]             # H offsets: m * sum of the ranks before
    assert rows[:, 3].tolist() == [0, 0, n * m, 2 * n * m]     # X offsets: matrix index * n * m
    bad = np.array([0], dtype=np.int32)
    assert lib.ms_nmf_plan(n, m, bad.ctypes.data_as(i32p), None, 1, ctypes.c_void_p(table.ctypes.data)) < 0


def test_host_copy_stream_copies_exactly():
    """ms_host_copy_stream (pure host code, non-temporal stores): every size and alignment, nothing outside the range."""
    import numpy as np

    from muscle_synergies_b200 import _native

    lib = ctypes.CDLL(_native.LIB_PATH)
    lib.ms_host_copy_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    rng = np.random.default_rng(0)
    for n in (0, 1, 15, 63, 64, 65, 127, 1000, 4096, (1 << 20) + 7):
        for src_off in (0, 3):
            for dst_off in (0, 5, 17, 48):
                src = rng.integers(0, 255, n + src_off + 64, dtype=np.uint8)
                dst = np.zeros(n + dst_off + 64, dtype=np.uint8)
                assert lib.ms_host_copy_stream(dst.ctypes.data + dst_off, src.ctypes.data + src_off, n) == 0
                assert (dst[dst_off : dst_off + n] == src[src_off : src_off + n]).all()
                assert dst[:dst_off].sum() == 0 and dst[dst_off + n :].sum() == 0
    assert lib.ms_host_copy_stream(None, None, 0) == 0
    assert lib.ms_host_copy_stream(None, None, 8) < 0
