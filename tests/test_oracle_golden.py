"""Pins the CPU oracles (oracle/vicon_oracle.py, oracle/vicon_oracle_c.c) against
  * the literals of the reference's own tests (tests/func/conftest.py:96-311,
    tests/func/test_data_loading.py:9-61), restated by hand below, and
  * golden vectors produced by running the unmodified reference (oracle/make_golden.py).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits, load_npz_u64
from oracle import vicon_oracle as vo
from oracle import vicon_oracle_fast as vof
from tools.synth_vicon import synth_layout, synth_vicon

ABRIDGED = os.path.join(GOLDEN, "abridged_data.csv")
NAN = float("nan")

# tests/func/conftest.py:101-162 (EMG), :174-181 / :193-200 (force plates), :212-260 (markers)
EXP_EMG = [
    [0.0037236, 0.00722359, 0.00344124, 0.00149971, -0.000798493, -0.00196037, -0.00602333, -0.00232391],
    [0.00463913, 0.00478218, 0.00206795, 0.000889358, -3.56e-05, -0.00150261, -0.00373451, -0.0036972],
    [0.00448654, 0.00142525, 0.000389481, -2.62e-05, -0.000798493, -0.00241814, 0.00191124, -0.00537567],
    [0.00235031, -0.00147392, -0.00098381, -0.0021624, -0.000493317, -0.000587082, 0.00786217, -0.00644379],
    [0.00204514, -0.00223686, -0.000220871, -0.0021624, -0.00156143, 0.00200691, 0.0128976, -0.00522308],
    [0.000519257, 5.20e-05, 0.00115242, -0.000789109, -0.00140884, 0.00246468, 0.014576, -0.0012558],
]
EXP_FP1 = [[0, 0, 0, 0, 0, 0, 232, 254, 0]] * 6
EXP_FP2 = [[0, 0, 0, 0, 0, 0, 232, 769, 0]] * 6
EXP_HV = [[209.331, 1219.74, 1780.67], [209.475, 1219.82, 1780.88]]
EXP_CLE = [[227.725, 1091.81, 496.721], [227.702, 1091.8, 496.729]]
EXP_NAN = [[NAN, NAN, NAN], [NAN, NAN, NAN]]
FP_COLS = ["Fx", "Fy", "Fz", "Mx", "My", "Mz", "Cx", "Cy", "Cz"]
FP_UNITS = ["N", "N", "N", "N.mm", "N.mm", "N.mm", "mm", "mm", "mm"]
EMG_COLS = ["VL", "RF", "GMED", "TFL", "GMAXS", "GMAXI", "BF", "ST"]

EXPECTED = [
    ("Imported AMTI OR6 Series Force Plate #1", vo.FORCE_PLATE, FP_COLS, FP_UNITS, EXP_FP1),
    ("Imported AMTI OR6 Series Force Plate #2", vo.FORCE_PLATE, FP_COLS, FP_UNITS, EXP_FP2),
    ("EMG2000 - Voltage", vo.EMG, EMG_COLS, ["V"] * 8, EXP_EMG),
    ("Angelica:HV", vo.TRAJECTORY_MARKER, ["X", "Y", "Z"], ["mm"] * 3, EXP_HV),
    ("Angelica:CM_E", vo.TRAJECTORY_MARKER, ["X", "Y", "Z"], ["mm"] * 3, EXP_NAN),
    ("Angelica:CL_E", vo.TRAJECTORY_MARKER, ["X", "Y", "Z"], ["mm"] * 3, EXP_CLE),
    ("Angelica:ELAST_DP", vo.TRAJECTORY_MARKER, ["X", "Y", "Z"], ["mm"] * 3, EXP_NAN),
]


def check_against_meta(res, meta, arrays):
    devs = res.all_devices()
    assert len(devs) == len(meta["devices"])
    for i, (dev, m) in enumerate(zip(devs, meta["devices"])):
        assert dev.name == m["name"]
        assert dev.dev_type == m["dev_type"]
        assert list(dev.units) == m["units"]
        assert list(dev.coords) == m["coords"]
        got = bits(vo.device_array(dev))
        assert list(got.shape) == m["shape"]
        assert (got == arrays[f"dev{i}"]).all(), dev.name
    assert res.num_frames == meta["num_frames"]


def test_abridged_literals_of_the_reference_tests():
    res = vo.load_vicon_file_oracle(ABRIDGED)
    assert res.freq == {1: 300, 2: 100}  # test_data_loading.py:28-32
    for dev, (name, kind, cols, units, rows) in zip(res.all_devices(), EXPECTED):
        assert (dev.name, dev.dev_type, dev.coords, list(dev.units)) == (name, kind, cols, units)
        want = bits(np.array(rows, dtype=np.float64))
        assert (bits(vo.device_array(dev)) == want).all(), name
    # conftest.py:267-311: (frame, subframe) -> index, and the invalid pairs
    pairs = [(1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2)]
    assert [vo.to_index(1, f, s, 2, 3) for f, s in pairs] == [0, 1, 2, 3, 4, 5]
    assert [vo.to_index(2, f, s, 2, 3) for f, s in pairs] == [0, 0, 0, 1, 1, 1]
    for f, s in [(-1, 0), (0, 3), (1, 3), (3, 0), (3, 2)]:
        with pytest.raises(IndexError):
            vo.to_index(1, f, s, 2, 3)
    # test_data_loading.py:47-51
    hv = vo.device_array(res.traj[0])
    assert list(hv[vo.to_index(2, 2, 2, 2, 3)]) == [209.475, 1219.82, 1780.88]


def test_abridged_against_reference_arrays():
    info = json.load(open(os.path.join(GOLDEN, "abridged_expected.json")))
    arrays = load_npz_u64(os.path.join(GOLDEN, "abridged_expected.npz"))
    check_against_meta(vo.load_vicon_file_oracle(ABRIDGED), info["meta"], arrays)


def test_variants_against_reference(variants_table):
    for name, info in variants_table.items():
        path = os.path.join(GOLDEN, "variants", name + ".csv")
        if info["raises"] is None:
            res = vo.load_vicon_file_oracle(path)
            check_against_meta(res, info["meta"], load_npz_u64(os.path.join(GOLDEN, "variants", name + ".npz")))
        else:
            with pytest.raises(Exception) as err:
                vo.load_vicon_file_oracle(path)
            assert type(err.value).__name__ == info["raises"], name
            if info["raises"] == "RuntimeError":
                assert str(err.value).replace(path, name + ".csv") == info["message"], name
                assert type(err.value.__cause__).__name__ == info["cause"], name


@pytest.mark.parametrize("tag", ["lf", "crlf", "narrow"])
def test_small_synthetic_against_reference(tag, tmp_path):
    info = json.load(open(os.path.join(GOLDEN, f"synth_small_{tag}.json")))
    blob = synth_vicon(**info["generator"])
    assert hashlib.sha256(blob.tobytes()).hexdigest() == info["csv_sha256"], "generator output changed"
    path = tmp_path / "t.csv"
    blob.tofile(path)
    arrays = load_npz_u64(os.path.join(GOLDEN, f"synth_small_{tag}.npz"))
    check_against_meta(vo.load_vicon_file_oracle(str(path)), info["meta"], arrays)
    # the C restatement agrees too
    dev, traj = vof.parse(blob)
    n1 = sum(1 for d in info["meta"]["devices"] if d["dev_type"] != "TRAJECTORY_MARKER")
    want1 = np.concatenate([arrays[f"dev{i}"] for i in range(n1)], axis=1)
    want2 = np.concatenate([arrays[f"dev{i}"] for i in range(n1, len(info["meta"]["devices"]))], axis=1)
    assert (bits(dev) == want1).all() and (bits(traj) == want2).all()


def test_c_oracle_on_d_layout_against_reference_digests():
    info = json.load(open(os.path.join(GOLDEN, "synth_D_digest.json")))
    blob = synth_layout("D", seed=0)
    assert hashlib.sha256(blob.tobytes()).hexdigest() == info["csv_sha256"], "generator output changed"
    dev, traj = vof.parse(blob)
    col = 0
    for i, m in enumerate(info["meta"]["devices"]):
        src = dev if m["dev_type"] != "TRAJECTORY_MARKER" else traj
        if m["dev_type"] == "TRAJECTORY_MARKER" and i == 3:
            col = 0
        w = m["shape"][1]
        got = np.ascontiguousarray(src[:, col : col + w]).view(np.uint64)
        assert hashlib.sha256(got.tobytes()).hexdigest() == info["sha256"][f"dev{i}"], m["name"]
        col += w


def test_c_oracle_on_variants(variants_table):
    for name, info in variants_table.items():
        if info["raises"] is not None or name in ("quoted_number",):
            continue
        blob = np.fromfile(os.path.join(GOLDEN, "variants", name + ".csv"), dtype=np.uint8)
        dev, traj = vof.parse(blob)
        arrays = load_npz_u64(os.path.join(GOLDEN, "variants", name + ".npz"))
        devices = info["meta"]["devices"]
        n1 = sum(1 for d in devices if d["dev_type"] != "TRAJECTORY_MARKER")
        want1 = np.concatenate([arrays[f"dev{i}"] for i in range(n1)], axis=1)
        want2 = np.concatenate([arrays[f"dev{i}"] for i in range(n1, len(devices))], axis=1)
        assert (bits(dev) == want1).all() and (bits(traj) == want2).all(), name
