"""float() conformance of the DEVICE build of the field parser (reader.py:940-948).

tests/test_parse_double_host.py exercises csrc/ms_parse_double.cuh compiled for the host; the device build goes
through other intrinsics (__umul64hi, __dmul_rn, __fma_rn, IDP4A) and, in the single-pass kernel, through a
different fast path.  Here the same material - the known-answer table, the four fuzz modes of
tests/native/parse_harness.cpp, halfway cases, a grammar fuzz - is written into the data rows of a two-section
file and parsed ON THE GPU through the public API (ctypes -> C ABI), by both loader paths; every stored double is
compared bit for bit with CPython float(), and every string float() rejects must surface as the reference's error
for exactly that field."""
import json
import os
import random
import struct

import numpy as np
import pytest

from conftest import GOLDEN, bits

pytestmark = pytest.mark.gpu
FIELDS = 30  # test strings per data row


@pytest.fixture(scope="module")
def ms():
    import __graft_entry__ as g

    g.build()
    import muscle_synergies_b200 as ms

    return ms


@pytest.fixture(params=["single_pass", "two_pass"])
def loader_path(request, ms):
    from muscle_synergies_b200.vicon_data import loader as loader_mod

    old = loader_mod.FORCE_PATH
    loader_mod.FORCE_PATH = "fused" if request.param == "single_pass" else "two_pass"
    yield request.param
    loader_mod.FORCE_PATH = old


def py_float_bits(text: str):
    """bits of float(text), or None when float() raises."""
    try:
        return struct.unpack("<Q", struct.pack("<d", float(text)))[0]
    except ValueError:
        return None


def usable(text: str) -> bool:
    """Can stand as an unquoted csv field that the device sees byte for byte."""
    return text != "" and all(c not in text for c in ',\r\n"') and all(ord(c) < 0x80 for c in text)


def sm64(state):
    state[0] = (state[0] + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state[0]
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def fuzz_strings(mode: int, n: int, seed: int):
    """The generators of ms_host_fuzz (tests/native/parse_harness.cpp), restated."""
    st = [seed]
    out = []
    for _ in range(n):
        if mode in (0, 1):
            b = sm64(st)
            if (b >> 52) & 0x7FF == 0x7FF:
                b &= ~(1 << 62)
            d = struct.unpack("<d", struct.pack("<Q", b))[0]
            out.append("%.17g" % d if mode == 0 else "%.*e" % (sm64(st) % 25, d))
        elif mode == 2:
            nd = 1 + sm64(st) % 40
            dot = sm64(st) % (nd + 1)
            s = "-" if sm64(st) & 1 else ""
            for k in range(nd):
                if k == dot:
                    s += "."
                s += chr(ord("0") + sm64(st) % 10)
            if dot == nd and sm64(st) & 1:
                s += "."
            if sm64(st) % 4:
                e = sm64(st) % 656 - 345
                s += ("e" if sm64(st) & 1 else "E") + str(e)
            out.append(s)
        else:
            r = sm64(st)
            u = (sm64(st) >> 11) / 9007199254740992.0 - 0.5
            v = u * (0.05, 3000.0, 1e5, 2e-4)[r % 4]
            out.append("%.2E" % v if v != 0 and abs(v) < 1e-4 else "%.6g" % v)
    return out


def halfway_strings(n: int, seed: int):
    from fractions import Fraction

    rnd = random.Random(seed)
    out = []
    for i in range(n):
        e = rnd.choice([0, 1, 2, 52, 500, 1000, 1022, 1023, 1024, 1075, 1100, 1500, 2000, 2045, 2046]) if i % 2 else rnd.randrange(0, 2047)
        b = (e << 52) | rnd.getrandbits(52)
        x = struct.unpack("<d", struct.pack("<Q", b))[0]
        y = struct.unpack("<d", struct.pack("<Q", b + 1))[0]
        if y == float("inf"):
            continue
        mid = (Fraction(x) + Fraction(y)) / 2
        k = mid.denominator.bit_length() - 1
        digits = str(mid.numerator * 5 ** k)
        s = digits if k == 0 else digits.rjust(k + 1, "0")[:-k] + "." + digits.rjust(k + 1, "0")[-k:]
        out += [s, s + "1", s + "0" * 40 + "1"]
    return out


def vicon_file(rows):
    """A Devices section whose EMG device has FIELDS channels holding `rows` (lists of FIELDS strings), and a minimal
    Trajectories section.  LF line ends."""
    names = ",".join(f"m{i}" for i in range(FIELDS))
    lines = ["Devices", "1000", ",,P #1 - Force,,,P #1 - Moment,,,P #1 - CoP,,,EMG - V",
             "Frame,Sub Frame,Fx,Fy,Fz,Mx,My,Mz,Cx,Cy,Cz," + names, ",,N,N,N,N.mm,N.mm,N.mm,mm,mm,mm," + ",".join(["V"] * FIELDS)]
    for r, fields in enumerate(rows):
        lines.append(f"{r // 10 + 1},{r % 10},0,0,0,0,0,0,0,0,0," + ",".join(fields))
    lines.append("," * (10 + FIELDS))
    lines += ["Trajectories", "100", ",,S:M0", "Frame,Sub Frame,X,Y,Z", ",,mm,mm,mm", "1,0,1,2,3", ""]
    return "\n".join(lines).encode("ascii")


def test_every_accepted_string_gives_float_s_bits(ms, loader_path):
    table = json.load(open(os.path.join(GOLDEN, "float_table.json")))
    strings = [t for t, want in table.items() if want != "ValueError"]
    for mode in range(4):
        strings += fuzz_strings(mode, 60_000, 2024 + mode)
    strings += halfway_strings(1500, 5)
    rnd = random.Random(11)
    alphabet = "0123456789" * 3 + ".eE+-_ \tinfatyINFNAT"
    strings += ["".join(rnd.choice(alphabet) for _ in range(rnd.randrange(1, 10))) for _ in range(150_000)]
    strings += [" %s " % s for s in strings[:2000]] + ["\t%s" % s for s in strings[2000:3000]]
    good = [(s, py_float_bits(s)) for s in strings if usable(s)]
    good = [(s, b) for s, b in good if b is not None]
    assert len(good) > 200_000
    while len(good) % FIELDS:
        good.append(("1", py_float_bits("1")))
    rows = [[s for s, _ in good[i : i + FIELDS]] for i in range(0, len(good), FIELDS)]
    want = np.array([b for _, b in good], dtype=np.uint64).reshape(len(rows), FIELDS)
    data = ms.load_vicon_bytes(np.frombuffer(vicon_file(rows), dtype=np.uint8), name="floats.csv")
    got = bits(data.emg.tensor.cpu().numpy().T)
    assert got.shape == want.shape
    bad = np.argwhere(got != want)
    assert bad.size == 0, [(rows[r][c], hex(int(got[r, c])), hex(int(want[r, c]))) for r, c in bad[:10]]


def test_every_rejected_string_raises_for_that_field(ms, loader_path):
    """One bad field per row; the first one in file order is what the reference reports (load_csv.py:128-134).
    Fix it, load again, and so on down the file."""
    table = json.load(open(os.path.join(GOLDEN, "float_table.json")))
    rnd = random.Random(13)
    alphabet = "0123456789" * 2 + ".eE+-_ \tinfatyINFNAT\x0b\x1cx"
    cands = [t for t, want in table.items() if want == "ValueError"]
    cands += ["".join(rnd.choice(alphabet) for _ in range(rnd.randrange(1, 9))) for _ in range(6000)]
    cands += ["1e", "1e+", "--1", "1..2", "0x10", "1_", "_1", "1__0", ".", "-", "+", "e5", "1 2", "0.5.", "nan1", "in", "- 1"]
    bad = [s for s in cands if usable(s) and py_float_bits(s) is None][:400]
    assert len(bad) >= 300
    rows, where = [], []
    for s in bad:
        fields = ["%.6g" % rnd.uniform(-50, 50) for _ in range(FIELDS)]
        col = rnd.randrange(FIELDS)
        fields[col] = s
        rows.append(fields)
        where.append(col)
    for i, s in enumerate(bad):
        blob = np.frombuffer(vicon_file(rows), dtype=np.uint8)
        with pytest.raises(RuntimeError) as err:
            ms.load_vicon_bytes(blob, name="bad.csv")
        assert str(err.value) == f"error parsing line {6 + i} of file bad.csv: could not convert string to float: {s!r}", (i, s)
        assert isinstance(err.value.__cause__, ValueError)
        rows[i][where[i]] = "1"
    ms.load_vicon_bytes(np.frombuffer(vicon_file(rows), dtype=np.uint8), name="fixed.csv")


def test_integration_stub_runs_verbatim(ms):
    """The binding INTEGRATION.md shows a reference maintainer (ctypes, no torch types in the signatures) is executed
    as written and reproduces the reference's arrays for sample_data/abridged_data.csv."""
    import re

    from conftest import ROOT, load_npz_u64

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "def parse_data_rows" in b)
    scope = {"__name__": "integration_stub", "MS_B200_LIB": os.path.join(ROOT, "muscle_synergies_b200", "libms_b200.so")}
    exec(compile(stub, "INTEGRATION.md", "exec"), scope)  # noqa: S102 - our own documentation
    raw = open(os.path.join(GOLDEN, "abridged_data.csv"), "rb").read()
    dev, traj = scope["parse_data_rows"](raw)
    want = load_npz_u64(os.path.join(GOLDEN, "abridged_expected.npz"))
    info = json.load(open(os.path.join(GOLDEN, "abridged_expected.json")))
    col = {0: 0, 1: 0}
    for i, m in enumerate(info["meta"]["devices"]):
        sec = 1 if m["dev_type"] == "TRAJECTORY_MARKER" else 0
        block = traj if sec else dev
        width = m["shape"][1]
        got = bits(np.ascontiguousarray(block[col[sec] : col[sec] + width].T))
        assert (got == want[f"dev{i}"]).all(), m["name"]
        col[sec] += width


def test_unicode_digits_and_spaces_as_float_of_str_takes_them(ms, loader_path):
    """float(str) rewrites Unicode decimal digits and spaces to ASCII before parsing (SURVEY.md Appendix B:
    float("１２") == 12.0); so does the device parser, on the UTF-8 bytes of the field."""
    import codecs
    import unicodedata

    from muscle_synergies_b200.vicon_data.header import file_encoding

    if codecs.lookup(file_encoding()).name != "utf-8":
        pytest.skip("open(filename) does not decode UTF-8 under this locale: the reference cannot read such a file either")
    rnd = random.Random(31)
    zeros = [cp for cp in range(0x80, 0x20000) if unicodedata.decimal(chr(cp), None) == 0]
    fields = ["１２", "٣", " 1.5 ", "１e２", "१२३.४", "-٠.٥", "1_෩.᱆"]
    while len(fields) < 20 * FIELDS:
        digits = "".join(chr(rnd.choice(zeros) + rnd.randrange(10)) if rnd.random() < 0.6 else rnd.choice("0123456789")
                         for _ in range(rnd.randrange(1, 9)))
        k = rnd.randrange(len(digits) + 1)
        text = rnd.choice(["", "-", " "]) + digits[:k] + ("." if rnd.random() < 0.6 else "") + digits[k:]
        if py_float_bits(text) is not None:
            fields.append(text)
    fields = fields[: len(fields) // FIELDS * FIELDS]
    rows = [fields[i : i + FIELDS] for i in range(0, len(fields), FIELDS)]
    names = ",".join(f"m{i}" for i in range(FIELDS))
    lines = ["Devices", "1000", ",,P #1 - Force,,,P #1 - Moment,,,P #1 - CoP,,,EMG - V",
             "Frame,Sub Frame,Fx,Fy,Fz,Mx,My,Mz,Cx,Cy,Cz," + names, ",,N,N,N,N.mm,N.mm,N.mm,mm,mm,mm," + ",".join(["V"] * FIELDS)]
    for r, f in enumerate(rows):
        lines.append(f"{r // 10 + 1},{r % 10},0,0,0,0,0,0,0,0,0," + ",".join(f))
    lines.append("," * (10 + FIELDS))
    lines += ["Trajectories", "100", ",,S:M0", "Frame,Sub Frame,X,Y,Z", ",,mm,mm,mm", "1,0,1,2,3", ""]
    blob = np.frombuffer("\n".join(lines).encode("utf-8"), dtype=np.uint8)
    data = ms.load_vicon_bytes(blob, name="unicode.csv")
    got = bits(data.emg.tensor.cpu().numpy().T)
    want = np.array([py_float_bits(t) for t in fields], dtype=np.uint64).reshape(len(rows), FIELDS)
    bad = np.argwhere(got != want)
    assert bad.size == 0, [(rows[r][c], hex(int(got[r, c])), hex(int(want[r, c]))) for r, c in bad[:10]]
