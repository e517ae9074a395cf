"""The device field parser (csrc/ms_parse_double.cuh), compiled for the host, against
CPython float() - the function the reference calls per field (reader.py:944-948).
Bit-exact or it fails."""
import ctypes
import json
import os
import random
import struct
from fractions import Fraction

import pytest

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def harness():
    import __graft_entry__ as g

    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "native", "libms_parse_harness.so"))
    lib.ms_host_parse.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]
    lib.ms_host_fuzz.restype = ctypes.c_int64
    lib.ms_host_fuzz.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    return lib


def dev_parse(lib, s: bytes):
    b = ctypes.c_uint64(0)
    st = lib.ms_host_parse(s, len(s), ctypes.byref(b))
    return st, b.value


def py_parse(s: bytes):
    try:
        text = s.decode("ascii")
    except UnicodeDecodeError:
        return 2, 0
    try:
        v = float(text)
    except ValueError:
        return 1, 0
    return 0, struct.unpack("<Q", struct.pack("<d", v))[0]


def agree(lib, s: bytes):
    a, b = dev_parse(lib, s), py_parse(s)
    assert a[0] == b[0] and (a[0] != 0 or a[1] == b[1]), (s, a, b)


def test_known_answer_table(harness):
    table = json.load(open(os.path.join(GOLDEN, "float_table.json")))
    for text, want in table.items():
        st, got = dev_parse(harness, text.encode("latin1"))
        if want == "ValueError":
            assert st == 1, text
        else:
            assert st == 0 and got == int(want, 16), (text, hex(got), want)
        agree(harness, text.encode("latin1"))


def py_parse_text(text: str):
    try:
        return 0, struct.unpack("<Q", struct.pack("<d", float(text)))[0]
    except ValueError:
        return 1, 0


def test_non_ascii_fields_follow_float_of_str(harness):
    """float(str) maps Unicode decimal digits and spaces to ASCII first (float("\uff11\uff12")
    is 12.0); bytes that are not UTF-8 cannot come out of open(filename) at all and are flagged."""
    import unicodedata

    rnd = random.Random(21)
    zeros = [cp for cp in range(0x80, 0x110000) if unicodedata.decimal(chr(cp), None) == 0]
    spaces = [chr(cp) for cp in range(0x80, 0x3100) if chr(cp).isspace()]
    cases = ["\uff11\uff12", "\u0663", "\u00a01.5\u2003", "1\u00e9", "x\uff11", "\uff11e\uff12", "\u0967\u0968\u0969.\u096a", "\u2028",
             "1\u00a02", "-\u0660.\u0665", "\U0001d7ce\U0001d7cf", "nan\u3000", "\u221e", "1\x7f"]
    for _ in range(3000):
        parts = []
        for _ in range(rnd.randrange(1, 8)):
            r = rnd.random()
            if r < 0.45:
                parts.append(chr(rnd.choice(zeros) + rnd.randrange(10)))
            elif r < 0.7:
                parts.append(rnd.choice("0123456789.-+e_"))
            elif r < 0.85:
                parts.append(rnd.choice(spaces))
            else:
                parts.append(chr(rnd.choice([0xe9, 0x3b1, 0x4e00, 0x1f600, 0x7f, 0xb2, 0x2160])))  # letters, superscript two, roman one
        cases.append("".join(parts))
    for text in cases:
        raw = text.encode("utf-8")
        got, want = dev_parse(harness, raw), py_parse_text(text)
        assert got[0] == want[0] and (got[0] != 0 or got[1] == want[1]), (text, got, want)
    for raw in (b"1\xff", b"\xc0\xb1", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xe0\x80"):  # stray, overlong, surrogate, > U+10FFFF, cut short
        assert dev_parse(harness, raw)[0] == 2, raw


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_fuzz_against_strtod(harness, mode):
    buf = ctypes.create_string_buffer(512)
    bad = harness.ms_host_fuzz(2024 + mode, 400_000, mode, 0, buf, 512)
    assert bad == 0, buf.value


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_exact_fallback_alone_against_strtod(harness, mode):
    buf = ctypes.create_string_buffer(512)
    bad = harness.ms_host_fuzz(77 + mode, 30_000, mode, 1, buf, 512)
    assert bad == 0, buf.value


def _exact_decimal(fr: Fraction) -> str:
    n, d = fr.numerator, fr.denominator
    k = d.bit_length() - 1
    assert d == 1 << k
    digits = str(n * 5 ** k)
    if k == 0:
        return digits
    digits = digits.rjust(k + 1, "0")
    return digits[:-k] + "." + digits[-k:]


def test_halfway_cases_round_to_even(harness):
    rnd = random.Random(5)
    for i in range(1500):
        e = rnd.choice([0, 1, 2, 52, 500, 1000, 1022, 1023, 1024, 1075, 1100, 1500, 2000, 2045, 2046]) if i % 2 else rnd.randrange(0, 2047)
        b = (e << 52) | rnd.getrandbits(52)
        x = struct.unpack("<d", struct.pack("<Q", b))[0]
        y = struct.unpack("<d", struct.pack("<Q", b + 1))[0]
        if y == float("inf"):
            continue
        s = _exact_decimal((Fraction(x) + Fraction(y)) / 2)
        for v in (s, s + "1", s + "0" * 40 + "1", s[:-1] + str(int(s[-1]) - 1) if s[-1] != "0" else s + "0"):
            agree(harness, v.encode())
        if "." in s:
            ip, fp = s.split(".")
            agree(harness, (ip + fp + "e-%d" % len(fp)).encode())


def test_grammar_fuzz(harness):
    rnd = random.Random(9)
    alphabet = "0123456789" * 3 + ".eE+-_ \tinfatyINFNAT" + "\x0b\x1cx,"
    for _ in range(150_000):
        s = "".join(rnd.choice(alphabet) for _ in range(rnd.randrange(1, 10)))
        agree(harness, s.encode())


def test_division_free_clinger_step_is_exact_exhaustively(tmp_path):
    """a / 10^k via (a*y, fma, fma) with y = RN(1/10^k) equals the IEEE quotient for EVERY integer
    a < 2^26 and k <= 22 (tests/native/div_check.c; the same program was run once for a < 2^32)."""
    import subprocess

    exe = tmp_path / "div_check"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-fopenmp", os.path.join(ROOT, "tests", "native", "div_check.c"),
                           "-o", str(exe), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-400:]
    assert "TOTAL one-correction 0 two-correction 0" in out.stdout
