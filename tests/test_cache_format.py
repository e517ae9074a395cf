"""Trial cache file format (host side only; the GPU round trip is in tests/test_cache_gpu.py)."""
import numpy as np
import pytest

from muscle_synergies_b200 import cache


def _blocks():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(26, 1000))
    a[3, 10] = np.nan
    a[4, 11] = -0.0
    b = rng.normal(size=(120, 50))
    b[:, 7] = np.frombuffer(np.uint64(0x7FF8000000000000).tobytes(), dtype=np.float64)[0]
    return [a, b]


HEAD = {"source": {"path": "x.csv", "size": 1, "mtime_ns": 2}, "sampling": {"forces_emg": 2000, "traj": 100, "num_frames": 50},
        "devices": []}


def test_round_trip_is_bit_exact(tmp_path):
    p = tmp_path / "t.msb200"
    blocks = _blocks()
    cache.write_trial_file(p, HEAD, blocks)
    head, got = cache.read_trial_file(p)
    assert head["sampling"] == HEAD["sampling"] and head["source"] == HEAD["source"] and head["version"] == cache.VERSION
    assert [s["offset"] % cache.ALIGN for s in head["sections"]] == [0, 0]
    for want, have in zip(blocks, got):
        assert have.shape == want.shape
        assert np.array_equal(have.view(np.uint64), want.view(np.uint64))
    assert cache.read_trial_header(p) == head


def test_empty_sections(tmp_path):
    p = tmp_path / "e.msb200"
    cache.write_trial_file(p, HEAD, [np.empty((26, 0)), np.empty((0, 0))])
    _, got = cache.read_trial_file(p)
    assert [g.shape for g in got] == [(26, 0), (0, 0)]


def test_damage_is_detected(tmp_path):
    p = tmp_path / "t.msb200"
    cache.write_trial_file(p, HEAD, _blocks())
    raw = bytearray(p.read_bytes())
    head = cache.read_trial_header(p)
    bad = tmp_path / "bad.msb200"
    flipped = bytearray(raw)
    flipped[head["sections"][1]["offset"] + 5] ^= 0x40
    bad.write_bytes(flipped)
    with pytest.raises(cache.CacheError, match="checksum mismatch in section 1"):
        cache.read_trial_file(bad)
    cache.read_trial_file(bad, verify=False)  # the caller may skip the check
    bad.write_bytes(raw[: head["sections"][1]["offset"] + 100])
    with pytest.raises(cache.CacheError, match="truncated section 1"):
        cache.read_trial_file(bad)
    bad.write_bytes(b"Devices\r\n2000\r\n")
    with pytest.raises(cache.CacheError, match="not a muscle_synergies_b200 trial cache"):
        cache.read_trial_file(bad)
    wrong_version = bytearray(raw)
    wrong_version[8] = 9
    bad.write_bytes(wrong_version)
    with pytest.raises(cache.CacheError, match="format version 9"):
        cache.read_trial_file(bad)


def test_load_trial_needs_a_gpu(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from muscle_synergies_b200._native import NativeError

    p = tmp_path / "t.msb200"
    cache.write_trial_file(p, HEAD, _blocks())
    with pytest.raises(NativeError):
        cache.load_trial(p)
