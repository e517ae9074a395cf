"""Host-side helpers that need no GPU: threaded file reads, lazy result tables, tracker batch conversions."""
import numpy as np
import pytest


def test_read_file_into_matches_plain_read(tmp_path):
    from muscle_synergies_b200.vicon_data import loader

    rng = np.random.default_rng(0)
    old = loader._READ_CHUNK
    loader._READ_CHUNK = 1 << 16  # many chunks on a small file
    try:
        for size in (0, 1, 65535, 65536, 65537, 1_000_003):
            data = rng.integers(0, 256, size, dtype=np.uint8)
            path = tmp_path / f"f{size}"
            data.tofile(path)
            buf = np.full(size + 32, 0xEE, dtype=np.uint8)
            seen = []
            loader.read_file_into(str(path), buf, size, lambda off, length: seen.append((off, length)))
            assert np.array_equal(buf[:size], data) and (buf[size:] == 0xEE).all()
            # chunks are reported in file order and tile the file exactly
            assert [o for o, _ in seen] == sorted(o for o, _ in seen)
            assert sum(length for _, length in seen) == size
            if seen:
                assert seen[0][0] == 0 and seen[-1][0] + seen[-1][1] == size
    finally:
        loader._READ_CHUNK = old
    with pytest.raises(FileNotFoundError):
        loader.read_file_into(str(tmp_path / "missing"), np.empty(8, dtype=np.uint8), 8)


def test_lazy_component_frames():
    from muscle_synergies_b200.pipeline import _Frames

    arrays = {2: np.arange(6, dtype=np.float32).reshape(2, 3), 3: np.ones((3, 3), dtype=np.float32)}
    frames = _Frames(arrays, ["a", "b", "c"])
    assert list(frames) == [2, 3] and len(frames) == 2 and 3 in frames
    df = frames[2]
    assert list(df.columns) == ["a", "b", "c"] and df.shape == (2, 3) and df.to_numpy().dtype == np.float64
    assert frames[2] is df  # built once
    with pytest.raises(KeyError):
        frames[5]


def test_batch_index_conversions_equal_the_scalar_ones():
    from muscle_synergies_b200.vicon_data.data_model import ForcesEMGFrameTracker, TrajFrameTracker
    from muscle_synergies_b200.vicon_data.definitions import SamplingFreq

    freq = SamplingFreq(2000, 100, 50)
    emg, traj = ForcesEMGFrameTracker(freq), TrajFrameTracker(freq)
    idx = [0, 1, 19, 20, 999]
    assert emg.to_framesubfr_many(idx) == [emg.to_framesubfr(i) for i in idx]
    pairs = [(1, 0), (1, 19), (2, 0), (50, 19)]
    assert emg.to_index_many(pairs) == [emg.to_index(f, s) for f, s in pairs]
    assert traj.to_index_many(pairs) == [traj.to_index(f, s) for f, s in pairs]
    for bad in ([1000], [-1], [0.5]):
        with pytest.raises(IndexError):
            emg.to_framesubfr_many(bad)
    for bad in ([(0, 0)], [(51, 0)], [(1, 20)], [(1, -1)]):
        with pytest.raises(IndexError):
            emg.to_index_many(bad)
