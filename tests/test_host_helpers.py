"""Host-side helpers that need no GPU: threaded file reads, lazy result tables, tracker batch conversions."""
import numpy as np
import pytest


@pytest.mark.parametrize("mode", ["preadv", "mmap"])
def test_read_file_into_matches_plain_read(tmp_path, mode):
    """Both ways a file reaches the staging buffer: preadv, and a mapping of the file copied with non-temporal stores."""
    import __graft_entry__ as g

    g.build()
    from muscle_synergies_b200.vicon_data import loader

    rng = np.random.default_rng(0)
    old, old_mode = loader._READ_CHUNK, loader._READ_MODE
    loader._READ_CHUNK = 1 << 16  # many chunks on a small file
    loader._READ_MODE = mode
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
        loader._READ_CHUNK, loader._READ_MODE = old, old_mode
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


def test_batch_factor_views_are_made_on_access():
    """analysis._assemble: the factors of a batch stay one packed array; item p is a view of the right shape."""
    from muscle_synergies_b200.analysis import _assemble

    ranks = np.array([1, 3, 2], dtype=np.int32)
    seeds = np.array([7, 8, 9], dtype=np.int64)
    n, m = 5, 4
    Wall = np.arange(n * ranks.sum(), dtype=np.float32)
    Hall = np.arange(m * ranks.sum(), dtype=np.float32) + 100
    res = _assemble(ranks, seeds, n, m, Wall, Hall, np.zeros(3, np.int32), np.zeros(3, np.float32), np.zeros((3, m + 1), np.float32))
    assert len(res.W) == 3 and len(res.H) == 3
    assert res.W[1].shape == (n, 3) and res.H[1].shape == (3, m)
    assert res.W[1][0, 0] == n * 1 and res.H[2][0, 0] == 100 + m * 4
    assert res.W[-1].shape == (n, 2) and [w.shape for w in res.W[0:2]] == [(n, 1), (n, 3)]
    assert np.shares_memory(res.W[2], Wall)
    import pytest

    with pytest.raises(IndexError):
        res.W[3]


def test_stream_entry_points_exist_without_a_gpu():
    """The batch extensions are part of the package's surface (they need a GPU to run, not to import)."""
    from muscle_synergies_b200.segment import PendingSegmenter, Segmenter
    from muscle_synergies_b200.vicon_data import ViconLoader

    assert callable(getattr(ViconLoader, "load_device_many")) and callable(getattr(ViconLoader, "load_many"))
    assert isinstance(getattr(ViconLoader, "work_stream"), property)
    assert callable(Segmenter.begin) and callable(PendingSegmenter.finish)


def test_reader_pool_is_sized_to_the_ranks_share_of_the_box(tmp_path, monkeypatch):
    """One process per GPU (torchrun sets LOCAL_WORLD_SIZE): each rank's reader pool takes its share of the cores, so
    that eight ranks do not run eight full-size pools on one box; few threads -> mapped file + streamed copy."""
    import os

    from muscle_synergies_b200.vicon_data import loader

    cpus = len(os.sched_getaffinity(0))
    data = np.arange(3 << 20, dtype=np.uint8)
    path = tmp_path / "f.bin"
    data.tofile(path)
    old_pool = loader._read_pool
    try:
        for world, want in ((1, max(2, min(loader._READ_THREADS_MAX, cpus - 1))),
                            (8, max(2, min(loader._READ_THREADS_MAX, max(1, cpus // 8) - 1)))):
            monkeypatch.setenv("LOCAL_WORLD_SIZE", str(world))
            loader._read_pool = None
            buf = np.zeros(data.size, dtype=np.uint8)
            loader.read_file_into(str(path), buf, data.size)
            assert loader._read_pool._max_workers == want
            assert np.array_equal(buf, data)
            loader._read_pool.shutdown()
    finally:
        loader._read_pool = old_pool
