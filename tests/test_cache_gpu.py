"""Binary trial cache round trip through the GPU loader."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ms():
    import __graft_entry__ as g

    g.build()
    import muscle_synergies_b200 as ms

    return ms


def _same(a, b):
    assert (a.name, a.dev_type, a.units, a.columns, a.sampling_frequency) == (b.name, b.dev_type, b.units, b.columns, b.sampling_frequency)
    assert np.array_equal(bits(a.df.to_numpy()), bits(b.df.to_numpy()))
    assert np.array_equal(bits(a.tensor.cpu().numpy()), bits(b.tensor.cpu().numpy()))


def test_abridged_round_trip(ms, tmp_path):
    from muscle_synergies_b200 import cache

    data = ms.load_vicon_file(os.path.join(GOLDEN, "abridged_data.csv"))
    p = tmp_path / "abridged.msb200"
    cache.save_trial(data, p)
    back = cache.load_trial(p, verify=True)
    for a, b in zip(list(data.forcepl) + [data.emg] + list(data.traj), list(back.forcepl) + [back.emg] + list(back.traj)):
        _same(a, b)
    assert len(back.forcepl) == 2 and len(back.traj) == len(data.traj)
    assert back.emg.to_index(2, 1) == data.emg.to_index(2, 1) and back.traj[0].to_index(2, 0) == 1
    with pytest.raises(IndexError):
        back.emg.to_index(3, 0)
    assert back.emg[(1, 0):(2, 1)].equals(data.emg[(1, 0):(2, 1)])


def test_cached_loader_and_segmentation(ms, tmp_path):
    from muscle_synergies_b200 import cache
    from muscle_synergies_b200.segment import Segmenter
    from tools.synth_vicon import synth_layout

    csv = tmp_path / "trial.csv"
    csv.write_bytes(synth_layout("D", seed=3).tobytes())
    first = cache.load_vicon_file_cached(csv)
    cpath = cache.cache_path_for(csv)
    assert os.path.exists(cpath)
    stamp = os.stat(cpath).st_mtime_ns
    second = cache.load_vicon_file_cached(csv)  # served from the cache
    assert os.stat(cpath).st_mtime_ns == stamp
    _same(first.emg, second.emg)
    _same(first.traj[5], second.traj[5])
    assert Segmenter(second).transitions == Segmenter(first).transitions
    # touching the CSV invalidates the cache
    os.utime(csv, ns=(os.stat(csv).st_atime_ns, os.stat(csv).st_mtime_ns + 10**9))
    third = cache.load_vicon_file_cached(csv)
    assert os.stat(cpath).st_mtime_ns != stamp
    _same(first.forcepl[1], third.forcepl[1])
    # a damaged cache is ignored and rewritten
    with open(cpath, "r+b") as f:
        f.write(b"garbage!")
    fourth = cache.load_vicon_file_cached(csv)
    _same(first.emg, fourth.emg)
    assert cache.read_trial_header(cpath)["source"]["size"] == os.stat(csv).st_size
