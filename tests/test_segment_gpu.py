"""Windowing parity: the CUDA transition search / window gather against the reference's own
Segmenter output (golden, generated in the build container) and the numpy oracle."""
import hashlib
import json
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


@pytest.fixture(scope="module")
def d_trial(ms):
    from tools.synth_vicon import synth_layout

    return ms.load_vicon_bytes(synth_layout("D", seed=0), name="D.csv")


def test_segmenter_matches_reference_golden(ms, d_trial):
    from muscle_synergies_b200.segment import Cycle, Segmenter, Trecho

    gold = json.load(open(os.path.join(GOLDEN, "segment_D.json")))
    seg = Segmenter(d_trial)
    assert seg.transitions == gold["transitions"]
    windows = []
    slices = []
    for w in gold["windows"]:
        trecho, cycle = Trecho[w["trecho"]], Cycle[w["cycle"]]
        assert seg.ith_phase(trecho, w["order"]).name == w["phase"]
        for ref in (w["phase"], w["order"]):
            sl = seg.get_times_of(trecho, cycle, ref)
            assert [int(x) for x in sl.start] == w["start"] and [int(x) for x in sl.stop] == w["stop"]
        slices.append(sl)
        windows.append(w)
    # rows through the reference API (host DataFrame) and through the GPU gather
    emg_cuts = Segmenter.cut(d_trial.emg, slices)
    traj_cuts = Segmenter.cut(d_trial.traj[0], slices)
    for w, sl, ec, tc in zip(windows, slices, emg_cuts, traj_cuts):
        host = d_trial.emg[sl]
        assert list(host.shape) == w["emg_shape"]
        assert hashlib.sha256(bits(host.to_numpy()).tobytes()).hexdigest() == w["emg_sha256"]
        got = np.ascontiguousarray(ec.cpu().numpy().T)
        assert list(got.shape) == w["emg_shape"]
        assert hashlib.sha256(bits(got).tobytes()).hexdigest() == w["emg_sha256"]
        got_t = np.ascontiguousarray(tc.cpu().numpy().T)
        assert list(got_t.shape) == w["traj0_shape"]
        assert hashlib.sha256(bits(got_t).tobytes()).hexdigest() == w["traj0_sha256"]
    cyc = seg.get_times_of(Trecho.SECOND, Cycle.SECOND)
    assert [int(x) for x in cyc.start] == gold["cycle_2_2"]["start"] and [int(x) for x in cyc.stop] == gold["cycle_2_2"]["stop"]
    tre = seg.get_times_of(3)
    assert [int(x) for x in tre.start] == gold["trecho_3"]["start"] and [int(x) for x in tre.stop] == gold["trecho_3"]["stop"]
    assert seg.get_times_of((Trecho.SECOND, Cycle.SECOND)) == cyc
    with pytest.raises(ValueError):
        seg.get_times_of(Trecho.FIRST, None, 1)


def _random_reactions(rnd, n):
    """Piecewise on/off plates with runs of random length, some shorter than 10."""
    left = np.zeros(n)
    right = np.zeros(n)
    i = 0
    while i < n:
        run = int(rnd.choice([1, 2, 3, 5, 9, 10, 11, 15, 40, 200]))
        state = rnd.integers(0, 4)
        left[i : i + run] = rnd.normal(size=min(run, n - i)) * 100 if state & 1 else 0.0
        right[i : i + run] = rnd.normal(size=min(run, n - i)) * 100 if state & 2 else 0.0
        i += run
    # a few NaNs (count as loaded) and negative zeros (count as unloaded)
    for _ in range(5):
        left[rnd.integers(0, n)] = np.nan
        right[rnd.integers(0, n)] = -0.0
    return left, right


@pytest.mark.parametrize("seed", range(8))
def test_transition_search_against_numpy_oracle(ms, seed):
    import torch

    from muscle_synergies_b200.segment import transition_indices
    from oracle import segment_oracle as so

    rnd = np.random.default_rng(seed)
    n = int(rnd.integers(200, 60000))
    left, right = _random_reactions(rnd, n)
    dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    for min_phase, k in ((10, 40), (3, 12), (1, 7), (25, 4)):
        try:
            want = so.transition_indices(left, right, min_phase, k)
        except ValueError:
            with pytest.raises(ValueError):
                transition_indices(dl, dr, min_phase, k)
            continue
        assert transition_indices(dl, dr, min_phase, k) == want
    assert transition_indices(dl, dr, 10, 0) == so.transition_indices(left, right, 10, 0)


def test_run_touching_the_end_counts(ms):
    import torch

    from muscle_synergies_b200.segment import transition_indices
    from oracle import segment_oracle as so

    left = np.array([0.0] * 20 + [5.0] * 4)  # a 4-sample one-leg run cut by the end of the signal
    right = np.zeros(24)
    want = so.transition_indices(left, right, 10, 1)
    assert want == [20]
    assert transition_indices(torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda(), 10, 1) == want


def test_full_size_trial_segments(ms):
    """configs[1]: 10-minute trial - 40 transitions, 32 windows, cut rows equal the source rows."""
    import torch

    from muscle_synergies_b200.segment import Segmenter
    from oracle import segment_oracle as so
    from tools.synth_vicon import synth_layout

    data = ms.load_vicon_bytes(synth_layout("T10", seed=0), name="T10")
    seg = Segmenter(data)
    left = data.forcepl[0].tensor[2].cpu().numpy()
    right = data.forcepl[1].tensor[2].cpu().numpy()
    assert seg.transitions == so.transition_indices(left, right, 10, 40)
    n_sub = 20
    want = so.organize(seg.transitions, left, right, n_sub)
    wins = seg.all_phase_windows()
    assert len(wins) == 32
    for (trecho, cycle, phase, sl), w in zip(wins, want):
        assert phase.name == w["phase"] and tuple(sl.start) == w["start"] and tuple(sl.stop) == w["stop"]
    slices = [w[3] for w in wins]
    for dev, section in ((data.emg, 1), (data.forcepl[1], 1), (data.traj[5], 2)):
        cuts = Segmenter.cut(dev, slices)
        for sl, cut in zip(slices, cuts):
            a, b = so.window_rows(section, tuple(sl.start), tuple(sl.stop), n_sub)
            assert torch.equal(cut.view(torch.int64), dev.tensor[:, a:b].contiguous().view(torch.int64))


def test_fused_submission_equals_step_by_step(ms, d_trial):
    """defer_check + cut_phases_of (one host wait) against the three-step path, EMG and a marker."""
    import torch

    from muscle_synergies_b200.segment import Segmenter
    from tools.synth_vicon import synth_layout

    blob = synth_layout("D", seed=0)
    loader = ms.ViconLoader()
    d = torch.empty(loader.padded_size(blob.nbytes), dtype=torch.uint8, device="cuda")
    d[: blob.nbytes].copy_(torch.from_numpy(blob))
    data = loader.load_device(d, n=blob.nbytes, name="D", defer_check=True)
    seg = Segmenter(data, cut_phases_of=(data.emg, data.traj[3], data.forcepl[1]))
    ref = Segmenter(d_trial)
    assert seg.transitions == ref.transitions
    windows = [w[3] for w in ref.all_phase_windows()]
    assert [w[3] for w in seg.all_phase_windows()] == windows
    for fused_dev, plain_dev in ((data.emg, d_trial.emg), (data.traj[3], d_trial.traj[3]), (data.forcepl[1], d_trial.forcepl[1])):
        got = seg.phase_cuts(fused_dev)
        want = Segmenter.cut(plain_dev, windows)
        assert len(got) == len(want) == 32
        for g, w in zip(got, want):
            assert g.shape == w.shape
            assert torch.equal(g.view(torch.int64), w.view(torch.int64))
    # a device that was not pre-cut is gathered on demand
    later = seg.phase_cuts(data.traj[0])
    assert all(torch.equal(a.view(torch.int64), b.view(torch.int64)) for a, b in zip(later, Segmenter.cut(d_trial.traj[0], windows)))


def test_deferred_check_raises_the_load_error(ms):
    import torch

    from muscle_synergies_b200.segment import Segmenter
    from tools.synth_vicon import synth_layout

    blob = synth_layout("D", seed=0).copy()
    text = blob.tobytes()
    pos = text.index(b"\r\n", 40_000) + 2  # start of a data row well inside the Devices section
    comma = text.index(b",", text.index(b",", pos) + 1) + 1
    bad = bytearray(text)
    bad[comma : comma + 1] = b"x"
    loader = ms.ViconLoader()
    d = torch.empty(loader.padded_size(len(bad)), dtype=torch.uint8, device="cuda")
    d[: len(bad)].copy_(torch.frombuffer(bad, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="error parsing line") as eager:
        loader.load_device(d, n=len(bad), name="bad.csv")
    data = loader.load_device(d, n=len(bad), name="bad.csv", defer_check=True)  # nothing raised yet
    with pytest.raises(RuntimeError) as deferred:
        data.check()
    assert str(deferred.value) == str(eager.value)
    data.check()  # reported once
    data = loader.load_device(d, n=len(bad), name="bad.csv", defer_check=True)
    with pytest.raises(RuntimeError) as via_segmenter:
        Segmenter(data)
    assert str(via_segmenter.value) == str(eager.value)
