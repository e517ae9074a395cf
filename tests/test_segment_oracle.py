"""Pins oracle/segment_oracle.py against the reference Segmenter's own output (golden)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import segment_oracle as so
from oracle import vicon_oracle_fast as vof
from oracle.refstub import reference_available
from tools.synth_vicon import synth_layout


@pytest.fixture(scope="module")
def d_arrays():
    return vof.parse(synth_layout("D", seed=0))


def test_transitions_and_windows_match_reference(d_arrays):
    gold = json.load(open(os.path.join(GOLDEN, "segment_D.json")))
    dev, _traj = d_arrays
    left, right = dev[:, 2], dev[:, 11]  # Fz of plate 1 and plate 2
    t = so.transition_indices(left, right)
    assert t == gold["transitions"]
    got = so.organize(t, left, right, 20)
    assert len(got) == len(gold["windows"]) == 32
    names = {0: "FIRST", 1: "SECOND", 2: "THIRD", 3: "FOURTH"}
    for g, w in zip(got, gold["windows"]):
        assert names[g["trecho"]] == w["trecho"] and names[g["cycle"]] == w["cycle"]
        assert g["phase"] == w["phase"] and g["order"] == w["order"]
        assert list(g["start"]) == w["start"] and list(g["stop"]) == w["stop"]
        a, b = so.window_rows(1, g["start"], g["stop"], 20)
        assert [b - a, 8] == w["emg_shape"]
        a, b = so.window_rows(2, g["start"], g["stop"], 20)
        assert [b - a, 3] == w["traj0_shape"]


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference checkout not present")
def test_oracle_equals_live_reference_on_random_signals():
    import pandas as pd

    from oracle.refstub import import_reference

    _ms, seg = import_reference()
    rnd = np.random.default_rng(3)
    for _ in range(6):
        n = int(rnd.integers(300, 5000))
        left, right = np.zeros(n), np.zeros(n)
        i = 0
        while i < n:
            run = int(rnd.choice([1, 3, 9, 10, 11, 30, 120]))
            state = rnd.integers(0, 4)
            left[i : i + run] = 1.0 if state & 1 else 0.0
            right[i : i + run] = 1.0 if state & 2 else 0.0
            i += run
        found_all = so.transition_indices(left, right, 10, 0)
        k = min(len(found_all), 12)
        if k == 0:
            continue
        ref = [int(x) for x in seg._transition_indices(pd.Series(left), pd.Series(right), 10, k)]
        assert ref == so.transition_indices(left, right, 10, k)
