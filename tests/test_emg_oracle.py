"""Pins oracle/emg_oracle.py against the live reference preprocessing functions (build container only)."""
import numpy as np
import pandas as pd
import pytest

from oracle import emg_oracle as eo
from oracle.refstub import reference_available


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference checkout not present")
def test_oracle_equals_reference_functions():
    from oracle.refstub import import_reference

    ms, _ = import_reference()
    rng = np.random.default_rng(0)
    x = rng.normal(0.001, 0.01, (5000, 6))
    df = pd.DataFrame(x, columns=list("abcdef"))
    assert np.array_equal(ms.zero_center(df).to_numpy(), eo.zero_center(x))
    assert np.array_equal(ms.rms(df, 0.5, sampling_frequency=2000).to_numpy(), eo.rms(x, 1000))
    assert np.array_equal(ms.normalize(df).to_numpy(), eo.normalize(x))
    assert np.array_equal(ms.time_normalize(df, 200).to_numpy(), eo.time_normalize(x, 200))
    # filters: the oracle makes the reference's scipy calls
    for kw in (dict(order=4), dict(order=2, zero_lag=False), dict(order=3, filter_type="cheby1", cheby_param=1.0),
               dict(order=4, filter_type="cheby2", cheby_param=30.0)):
        want = ms.analysis.linear_envelope(df, 6.0, 2000, **kw).to_numpy()
        assert np.array_equal(want, eo.linear_envelope(x, 6.0, 2000, **kw))
    want = ms.analysis.digital_filter(df, (20.0, 450.0), 2000, 4, band_type="bandpass").to_numpy()
    assert np.array_equal(want, eo.digital_filter(x, (20.0, 450.0), 2000, 4, band_type="bandpass"))
    want = ms.analysis.linear_envelope(df, 6.0, 2000, 4, zero_center_=False).to_numpy()
    assert np.array_equal(want, eo.linear_envelope(x, 6.0, 2000, 4, zero_center_=False))


def test_same_convolution_window_alignment():
    x = np.zeros((50, 1))
    x[20, 0] = 3.0
    out = eo.rms(x, 10)[:, 0]
    # np.convolve "same" with an even window: sample 20 contributes to outputs 16..25
    assert np.flatnonzero(out).tolist() == list(range(16, 26))
