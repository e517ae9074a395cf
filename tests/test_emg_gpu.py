"""EMG envelope chain against the numpy/scipy oracle that makes the
reference's own calls.  float64, stated tolerance: relative 1e-9 (sums are ordered differently)."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def emg():
    import __graft_entry__ as g

    g.build()
    from muscle_synergies_b200 import emg

    return emg


def test_dataframe_level_functions(emg):
    from oracle import emg_oracle as eo

    rng = np.random.default_rng(1)
    x = rng.normal(0.002, 0.01, (7001, 5))
    df = pd.DataFrame(x, columns=list("abcde"))
    np.testing.assert_allclose(emg.zero_center(df).to_numpy(), eo.zero_center(x), rtol=RTOL, atol=1e-15)
    for win in (1, 2, 7, 1000, 1001, 7001):
        np.testing.assert_allclose(emg.rms(df, win).to_numpy(), eo.rms(x, win), rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(emg.rms(df, 0.5, sampling_frequency=2000).to_numpy(), eo.rms(x, 1000), rtol=RTOL, atol=1e-13)
    # spans several blocks of the shared-memory kernel; the last window is too long for it (running-sum kernel)
    xl = rng.normal(0.002, 0.01, (60_000, 2))
    dfl = pd.DataFrame(xl, columns=list("ab"))
    for win in (3, 1000, 4097, 20_000, 30_000):
        np.testing.assert_allclose(emg.rms(dfl, win).to_numpy(), eo.rms(xl, win), rtol=RTOL, atol=1e-13)
    # a quiet stretch right after a burst: window sums a million times smaller than the running total
    xb = np.concatenate([rng.normal(0, 1.0, (3000, 2)), rng.normal(0, 1e-3, (5000, 2))])
    dfb = pd.DataFrame(xb, columns=list("ab"))
    for win in (50, 1000):
        got, want = emg.rms(dfb, win).to_numpy(), eo.rms(xb, win)
        np.testing.assert_allclose(got[4200:], want[4200:], rtol=RTOL)  # relative, on the quiet part alone
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(emg.normalize(df).to_numpy(), eo.normalize(x), rtol=RTOL)
    for r in (2, 200, 7001, 9000):
        got = emg.time_normalize(df, r)
        np.testing.assert_allclose(got.to_numpy(), eo.time_normalize(x, r), rtol=RTOL, atol=1e-13)  # 1e-11 of the signal scale
        assert list(got.columns) == list(df.columns) and np.allclose(got.index, np.linspace(0, 1, r))
    inplace = df.copy()
    assert emg.zero_center(inplace, inplace=True) is inplace
    from scipy import interpolate

    cubic = interpolate.interp1d(np.linspace(0, 1, len(df)), df, axis=0, kind="cubic", fill_value="extrapolate")(np.linspace(0, 1, 10))
    assert np.array_equal(emg.time_normalize(df, 10, kind="cubic").to_numpy(), cubic)  # other kinds: scipy, as in the reference


def test_envelope_windows_on_a_segmented_trial(emg):
    """Load -> segment -> per-cycle envelopes on the GPU, against the oracle on the host arrays."""
    import muscle_synergies_b200 as ms
    from muscle_synergies_b200.segment import Cycle, Segmenter, Trecho
    from oracle import emg_oracle as eo
    from tools.synth_vicon import synth_layout

    data = ms.load_vicon_bytes(synth_layout("D", seed=0), name="D")
    seg = Segmenter(data)
    windows = [seg.get_times_of(t, c) for t in Trecho for c in Cycle]
    got = emg.envelope_windows(data.emg, windows, window_size=0.5, reduce_to=200).cpu().numpy()
    ranges = []
    for w in windows:
        sl = data.emg.to_index(w)
        ranges.append((sl.start, sl.stop))
    want = eo.envelope_windows(data.emg.df.to_numpy(), ranges, 1000, 200)
    assert got.shape == want.shape == (8, 200, 8)
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-12)
    assert np.abs(got).max() <= 1.0 + 1e-12


FILTER_RTOL = 1e-9  # of the signal scale: the chunked recursion carries states through a matrix power


def test_filters_against_scipy_calls_of_the_reference(emg):
    """digital_filter / linear_envelope (analysis.py:252-432): chunk-parallel recursion on the GPU vs
    scipy's sequential one, over lengths around the chunk size and the padding."""
    from oracle import emg_oracle as eo

    rng = np.random.default_rng(3)
    for n in (40, 1023, 1024, 1025, 2048, 7001, 50_000):
        x = rng.normal(0.002, 0.01, (n, 3))
        df = pd.DataFrame(x, columns=list("abc"))
        scale = np.abs(x).max()
        cases = [dict(order=4), dict(order=2, zero_lag=False), dict(order=1), dict(order=7),
                 dict(order=3, filter_type="cheby1", cheby_param=1.0), dict(order=4, filter_type="cheby2", cheby_param=30.0)]
        for kw in cases:
            want = eo.linear_envelope(x, 6.0, 2000, **kw)
            got = emg.linear_envelope(df, 6.0, 2000, **kw).to_numpy()
            np.testing.assert_allclose(got, want, rtol=0, atol=FILTER_RTOL * scale, err_msg=f"n={n} {kw}")
        for band, freqs in (("bandpass", (20.0, 450.0)), ("highpass", 20.0), ("bandstop", (55.0, 65.0))):
            for zero_lag in (True, False):
                want = eo.digital_filter(x, freqs, 2000, 4, band_type=band, zero_lag=zero_lag)
                got = emg.digital_filter(df, freqs, 2000, 4, band_type=band, zero_lag=zero_lag).to_numpy()
                np.testing.assert_allclose(got, want, rtol=0, atol=FILTER_RTOL * scale, err_msg=f"n={n} {band} {zero_lag}")
    want = eo.linear_envelope(x, 6.0, 2000, 4, zero_center_=False)
    np.testing.assert_allclose(emg.linear_envelope(df, 6.0, 2000, 4, zero_center_=False).to_numpy(), want, rtol=0,
                               atol=FILTER_RTOL * scale)
    inplace = df.copy()
    assert emg.linear_envelope(inplace, 6.0, 2000, 4, inplace=True) is inplace
    np.testing.assert_allclose(inplace.to_numpy(), eo.linear_envelope(x, 6.0, 2000, 4), rtol=0, atol=FILTER_RTOL * scale)
    with pytest.raises(ValueError, match="greater than padlen, which is 15"):
        emg.linear_envelope(df.iloc[:15], 6.0, 2000, 4)
    with pytest.raises(ValueError, match="filter type not understood"):
        emg.digital_filter(df, 6.0, 2000, 4, filter_type="bessel")


def test_linear_envelope_windows_on_a_segmented_trial(emg):
    import muscle_synergies_b200 as ms
    from muscle_synergies_b200.segment import Cycle, Segmenter, Trecho
    from oracle import emg_oracle as eo
    from tools.synth_vicon import synth_layout

    data = ms.load_vicon_bytes(synth_layout("D", seed=0), name="D")
    seg = Segmenter(data)
    windows = [seg.get_times_of(t, c) for t in Trecho for c in Cycle]
    got = emg.envelope_windows(data.emg, windows, method="linear_envelope", critical_freqs=6.0, order=4).cpu().numpy()
    ranges = [(data.emg.to_index(w).start, data.emg.to_index(w).stop) for w in windows]
    want = eo.envelope_windows_linear(data.emg.df.to_numpy(), ranges, 6.0, 2000, 4)
    assert got.shape == want.shape == (8, 200, 8)
    np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-10)
