"""Calls a user of the reference makes that do not touch the GPU: what the accelerated stages do not implement is
forwarded to scikit-learn / scipy exactly as the reference forwards it (analysis.py:551-594, 848-914)."""
import warnings

import numpy as np
import pandas as pd
import pytest


@pytest.fixture(scope="module")
def envelopes():
    rng = np.random.default_rng(0)
    t = np.linspace(0, 1, 200)[:, None]
    basis = np.abs(np.sin(np.pi * (rng.uniform(0.5, 3, (1, 3)) * t + rng.uniform(0, 1, (1, 3))))) ** 2
    X = basis @ rng.uniform(0, 1, (3, 8)) + 0.02 * rng.uniform(0, 1, (200, 8))
    return pd.DataFrame(X / X.max(axis=0), columns=[f"m{i}" for i in range(8)])


def test_find_synergies_as_the_tutorial_calls_it(envelopes):
    """docs/source/tutorials/Finding muscle synergies.ipynb cell 26: find_synergies(df, n_components=2,
    max_components=3, max_iter=50_000) - scikit-learn's default solver ("cd") and init, not the mu kernels."""
    from sklearn.decomposition import NMF

    from muscle_synergies_b200.analysis import find_synergies, vaf

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = find_synergies(envelopes, n_components=2, max_components=3, max_iter=50_000)
        assert list(res.vaf_values.index) == [2, 3]
        assert list(res.vaf_values.columns) == ["All signals"] + list(envelopes.columns)
        for k in (2, 3):
            model = NMF(n_components=k, max_iter=50_000, tol=1e-6)
            W = model.fit_transform(envelopes)
            assert type(res.model[k]).__name__ == "NMF" and res.model[k].solver == "cd"
            assert np.allclose(res.components[k].to_numpy(), model.components_, rtol=0, atol=1e-9)
            want = vaf(envelopes, components=model.components_, transformed_signal=W)
            assert np.allclose(res.vaf_values.loc[[k]].to_numpy(), want.to_numpy(), rtol=0, atol=1e-9)
        single = find_synergies(envelopes, 2, max_iter=2000, init="nndsvd")
        assert single.components.shape == (2, 8) and single.model.init == "nndsvd"
    with pytest.raises(ValueError):
        find_synergies(envelopes, 0)
    with pytest.raises(ValueError):
        find_synergies(envelopes, 2, 9)


@pytest.mark.parametrize("kind", ["cubic", "nearest", "quadratic", 3])
def test_time_normalize_kinds_go_to_scipy(envelopes, kind):
    from scipy import interpolate

    from muscle_synergies_b200.emg import time_normalize

    got = time_normalize(envelopes, 57, kind=kind)
    f = interpolate.interp1d(np.linspace(0, 1, 200), envelopes, axis=0, copy=False, kind=kind, fill_value="extrapolate")
    want = f(np.linspace(0, 1, 57))
    assert list(got.columns) == list(envelopes.columns) and np.array_equal(got.index.to_numpy(), np.linspace(0, 1, 57))
    assert np.array_equal(got.to_numpy(), want)
