"""Pins oracle/nmf_oracle.py (numpy restatement of sklearn's MU solver) against scikit-learn
itself - the third-party code the reference calls (analysis.py:862-863)."""
import warnings

import numpy as np
import pytest

from oracle import nmf_oracle as no


def envelopes(seed, n=200, m=16, k_true=4):
    rng = np.random.default_rng(seed)
    t = np.linspace(0, 1, n)[:, None]
    basis = np.abs(np.sin(np.pi * (rng.uniform(0.5, 3, (1, k_true)) * t + rng.uniform(0, 1, (1, k_true))))) ** 2
    X = basis @ rng.uniform(0, 1, (k_true, m)) + 0.02 * rng.uniform(0, 1, (n, m))
    return X / X.max(axis=0)


@pytest.mark.parametrize("k,seed", [(1, 0), (2, 3), (5, 7), (8, 11)])
def test_numpy_restatement_equals_sklearn_bit_for_bit(k, seed):
    from sklearn.decomposition import NMF

    X = envelopes(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = NMF(n_components=k, solver="mu", init="random", random_state=seed, max_iter=300, tol=1e-5)
        W_ref = model.fit_transform(X)
    W0, H0 = no.random_init(X, k, seed)
    W, H, n_iter = no.mu(X, W0, H0, max_iter=300, tol=1e-5)
    assert n_iter == model.n_iter_
    assert np.array_equal(W, W_ref) and np.array_equal(H, model.components_)
    assert np.isclose(no.frobenius(X, W, H), model.reconstruction_err_, rtol=1e-12)
