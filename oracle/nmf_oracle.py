"""CPU oracle for the NMF multiplicative-update EXTENSION stage.  TEST INFRASTRUCTURE ONLY.

The reference calls scikit-learn (src/muscle_synergies/analysis.py:862-863:
`NMF(n_components=k, **kwargs).fit_transform(matrix)`); scikit-learn is a third-party
dependency (requirements.txt: scikit-learn>=0.21,<=0.24; installed here: 1.9.0) that is not
part of /root/reference.  The oracle is therefore sklearn itself, run live by the tests, plus
this numpy restatement of the algorithm it executes for solver="mu", beta_loss="frobenius",
init="random" (sklearn/decomposition/_nmf.py: _initialize_nmf, _multiplicative_update_w/_h,
_fit_multiplicative_update), which the tests pin against sklearn bit for bit in float64.

The reference has no test and no golden value for an MU run: NMF parity is "unpinned" by the
reference and anchored on sklearn only - which is why the stage is declared an extension.
"""
import numpy as np

EPSILON = np.finfo(np.float32).eps


def random_init(X, k, seed):
    """_initialize_nmf(init="random"): H is drawn first, then W."""
    X = np.asarray(X)
    avg = np.sqrt(X.mean() / k)
    rng = np.random.RandomState(seed)
    H = avg * rng.standard_normal(size=(k, X.shape[1])).astype(X.dtype, copy=False)
    W = avg * rng.standard_normal(size=(X.shape[0], k)).astype(X.dtype, copy=False)
    np.abs(H, out=H)
    np.abs(W, out=W)
    return W, H


def frobenius(X, W, H):
    d = X - W @ H
    return np.sqrt(np.sum(d * d))


def mu(X, W, H, max_iter=200, tol=1e-4, check_every=10):
    """Returns (W, H, n_iter) after sklearn's MU loop; arithmetic in the dtype of X."""
    X = np.asarray(X)
    W = W.copy()
    H = H.copy()
    err0 = prev = frobenius(X, W, H)
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        den = W @ (H @ H.T)
        den[den == 0] = EPSILON
        W *= (X @ H.T) / den
        den = (W.T @ W) @ H
        den[den == 0] = EPSILON
        H *= (W.T @ X) / den
        if tol > 0 and n_iter % check_every == 0:
            err = frobenius(X, W, H)
            if (prev - err) / err0 < tol:
                break
            prev = err
    return W, H, n_iter


def vaf(X, W, H):
    """analysis.py:642-667: overall and per-column 1 - SS_res / SS_tot (uncentred)."""
    X = np.asarray(X, dtype=np.float64)
    R = X - np.asarray(W, dtype=np.float64) @ np.asarray(H, dtype=np.float64)
    overall = 1.0 - np.sum(R * R) / np.sum(X * X)
    cols = 1.0 - np.sum(R * R, axis=0) / np.sum(X * X, axis=0)
    return overall, cols
