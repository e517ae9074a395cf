"""Muscle-synergy extraction on the GPU: batched NMF by multiplicative updates (EXTENSION).

API mirror of the reference's `vaf`, `SynergyRunResult` and `find_synergies`
(src/muscle_synergies/analysis.py:597-667, 670-710, 713-914).  The reference hands the
factorisation to `sklearn.decomposition.NMF(n_components=k, **kwargs)` (:862-863); here the
solver="mu", beta_loss="frobenius", init="random" case runs as ONE CUDA launch for the whole
rank sweep x restarts grid (csrc/ms_nmf.cu), in fp32, with sklearn's initialisation
(`RandomState(seed)`: H drawn first, then W - sklearn/decomposition/_nmf.py:_initialize_nmf)
generated on the host so that a run is comparable to `NMF(solver="mu", init="random",
random_state=seed)` at the same iteration count.  Other solvers / inits raise
NotImplementedError: this stage is an extension, not a replacement of scikit-learn.
"""
import ctypes
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Mapping, Optional, Sequence, Union

import numpy as np
import pandas

from . import _native as nat


# ---- batched solver ---------------------------------------------------------------------------------
@dataclass
class NMFBatchResult:
    ranks: np.ndarray          # (P,)
    seeds: np.ndarray          # (P,)
    W: list                    # P arrays (n, k_p) float32
    H: list                    # P arrays (k_p, m) float32
    n_iter: np.ndarray         # (P,)
    err: np.ndarray            # (P,)  ||X - W H||_F
    vaf: np.ndarray            # (P, m + 1): overall, then per column


def sklearn_random_init(X: np.ndarray, k: int, seed: int):
    """`_initialize_nmf(X, k, init="random", random_state=seed)` of scikit-learn."""
    avg = np.sqrt(X.mean() / k)
    rng = np.random.RandomState(seed)
    H = avg * rng.standard_normal(size=(k, X.shape[1])).astype(X.dtype, copy=False)
    W = avg * rng.standard_normal(size=(X.shape[0], k)).astype(X.dtype, copy=False)
    np.abs(H, out=H)
    np.abs(W, out=W)
    return W, H


def nmf_mu_batched(X, ranks: Sequence[int], seeds: Sequence[int], max_iter: int = 200, tol: float = 1e-4,
                   check_every: int = 10, init=None, device=None, x_index: Optional[Sequence[int]] = None,
                   regime: Optional[str] = None) -> NMFBatchResult:
    """Runs len(ranks) MU factorisations in one kernel launch.

    X: non-negative (n samples x m muscles), or a stack (B, n, m) of such matrices (e.g. one per
    gait cycle) with x_index[p] naming the matrix of problem p.  ranks[p], seeds[p]: rank and
    sklearn `random_state` of problem p; `init` optionally gives the initial (W, H) pairs
    instead of the seeds.  regime: None (by size), "resident" or "stream"."""
    import torch

    lib = nat.lib()
    if not torch.cuda.is_available():
        raise nat.NativeError("nmf_mu_batched needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if isinstance(X, torch.Tensor):
        Xh = X.detach().to("cpu", torch.float64).numpy()
    else:
        Xh = np.asarray(X, dtype=np.float64)
    if Xh.ndim == 2:
        Xh = Xh[None]
    if Xh.ndim != 3 or Xh.size == 0:
        raise ValueError("X must be a non-empty (n, m) matrix or a (B, n, m) stack")
    if (Xh < 0).any():
        raise ValueError("Negative values in data passed to NMF (input X)")
    _, n, m = Xh.shape
    ranks = np.asarray(ranks, dtype=np.int32)
    seeds = np.asarray(seeds, dtype=np.int64)
    P = len(ranks)
    xi = np.zeros(P, dtype=np.int32) if x_index is None else np.asarray(x_index, dtype=np.int32)
    if len(xi) != P or len(seeds) != P or xi.min() < 0 or xi.max() >= Xh.shape[0]:
        raise ValueError("ranks, seeds and x_index must have one entry per problem")
    kmax = int(ranks.max())
    # short signals stay resident in shared memory; long ones stream from HBM every iteration
    resident = (n <= int(lib.ms_nmf_resident_max_rows(m, kmax))) if regime is None else (regime == "resident")
    w_parts, h_parts = [], []
    for p in range(P):
        if init is not None:
            W0, H0 = init[p]
        else:
            W0, H0 = sklearn_random_init(Xh[xi[p]], int(ranks[p]), int(seeds[p]))  # float64 draws, like sklearn
        w_parts.append(np.ascontiguousarray(W0, dtype=np.float32).ravel())
        h_parts.append(np.ascontiguousarray(H0, dtype=np.float32).ravel())
    stream = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        dX = torch.from_numpy(np.ascontiguousarray(Xh, dtype=np.float32)).to(dev)
        dW = torch.from_numpy(np.concatenate(w_parts)).to(dev)
        dH = torch.from_numpy(np.concatenate(h_parts)).to(dev)
        work = torch.empty(max(P * 32, int(lib.ms_nmf_stream_workspace_bytes(m, P))), dtype=torch.uint8, device=dev)
        d_iter = torch.empty(P, dtype=torch.int32, device=dev)
        d_err = torch.empty(P, dtype=torch.float32, device=dev)
        d_vaf = torch.empty((P, m + 1), dtype=torch.float32, device=dev)
        h_ranks = (ctypes.c_int32 * P)(*[int(k) for k in ranks])
        h_xi = (ctypes.c_int32 * P)(*[int(v) for v in xi])
        entry = lib.ms_nmf_mu_batched if resident else lib.ms_nmf_mu_stream
        nat.check(
            entry(
                dX.data_ptr(), n, m, h_ranks, h_xi, P, dW.data_ptr(), dH.data_ptr(), int(max_iter), ctypes.c_float(tol),
                int(check_every), work.data_ptr(), d_iter.data_ptr(), d_err.data_ptr(), d_vaf.data_ptr(),
                ctypes.c_void_p(stream.cuda_stream),
            ),
            "ms_nmf_mu_batched" if resident else "ms_nmf_mu_stream",
        )
        Wall, Hall = dW.cpu().numpy(), dH.cpu().numpy()
        n_iter, err, vafs = d_iter.cpu().numpy(), d_err.cpu().numpy(), d_vaf.cpu().numpy()
    Ws, Hs = [], []
    wo = ho = 0
    for k in ranks:
        k = int(k)
        Ws.append(Wall[wo : wo + n * k].reshape(n, k))
        Hs.append(Hall[ho : ho + k * m].reshape(k, m))
        wo += n * k
        ho += k * m
    return NMFBatchResult(ranks, seeds, Ws, Hs, n_iter, err, vafs)


# ---- reference API ------------------------------------------------------------------------------------
def vaf(original_df: pandas.DataFrame, transformed_signal=None, components=None, reconstructed_signal=None) -> pandas.DataFrame:
    """Variance accounted for, overall and per muscle (analysis.py:597-667)."""
    if reconstructed_signal is None:
        reconstructed_signal = transformed_signal @ components
    error = original_df - reconstructed_signal

    def ss(arr, axis):
        return np.sum(arr.to_numpy() ** 2, axis=axis)

    overall = 1 - ss(error, (0, 1)) / ss(original_df, (0, 1))
    per_column = 1 - ss(error, 0) / ss(original_df, 0)
    labels = ["All signals"] + original_df.columns.tolist()
    values = [overall] + list(per_column.reshape(-1))
    return pandas.DataFrame({lbl: [val] for (lbl, val) in zip(labels, values)})


@dataclass
class NMFModel:
    """What the reference keeps of a fitted `sklearn.decomposition.NMF`."""

    n_components: int
    components_: np.ndarray
    n_iter_: int
    reconstruction_err_: float
    n_features_in_: int
    random_state: int
    solver: str = "mu"
    beta_loss: str = "frobenius"
    init: str = "random"
    restarts: Optional[pandas.DataFrame] = None  # extension: one row per restart (seed, n_iter, err, VAF)

    def transform(self, X):  # pragma: no cover - kept for API familiarity
        raise NotImplementedError("transform of new data is not part of the accelerated path")


@dataclass
class SynergyRunResult:
    vaf_values: pandas.DataFrame
    components: Union[pandas.DataFrame, Mapping[int, pandas.DataFrame]]
    model: Union[NMFModel, Mapping[int, NMFModel]]
    transformed: Union[np.ndarray, Mapping[int, np.ndarray], None] = field(default=None, repr=False)


def find_synergies(
    processed_emg_df: pandas.DataFrame,
    n_components: int,
    max_components: Optional[int] = None,
    *,
    max_iter: int = 100_000,
    tol: float = 1e-6,
    n_restarts: int = 1,
    **sklearn_kwargs,
) -> SynergyRunResult:
    """Find muscle synergies with NMF (analysis.py:713-914) - all ranks and restarts in one launch.

    Accepted scikit-learn keywords: solver="mu", init="random", beta_loss="frobenius" (or 2),
    random_state=int.  `n_restarts` (extension) runs seeds random_state .. random_state+R-1 per
    rank and keeps, per rank, the restart with the smallest reconstruction error."""
    if processed_emg_df.empty:
        raise ValueError("empty EMG DataFrame")
    num_features = len(processed_emg_df.columns)
    if n_components < 1 or n_components > num_features:
        raise ValueError("invalid number of components")
    if max_components is not None and (max_components < n_components or max_components > num_features):
        raise ValueError("invalid number of components")
    kw = dict(sklearn_kwargs)
    solver = kw.pop("solver", "cd")
    init = kw.pop("init", None)
    beta_loss = kw.pop("beta_loss", "frobenius")
    seed = kw.pop("random_state", None)
    if solver != "mu" or init != "random" or beta_loss not in ("frobenius", 2, 2.0):
        raise NotImplementedError(
            'the CUDA stage implements NMF(solver="mu", init="random", beta_loss="frobenius") only; '
            "pass those keywords (the reference forwards them to scikit-learn)"
        )
    if kw:
        raise NotImplementedError(f"unsupported NMF keywords: {sorted(kw)}")
    if seed is None:
        seed = int(np.random.randint(0, 2**31 - 1))
    ranks_sweep = [n_components] if max_components is None else list(range(n_components, max_components + 1))
    ranks = [k for k in ranks_sweep for _ in range(n_restarts)]
    seeds = [int(seed) + r for _ in ranks_sweep for r in range(n_restarts)]
    X = processed_emg_df.to_numpy(dtype=np.float64)
    res = nmf_mu_batched(X, ranks, seeds, max_iter=max_iter, tol=tol)

    columns = processed_emg_df.columns
    labels = ["All signals"] + columns.tolist()
    per_rank = OrderedDict()
    for i, k in enumerate(ranks_sweep):
        sl = slice(i * n_restarts, (i + 1) * n_restarts)
        best = i * n_restarts + int(np.argmin(res.err[sl]))
        table = pandas.DataFrame(res.vaf[sl].astype(np.float64), columns=labels)
        table.insert(0, "reconstruction_err", res.err[sl].astype(np.float64))
        table.insert(0, "n_iter", res.n_iter[sl])
        table.insert(0, "random_state", res.seeds[sl])
        model = NMFModel(k, res.H[best].astype(np.float64), int(res.n_iter[best]), float(res.err[best]),
                         num_features, int(res.seeds[best]), restarts=table)
        transformed = res.W[best].astype(np.float64)
        comps = pandas.DataFrame(model.components_, columns=columns)
        vaf_values = vaf(processed_emg_df, components=model.components_, transformed_signal=transformed)
        per_rank[k] = SynergyRunResult(vaf_values, comps, model, transformed)
    if max_components is None:
        return per_rank[n_components]
    vaf_values = pandas.concat([r.vaf_values for r in per_rank.values()])
    vaf_values.set_index(np.array(tuple(per_rank.keys())), inplace=True)
    return SynergyRunResult(
        vaf_values,
        {k: r.components for k, r in per_rank.items()},
        {k: r.model for k, r in per_rank.items()},
        {k: r.transformed for k, r in per_rank.items()},
    )
