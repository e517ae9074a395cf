"""Muscle-synergy extraction on the GPU: batched NMF by multiplicative updates (EXTENSION).

API mirror of the reference's `vaf`, `SynergyRunResult` and `find_synergies`
(src/muscle_synergies/analysis.py:597-667, 670-710, 713-914).  The reference hands the
factorisation to `sklearn.decomposition.NMF(n_components=k, **kwargs)` (:862-863); here the
solver="mu", beta_loss="frobenius", init="random" case runs as ONE CUDA launch for the whole
rank sweep x restarts grid (csrc/ms_nmf.cu), in fp32, with sklearn's initialisation
(`RandomState(seed)`: H drawn first, then W - sklearn/decomposition/_nmf.py:_initialize_nmf)
generated on the host so that a run is comparable to `NMF(solver="mu", init="random",
random_state=seed)` at the same iteration count.  Every other call (sklearn's default solver="cd",
nndsvd inits, regularisation ...) is forwarded to scikit-learn, exactly as the reference does
(analysis.py:848-864): this stage is an extension, not a replacement of scikit-learn.
"""
import ctypes
import functools
import threading
from collections import OrderedDict
from dataclasses import dataclass, field
from collections.abc import Sequence
from typing import Mapping, Optional, Union

import numpy as np
import pandas

from . import _native as nat


# ---- batched solver ---------------------------------------------------------------------------------
@dataclass
class NMFBatchResult:
    ranks: np.ndarray          # (P,)
    seeds: np.ndarray          # (P,)
    W: Sequence                # P arrays (n, k_p) float32 (views of one packed array, made on access)
    H: Sequence                # P arrays (k_p, m) float32
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


def _unit_draws(n: int, m: int, k: int, seed: int):
    """|standard_normal| draws of sklearn's random init for (seed, k) on an n x m matrix: H first,
    then W.  The init itself is these times sqrt(X.mean() / k) (|a z| == a |z| exactly for a > 0)."""
    rng = np.random.RandomState(seed)
    H = np.abs(rng.standard_normal(size=(k, m)))
    W = np.abs(rng.standard_normal(size=(n, k)))
    return W.ravel(), H.ravel()


_unit_draws_cached = functools.lru_cache(maxsize=1024)(_unit_draws)  # gait-cycle sized problems only
_draw_cache = OrderedDict()  # (n, m, ranks, seeds, device) -> unit draws of the whole batch, on the device


def _batch_draws(torch, n, m, ranks, seeds, dev):
    with _cache_lock:
        return _batch_draws_locked(torch, n, m, ranks, seeds, dev)


def _batch_draws_locked(torch, n, m, ranks, seeds, dev):
    key = (n, m, ranks.tobytes(), seeds.tobytes(), str(dev))
    hit = _draw_cache.get(key)
    if hit is None:
        draws = _unit_draws_cached if n * m <= 1 << 16 else _unit_draws
        parts = [draws(n, m, int(k), int(s)) for k, s in zip(ranks, seeds)]
        hit = (torch.from_numpy(np.concatenate([p[0] for p in parts])).to(dev),
               torch.from_numpy(np.concatenate([p[1] for p in parts])).to(dev))
        if n * m > 1 << 16:
            return hit
        _draw_cache[key] = hit
        while len(_draw_cache) > 8:
            _draw_cache.popitem(last=False)
    else:
        _draw_cache.move_to_end(key)
    return hit


_plan_cache = OrderedDict()  # (n, m, ranks, x_index, device) -> (device problem table, kmax, device ranks, x_index, element -> problem maps)


_cache_lock = threading.Lock()  # the two caches below are process-wide; batches may be launched from several threads


def _batch_plan(torch, lib, n, m, ranks, xi, dev):
    with _cache_lock:
        return _batch_plan_locked(torch, lib, n, m, ranks, xi, dev)


def _batch_plan_locked(torch, lib, n, m, ranks, xi, dev):
    """The device-side description of a sweep (ms_nmf_plan's table, the ranks and matrix indices as int64 tensors), kept
    for the sweeps a caller repeats: a pipeline runs the same one for every trial, and each copy from pageable memory
    would wait for everything queued on the stream - the factorisation of the trial before."""
    key = (n, m, ranks.tobytes(), xi.tobytes(), str(dev))
    hit = _plan_cache.get(key)
    if hit is None:
        table = np.empty(len(ranks) * 32, dtype=np.uint8)
        kmax = int(lib.ms_nmf_plan(n, m, ranks.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                   xi.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(ranks), table.ctypes.data))
        nat.check(min(kmax, 0), "ms_nmf_plan")
        r64 = ranks.astype(np.int64)
        problem = np.arange(len(ranks), dtype=np.int64)
        # which problem every element of the packed W / H belongs to (the per-problem scale of the random init is
        # gathered through these; torch.repeat_interleave with device counts would wait for the device to size its output)
        hit = (torch.from_numpy(table).to(dev), kmax, torch.from_numpy(r64).to(dev), torch.from_numpy(xi.astype(np.int64)).to(dev),
               torch.from_numpy(np.repeat(problem, r64 * n)).to(dev), torch.from_numpy(np.repeat(problem, r64 * m)).to(dev))
        _plan_cache[key] = hit
        while len(_plan_cache) > 8:
            _plan_cache.popitem(last=False)
    else:
        _plan_cache.move_to_end(key)
    return hit


_pinned_pool = {}  # nbytes -> pinned uint8 tensors not on loan (results of deferred runs travel through them)


def _pinned_take(torch, nbytes: int):
    free = _pinned_pool.get(nbytes)
    return free.pop() if free else torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)


def _pinned_give(buf):
    free = _pinned_pool.setdefault(int(buf.numel()), [])
    if len(free) < 4:
        free.append(buf)


class PendingNMF:
    """A batch whose kernel is queued and whose results are on their way to pinned host memory: `result()` waits
    for the copies and builds the NMFBatchResult.  Lets a caller queue the next trial's GPU work before it
    looks at this one's (pipeline.synergies_for_files)."""

    def __init__(self, torch, stream, ranks, seeds, n, m, device_arrays, keep, negative=None):
        self._ranks, self._seeds, self._n, self._m = ranks, seeds, n, m
        self._keep = keep  # inputs of the kernel: alive until the copies below have run
        self._bufs = []
        self._negative = negative is not None  # a device flag "X has a negative entry" travels with the results
        if negative is not None:
            device_arrays = list(device_arrays) + [negative.to(torch.uint8).reshape(1)]
        for t in device_arrays:
            t = t.contiguous()
            buf = _pinned_take(torch, int(t.numel()) * t.element_size())
            buf.view(t.dtype).copy_(t.reshape(-1), non_blocking=True)
            self._bufs.append((buf, t.dtype, tuple(t.shape)))
            self._keep.append(t)
        self._event = torch.cuda.Event()
        self._event.record(stream)
        self._result = None

    def result(self) -> NMFBatchResult:
        if self._result is None:
            self._event.synchronize()
            out = []
            for buf, dtype, shape in self._bufs:
                out.append(buf.view(dtype).numpy().reshape(shape).copy())  # own memory: the pinned buffer goes back
                _pinned_give(buf)
            self._bufs, self._keep = [], []
            if self._negative and out.pop()[0]:
                raise ValueError("Negative values in data passed to NMF (input X)")
            self._result = _assemble(self._ranks, self._seeds, self._n, self._m, *out)
        return self._result


class _Factors(Sequence):
    """The factors of a batch, problem after problem in one array: item p is a view, made when it is asked for (a
    trial's batch has 1280 problems and its tables read 64 of them)."""

    def __init__(self, packed: np.ndarray, offsets: np.ndarray, shapes):
        self._packed, self._offsets, self._shapes = packed, offsets, shapes

    def __len__(self):
        return len(self._offsets) - 1

    def __getitem__(self, p):
        if isinstance(p, slice):
            return [self[i] for i in range(*p.indices(len(self)))]
        p = int(p)
        if p < 0:
            p += len(self)
        if not 0 <= p < len(self):
            raise IndexError(p)
        return self._packed[self._offsets[p] : self._offsets[p + 1]].reshape(self._shapes(p))


def _assemble(ranks, seeds, n, m, Wall, Hall, n_iter, err, vafs) -> NMFBatchResult:
    w_off = np.concatenate([[0], np.cumsum(ranks.astype(np.int64) * n)])
    h_off = np.concatenate([[0], np.cumsum(ranks.astype(np.int64) * m)])
    Ws = _Factors(Wall, w_off, lambda p: (n, int(ranks[p])))
    Hs = _Factors(Hall, h_off, lambda p: (int(ranks[p]), m))
    return NMFBatchResult(ranks, seeds, Ws, Hs, n_iter, err, vafs)


def nmf_mu_batched(X, ranks: Sequence[int], seeds: Sequence[int], max_iter: int = 200, tol: float = 1e-4,
                   check_every: int = 10, init=None, device=None, x_index: Optional[Sequence[int]] = None,
                   regime: Optional[str] = None, defer: bool = False) -> Union[NMFBatchResult, PendingNMF]:
    """Runs len(ranks) MU factorisations in one kernel launch.  defer=True returns a PendingNMF right after the
    launch (`.result()` gives the NMFBatchResult).

    X: non-negative (n samples x m muscles), or a stack (B, n, m) of such matrices (e.g. one per
    gait cycle) with x_index[p] naming the matrix of problem p; a numpy array, or a CUDA tensor
    (e.g. the output of `envelope_windows`), which then never leaves the device.  ranks[p], seeds[p]:
    rank and sklearn `random_state` of problem p; `init` optionally gives the initial (W, H) pairs
    instead of the seeds.  regime: None (by size), "resident" or "stream"."""
    import torch

    lib = nat.lib()
    if not torch.cuda.is_available():
        raise nat.NativeError("nmf_mu_batched needs a CUDA device; there is no CPU fallback")
    on_device = isinstance(X, torch.Tensor) and X.is_cuda
    if on_device:
        dev = X.device
        dX64 = X.detach().to(torch.float64)
        if dX64.ndim == 2:
            dX64 = dX64[None]
        shape = tuple(dX64.shape)
    else:
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        Xh = X.detach().to("cpu", torch.float64).numpy() if isinstance(X, torch.Tensor) else np.asarray(X, dtype=np.float64)
        if Xh.ndim == 2:
            Xh = Xh[None]
        shape = Xh.shape
    if len(shape) != 3 or 0 in shape:
        raise ValueError("X must be a non-empty (n, m) matrix or a (B, n, m) stack")
    # a deferred run on device data does not wait for the answer here (it would wait for everything queued before it,
    # the previous trial's factorisation included): the flag travels with the results and `.result()` raises
    negative = (dX64 < 0).any() if on_device and defer else None
    if negative is None and (bool((dX64 < 0).any()) if on_device else bool((Xh < 0).any())):
        raise ValueError("Negative values in data passed to NMF (input X)")
    B, n, m = shape
    ranks = np.ascontiguousarray(ranks, dtype=np.int32)
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    P = len(ranks)
    xi = np.zeros(P, dtype=np.int32) if x_index is None else np.ascontiguousarray(x_index, dtype=np.int32)
    if P < 1 or len(xi) != P or len(seeds) != P or xi.min() < 0 or xi.max() >= B:
        raise ValueError("ranks, seeds and x_index must have one entry per problem")
    if ranks.min() < 1:
        raise ValueError("ranks must be positive")
    kmax = int(ranks.max())
    # short signals stay resident in shared memory; long ones stream from HBM every iteration
    resident = (n <= int(lib.ms_nmf_resident_max_rows(m, kmax))) if regime is None else (regime == "resident")
    stream = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        if on_device:
            dX = dX64.to(torch.float32).contiguous()
        else:
            dX = torch.from_numpy(np.ascontiguousarray(Xh, dtype=np.float32)).to(dev)
        if init is not None:
            dW = torch.from_numpy(np.concatenate(
                [np.ascontiguousarray(init[p][0], dtype=np.float32).ravel() for p in range(P)])).to(dev)
            dH = torch.from_numpy(np.concatenate(
                [np.ascontiguousarray(init[p][1], dtype=np.float32).ravel() for p in range(P)])).to(dev)
        else:
            # sklearn's init="random": float64 draws scaled by sqrt(X.mean() / k), then the kernel's fp32
            unit_w, unit_h = _batch_draws(torch, n, m, ranks, seeds, dev)
            if on_device:
                means = dX64.mean(dim=(1, 2))
            else:
                means = torch.from_numpy(np.array([Xh[b].mean() for b in range(B)])).to(dev)
            _, _, d_ranks, d_xi, w_problem, h_problem = _batch_plan(torch, lib, n, m, ranks, xi, dev)
            avg = torch.sqrt(means[d_xi] / d_ranks)
            dW = (unit_w * avg[w_problem]).to(torch.float32)
            dH = (unit_h * avg[h_problem]).to(torch.float32)
        d_iter = torch.empty(P, dtype=torch.int32, device=dev)
        d_err = torch.empty(P, dtype=torch.float32, device=dev)
        d_vaf = torch.empty((P, m + 1), dtype=torch.float32, device=dev)
        if resident:
            # the problem table lives on the device for as long as the sweep is repeated: the launch copies nothing
            # and waits for nothing
            work, kmax_plan = _batch_plan(torch, lib, n, m, ranks, xi, dev)[:2]
            nat.check(
                lib.ms_nmf_mu_batched_planned(
                    dX.data_ptr(), n, m, work.data_ptr(), P, kmax_plan, dW.data_ptr(), dH.data_ptr(), int(max_iter),
                    ctypes.c_float(tol), int(check_every), d_iter.data_ptr(), d_err.data_ptr(), d_vaf.data_ptr(),
                    ctypes.c_void_p(stream.cuda_stream),
                ),
                "ms_nmf_mu_batched_planned",
            )
        else:
            work = torch.empty(int(lib.ms_nmf_stream_workspace_bytes(m, P)), dtype=torch.uint8, device=dev)
            h_ranks = ranks.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
            h_xi = xi.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
            nat.check(
                lib.ms_nmf_mu_stream(
                    dX.data_ptr(), n, m, h_ranks, h_xi, P, dW.data_ptr(), dH.data_ptr(), int(max_iter), ctypes.c_float(tol),
                    int(check_every), work.data_ptr(), d_iter.data_ptr(), d_err.data_ptr(), d_vaf.data_ptr(),
                    ctypes.c_void_p(stream.cuda_stream),
                ),
                "ms_nmf_mu_stream",
            )
        if defer:
            return PendingNMF(torch, stream, ranks, seeds, n, m, [dW, dH, d_iter, d_err, d_vaf], [dX, work], negative)
        Wall, Hall = dW.cpu().numpy(), dH.cpu().numpy()
        n_iter, err, vafs = d_iter.cpu().numpy(), d_err.cpu().numpy(), d_vaf.cpu().numpy()
    return _assemble(ranks, seeds, n, m, Wall, Hall, n_iter, err, vafs)


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


def _find_synergies_sklearn(processed_emg_df, n_components, max_components, **sklearn_kwargs) -> "SynergyRunResult":
    """find_synergies as the reference runs it (analysis.py:848-914): one sklearn.decomposition.NMF per rank."""
    from sklearn.decomposition import NMF

    def single(k):
        model = NMF(n_components=k, **sklearn_kwargs)
        transformed = model.fit_transform(processed_emg_df)
        values = vaf(processed_emg_df, components=model.components_, transformed_signal=transformed)
        comps = pandas.DataFrame(model.components_, columns=processed_emg_df.columns)
        return SynergyRunResult(values, comps, model, transformed)

    if max_components is None:
        return single(n_components)
    runs = OrderedDict((k, single(k)) for k in range(n_components, max_components + 1))
    vaf_values = pandas.concat([r.vaf_values for r in runs.values()])
    vaf_values.set_index(np.array(tuple(runs.keys())), inplace=True)
    return SynergyRunResult(vaf_values, {k: r.components for k, r in runs.items()}, {k: r.model for k, r in runs.items()},
                            {k: r.transformed for k, r in runs.items()})


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

    solver="mu", init="random", beta_loss="frobenius" (or 2), random_state=int run on the GPU, all ranks and
    restarts in one launch; `n_restarts` (extension) runs seeds random_state .. random_state+R-1 per rank and keeps,
    per rank, the restart with the smallest reconstruction error.  Any other scikit-learn configuration (the
    default solver="cd", nndsvd inits, regularisation ...) is forwarded to scikit-learn as the reference does."""
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
    if solver != "mu" or init != "random" or beta_loss not in ("frobenius", 2, 2.0) or kw:
        # Everything the batched kernels do not implement is what the reference itself does with it: the keywords go
        # to scikit-learn unchanged (analysis.py:848-864) - solver="cd" with an nndsvd* init is the tutorial's own call
        # (docs/source/tutorials/Finding muscle synergies.ipynb, cell 26).  This is the reference's code path, not a
        # fallback of the accelerated one: solver="mu", init="random" never comes here.
        if n_restarts != 1:
            raise ValueError("n_restarts is an extension of the batched mu solver; scikit-learn runs one initialisation")
        return _find_synergies_sklearn(processed_emg_df, n_components, max_components, max_iter=max_iter, tol=tol,
                                       **sklearn_kwargs)
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
