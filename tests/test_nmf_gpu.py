"""NMF-MU extension stage against scikit-learn (the code the reference calls) and the numpy
oracle.  Not a bit-parity claim: the kernel computes in fp32, sklearn in fp64.

Stated tolerance: same init, same iteration count ->
  |VAF_gpu - VAF_sklearn| <= 1e-4 (overall and per muscle),
  | ||X-WH||_F(gpu) - ||X-WH||_F(sklearn) | / ||X||_F <= 1e-3.
"""
import warnings

import numpy as np
import pandas as pd
import pytest

from test_nmf_oracle import envelopes

pytestmark = pytest.mark.gpu

VAF_TOL = 1e-4
ERR_TOL = 1e-3


@pytest.fixture(scope="module")
def analysis():
    import __graft_entry__ as g

    g.build()
    from muscle_synergies_b200 import analysis

    return analysis


def sklearn_run(X, k, seed, max_iter, tol):
    from sklearn.decomposition import NMF

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = NMF(n_components=k, solver="mu", init="random", random_state=seed, max_iter=max_iter, tol=tol)
        W = model.fit_transform(X)
    return W, model.components_, model


def test_rank_sweep_with_restarts_matches_sklearn_at_fixed_iterations(analysis):
    """BASELINE configs[3]: k = 1..8 x 20 restarts on 200 x 16 envelopes, 200 iterations."""
    from oracle import nmf_oracle as no

    X = envelopes(1)
    ranks = [k for k in range(1, 9) for _ in range(20)]
    seeds = [r for _ in range(1, 9) for r in range(20)]
    res = analysis.nmf_mu_batched(X, ranks, seeds, max_iter=200, tol=0.0)
    xnorm = np.linalg.norm(X)
    assert (res.n_iter == 200).all()
    for p in range(0, len(ranks), 7):  # every 7th problem against sklearn, all against the invariants
        W, H, model = sklearn_run(X, ranks[p], seeds[p], 200, 0.0)
        want_all, want_cols = no.vaf(X, W, H)
        assert abs(res.vaf[p, 0] - want_all) <= VAF_TOL, (p, res.vaf[p, 0], want_all)
        assert np.abs(res.vaf[p, 1:] - want_cols).max() <= VAF_TOL
        assert abs(res.err[p] - model.reconstruction_err_) / xnorm <= ERR_TOL
        # the factors themselves track sklearn's closely at this iteration count
        assert np.abs(res.W[p] @ res.H[p] - W @ H).max() <= 5e-3
    for p in range(len(ranks)):
        assert res.W[p].shape == (200, ranks[p]) and res.H[p].shape == (ranks[p], 16)
        assert (res.W[p] >= 0).all() and (res.H[p] >= 0).all()
        got_all, got_cols = no.vaf(X, res.W[p], res.H[p])  # the fused VAF equals a recomputation from W, H
        assert abs(got_all - res.vaf[p, 0]) <= 2e-5 and np.abs(got_cols - res.vaf[p, 1:]).max() <= 2e-5
        assert abs(np.linalg.norm(X - res.W[p].astype(np.float64) @ res.H[p]) - res.err[p]) / xnorm <= 1e-5
    # more components never explain less (best restart per rank)
    best = [res.vaf[i * 20 : (i + 1) * 20, 0].max() for i in range(8)]
    assert all(b2 >= b1 - 1e-4 for b1, b2 in zip(best, best[1:]))


def test_convergence_criterion_tracks_sklearn(analysis):
    X = envelopes(2)
    for k, seed in [(2, 0), (3, 5), (6, 9)]:
        res = analysis.nmf_mu_batched(X, [k], [seed], max_iter=5000, tol=1e-5)
        W, H, model = sklearn_run(X, k, seed, 5000, 1e-5)
        assert res.n_iter[0] % 10 == 0 or res.n_iter[0] == 5000
        # fp32 vs fp64 error differences may move the stop by a few checks
        assert abs(int(res.n_iter[0]) - model.n_iter_) <= max(50, 0.2 * model.n_iter_)
        assert abs(res.err[0] - model.reconstruction_err_) / np.linalg.norm(X) <= ERR_TOL


def test_monotone_decrease_of_the_objective(analysis):
    X = envelopes(3)
    errs = [analysis.nmf_mu_batched(X, [4], [1], max_iter=it, tol=0.0).err[0] for it in (1, 5, 20, 100, 400)]
    assert all(b <= a * (1 + 1e-6) for a, b in zip(errs, errs[1:]))


def test_find_synergies_api(analysis):
    """Same call and result surface as the reference (analysis.py:713-914)."""
    X = envelopes(4)
    cols = [f"M{i}" for i in range(16)]
    df = pd.DataFrame(X, columns=cols)
    res = analysis.find_synergies(df, 2, 4, max_iter=300, tol=0.0, solver="mu", init="random", random_state=3)
    assert list(res.vaf_values.index) == [2, 3, 4]
    assert list(res.vaf_values.columns) == ["All signals"] + cols
    assert set(res.components) == {2, 3, 4} and res.components[3].shape == (3, 16)
    assert list(res.components[3].columns) == cols
    for k in (2, 3, 4):
        W, H, model = sklearn_run(X, k, 3, 300, 0.0)
        ref = analysis.vaf(df, transformed_signal=W, components=H)
        assert np.abs(res.vaf_values.loc[k].to_numpy() - ref.iloc[0].to_numpy()).max() <= VAF_TOL
        assert res.model[k].n_iter_ == 300 and res.model[k].components_.shape == (k, 16)
    single = analysis.find_synergies(df, 3, max_iter=100, tol=0.0, solver="mu", init="random", random_state=0)
    assert single.vaf_values.shape == (1, 17) and single.components.shape == (3, 16)
    multi = analysis.find_synergies(df, 3, max_iter=100, tol=0.0, n_restarts=8, solver="mu", init="random", random_state=0)
    assert multi.model.restarts.shape[0] == 8
    assert multi.model.reconstruction_err_ <= single.model.reconstruction_err_ + 1e-6
    with pytest.raises(ValueError):
        analysis.find_synergies(df, 0, solver="mu", init="random")
    with pytest.raises(ValueError):
        analysis.find_synergies(df, 3, 17, solver="mu", init="random")
    with pytest.raises(ValueError):
        analysis.find_synergies(df.iloc[:0], 3, solver="mu", init="random")
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        default = analysis.find_synergies(df, 3, max_iter=300)  # sklearn's default solver="cd": forwarded, as in the reference
    assert type(default.model).__name__ == "NMF" and default.components.shape == (3, df.shape[1])


def test_other_shapes(analysis):
    from oracle import nmf_oracle as no

    rng = np.random.default_rng(0)
    for n, m, k in [(50, 4, 3), (333, 8, 8), (1000, 16, 16), (64, 33, 5)]:
        X = rng.uniform(0, 1, (n, m))
        res = analysis.nmf_mu_batched(X, [k], [4], max_iter=60, tol=0.0)
        W0, H0 = no.random_init(X, k, 4)
        W, H, _ = no.mu(X, W0, H0, max_iter=60, tol=0.0)
        want_all, want_cols = no.vaf(X, W, H)
        assert abs(res.vaf[0, 0] - want_all) <= VAF_TOL and np.abs(res.vaf[0, 1:] - want_cols).max() <= 5 * VAF_TOL


def test_end_to_end_cycles_to_synergies(analysis):
    """BASELINE configs[4] in miniature: load -> segment -> per-cycle envelopes -> NMF sweep over
    every cycle in ONE launch; each problem checked against sklearn on the same envelope."""
    import muscle_synergies_b200 as ms
    from muscle_synergies_b200 import emg
    from muscle_synergies_b200.segment import Cycle, Segmenter, Trecho
    from oracle import nmf_oracle as no
    from tools.synth_vicon import synth_layout

    data = ms.load_vicon_bytes(synth_layout("D", seed=0), name="D")
    seg = Segmenter(data)
    windows = [seg.get_times_of(t, c) for t in Trecho for c in Cycle]
    X = emg.envelope_windows(data.emg, windows, window_size=0.05, reduce_to=200)  # (8 cycles, 200, 8 muscles)
    assert X.shape == (8, 200, 8) and float(X.min()) >= 0.0
    ranks, seeds, xi = [], [], []
    for cyc in range(8):
        for k in (1, 2, 3, 4):
            for r in range(3):
                ranks.append(k), seeds.append(r), xi.append(cyc)
    res = analysis.nmf_mu_batched(X, ranks, seeds, max_iter=150, tol=0.0, x_index=xi)
    Xh = X.cpu().numpy()
    for p in range(0, len(ranks), 5):
        W, H, _ = sklearn_run(Xh[xi[p]], ranks[p], seeds[p], 150, 0.0)
        want_all, want_cols = no.vaf(Xh[xi[p]], W, H)
        assert abs(res.vaf[p, 0] - want_all) <= VAF_TOL
        assert np.abs(res.vaf[p, 1:] - want_cols).max() <= 2 * VAF_TOL


@pytest.mark.parametrize("n,k", [(200, 3), (5000, 4), (60001, 8)])
def test_streaming_regime_matches_sklearn_and_the_resident_kernel(analysis, n, k):
    """Long signals: X and W stream from HBM every iteration (ms_nmf_mu_stream)."""
    from oracle import nmf_oracle as no

    X = envelopes(5, n=n)
    res = analysis.nmf_mu_batched(X, [k, k, 2], [0, 1, 0], max_iter=40, tol=0.0, regime="stream")
    assert (res.n_iter == 40).all()
    for p, (kk, seed) in enumerate([(k, 0), (k, 1), (2, 0)]):
        W, H, model = sklearn_run(X, kk, seed, 40, 0.0)
        want_all, want_cols = no.vaf(X, W, H)
        assert abs(res.vaf[p, 0] - want_all) <= VAF_TOL
        assert np.abs(res.vaf[p, 1:] - want_cols).max() <= 2 * VAF_TOL
        assert abs(res.err[p] - model.reconstruction_err_) / np.linalg.norm(X) <= ERR_TOL
    if n <= 1000:  # short enough for the shared-memory resident kernel too
        same = analysis.nmf_mu_batched(X, [k, k, 2], [0, 1, 0], max_iter=40, tol=0.0, regime="resident")
        assert np.abs(same.vaf - res.vaf).max() <= 2e-5


def test_streaming_regime_with_unaligned_factor_offsets(analysis):
    """A problem whose W starts at an odd float offset (after a rank-3 problem on 4999 rows) must take the
    scalar global-memory path of the streaming kernel, the aligned ones the float4 path - same answers."""
    from oracle import nmf_oracle as no

    X = envelopes(6, n=4999)
    ranks, seeds = [3, 4, 8, 4], [0, 1, 2, 3]
    res = analysis.nmf_mu_batched(X, ranks, seeds, max_iter=30, tol=0.0, regime="stream")
    for p, (kk, seed) in enumerate(zip(ranks, seeds)):
        W, H, model = sklearn_run(X, kk, seed, 30, 0.0)
        want_all, want_cols = no.vaf(X, W, H)
        assert abs(res.vaf[p, 0] - want_all) <= VAF_TOL
        assert np.abs(res.vaf[p, 1:] - want_cols).max() <= 2 * VAF_TOL
        assert abs(res.err[p] - model.reconstruction_err_) / np.linalg.norm(X) <= ERR_TOL


def test_streaming_regime_convergence_stop(analysis):
    X = envelopes(6, n=3000)
    res = analysis.nmf_mu_batched(X, [3], [2], max_iter=4000, tol=1e-5, regime="stream")
    W, H, model = sklearn_run(X, 3, 2, 4000, 1e-5)
    assert res.n_iter[0] % 10 == 0
    assert abs(int(res.n_iter[0]) - model.n_iter_) <= max(60, 0.25 * model.n_iter_)
    assert abs(res.err[0] - model.reconstruction_err_) / np.linalg.norm(X) <= ERR_TOL
