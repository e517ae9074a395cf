"""Trial -> cycles -> envelopes -> synergies on the GPU (BASELINE configs[4]) against the CPU chain:
oracle envelopes (the reference's own numpy/scipy calls) + scikit-learn's NMF, the solver the
reference calls.  Tolerances are the NMF stage's (tests/test_nmf_gpu.py): |dVAF| <= 1e-4,
relative reconstruction error difference <= 1e-3, at the same init and iteration count."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
VAF_TOL = 1e-4
ERR_TOL = 1e-3


@pytest.fixture(scope="module")
def ms():
    import __graft_entry__ as g

    g.build()
    import muscle_synergies_b200 as ms

    return ms


@pytest.fixture(scope="module")
def trial(ms):
    from tools.synth_vicon import synth_layout

    return ms.load_vicon_bytes(synth_layout("D", seed=0), name="D.csv")


def test_trial_synergies_against_the_cpu_chain(ms, trial):
    from sklearn.decomposition import NMF

    from muscle_synergies_b200.pipeline import cycle_windows, trial_synergies
    from muscle_synergies_b200.segment import Cycle, Segmenter, Trecho
    from oracle import emg_oracle as eo
    from oracle import nmf_oracle as no

    seg = Segmenter(trial)
    out = trial_synergies(trial, 1, 4, n_restarts=3, random_state=7, max_iter=200, tol=0.0, segmenter=seg, keep_batch=True)
    wins = cycle_windows(seg)
    assert [(c.trecho, c.cycle) for c in out.cycles] == [(t, c) for t in Trecho for c in Cycle]
    assert len(out.restarts) == 8 * 4 * 3 and (out.restarts["n_iter"] == 200).all()
    ranges = [(trial.emg.to_index(w).start, trial.emg.to_index(w).stop) for (_, _, w) in wins]
    env = eo.envelope_windows(trial.emg.df.to_numpy(), ranges, 1000, 200)
    np.testing.assert_allclose(out.envelopes.cpu().numpy(), env, rtol=1e-8, atol=1e-12)
    muscles = list(trial.emg.df.columns)
    for ci in (0, 3, 7):
        X = env[ci]
        xnorm = np.linalg.norm(X)
        cyc = out.cycles[ci]
        assert cyc.window == wins[ci][2]
        assert list(cyc.vaf_values.columns) == ["All signals"] + muscles and list(cyc.vaf_values.index) == [1, 2, 3, 4]
        for k in (1, 2, 3, 4):
            errs = {}
            for seed in (7, 8, 9):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    model = NMF(n_components=k, solver="mu", init="random", random_state=seed, max_iter=200, tol=0.0)
                    W = model.fit_transform(X)
                errs[seed] = (model.reconstruction_err_, W, model.components_)
                row = out.restarts[(out.restarts.trecho == cyc.trecho.value) & (out.restarts.cycle == cyc.cycle.value)
                                   & (out.restarts.n_components == k) & (out.restarts.random_state == seed)]
                assert len(row) == 1
                assert abs(float(row.reconstruction_err.iloc[0]) - model.reconstruction_err_) / xnorm <= ERR_TOL
                assert abs(float(row["All signals"].iloc[0]) - no.vaf(X, W, model.components_)[0]) <= VAF_TOL
            # the restart kept is (one of) the best of the CPU runs
            best_err = min(e for (e, _, _) in errs.values())
            assert abs(errs[cyc.random_state[k]][0] - best_err) / xnorm <= ERR_TOL
            _, Wb, Hb = errs[cyc.random_state[k]]
            want_all, want_cols = no.vaf(X, Wb, Hb)
            got = cyc.vaf_values.loc[k].to_numpy()
            assert abs(got[0] - want_all) <= VAF_TOL and np.abs(got[1:] - want_cols).max() <= VAF_TOL
            assert cyc.components[k].shape == (k, len(muscles)) and cyc.transformed[k].shape == (200, k)
            assert np.abs(cyc.transformed[k] @ cyc.components[k].to_numpy() - Wb @ Hb).max() <= 5e-3
    assert out[Trecho.SECOND, Cycle.FIRST] is out.cycles[2] and out[2, 1] is out.cycles[2]


def test_device_resident_input_equals_host_input(ms, trial):
    """nmf_mu_batched on a CUDA tensor (no host round trip) and on the same values as numpy."""
    from muscle_synergies_b200.analysis import nmf_mu_batched
    from muscle_synergies_b200.pipeline import cycle_windows
    from muscle_synergies_b200.segment import Segmenter

    env = ms.envelope_windows(trial.emg, [w for (_, _, w) in cycle_windows(Segmenter(trial))])
    ranks, seeds, xi = [2, 3, 5, 2], [0, 1, 2, 3], [0, 1, 7, 7]
    a = nmf_mu_batched(env, ranks, seeds, max_iter=100, tol=0.0, x_index=xi)
    b = nmf_mu_batched(env.cpu().numpy(), ranks, seeds, max_iter=100, tol=0.0, x_index=xi)
    for p in range(4):
        # the two differ only by the rounding of X.mean() in the init scale
        np.testing.assert_allclose(a.W[p], b.W[p], rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(a.H[p], b.H[p], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(a.vaf, b.vaf, atol=1e-5)


def test_synergies_for_files(ms, tmp_path):
    from muscle_synergies_b200.pipeline import synergies_for_files
    from tools.synth_vicon import synth_layout

    paths = []
    for seed in (0, 1):
        p = tmp_path / f"trial{seed}.csv"
        p.write_bytes(synth_layout("D", seed=seed).tobytes())
        paths.append(str(p))
    bad = tmp_path / "bad.csv"
    bad.write_bytes(b"not a vicon file\r\n")
    paths.insert(1, str(bad))
    got = list(synergies_for_files(paths, min_components=2, max_components=3, n_restarts=2, max_iter=50))
    assert [p for p, _ in got] == paths
    assert isinstance(got[1][1], Exception)
    for i in (0, 2):
        res = got[i][1]
        assert len(res.cycles) == 8
        v = res.cycles[0].vaf_values["All signals"]
        assert 0.0 < v.loc[2] <= v.loc[3] + 1e-4 <= 1.0 + 1e-4


def test_sharded_pipeline_matches_single_process(tmp_path):
    """synergies_for_files_sharded with the ranks played one after the other in this process: by file (3 files over
    2 ranks) and by (rank, restart) (1 file over 2 ranks).  The merged table equals the single-process one."""
    import numpy as np

    from muscle_synergies_b200.pipeline import merge_tables, synergies_for_files_sharded
    from tools.synth_vicon import synth_vicon

    paths = []
    for i in range(3):
        path = str(tmp_path / f"trial{i}.csv")
        synth_vicon(seed=300 + i, seconds=8.0 + i, n_emg=16, n_markers=4).tofile(path)
        paths.append(path)
    kw = dict(min_components=2, max_components=3, n_restarts=4, random_state=5, max_iter=60, tol=0.0)
    for subset in (paths, paths[:1]):
        single = synergies_for_files_sharded(subset, rank=0, world=1, gather=False, **kw)
        parts = [synergies_for_files_sharded(subset, rank=r, world=2, gather=False, **kw) for r in range(2)]
        if len(subset) >= 2:
            assert not ({row["file"] for row in parts[0]} & {row["file"] for row in parts[1]})  # a file has one owner
        else:
            assert parts[0] and parts[1]  # both ranks worked on the one file (different restarts)
        merged = merge_tables(parts)
        assert len(merged) == len(single) == len(subset) * 8 * 2
        for a, b in zip(merged, single):
            assert (a["file"], a["trecho"], a["cycle"], a["n_components"]) == (b["file"], b["trecho"], b["cycle"], b["n_components"])
            assert a["random_state"] == b["random_state"] and a["n_iter"] == b["n_iter"]
            assert np.array_equal(a["vaf"], b["vaf"]) and a["reconstruction_err"] == b["reconstruction_err"]
