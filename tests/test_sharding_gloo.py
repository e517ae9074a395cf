"""The N>1 path on CPU: trials are partitioned over ranks with no data-path collective and the
small per-file results are gathered on the host (gloo, world size 2)."""
import os
import socket

import pytest

from muscle_synergies_b200.sharding import shard


def test_round_robin_and_balanced_partitions_cover_everything_once():
    files = [f"trial_{i}.csv" for i in range(11)]
    for world in (1, 2, 4, 8):
        parts = [shard(files, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(files)
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    sizes = [100, 1, 1, 1, 50, 50, 1, 1, 98, 1, 1]
    parts = [shard(files, r, 2, sizes) for r in range(2)]
    assert sorted(sum(parts, [])) == sorted(files)
    loads = [sum(sizes[files.index(f)] for f in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 10
    with pytest.raises(ValueError):
        shard(files, 2, 2)


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist

    from muscle_synergies_b200.sharding import gather_results, shard
    from oracle.vicon_oracle import load_vicon_file_oracle
    from tools.synth_vicon import synth_vicon

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seeds = list(range(5))
        mine = shard(seeds, rank, world)
        local = {}
        for seed in mine:
            # stands in for the per-rank GPU load: every rank handles only its own trials
            path = os.path.join(tmpdir, f"t{seed}.csv")
            synth_vicon(seed=seed, seconds=0.05 + 0.01 * seed, n_emg=4, n_markers=2).tofile(path)
            res = load_vicon_file_oracle(path)
            local[seed] = (res.num_frames, len(res.emg.rows))
        merged = {}
        for part in gather_results(local):
            assert not (set(part) & set(merged)), "a trial was processed by two ranks"
            merged.update(part)
        assert sorted(merged) == seeds
        for seed, (frames, rows) in merged.items():
            assert rows == frames * 20
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


def _fake_analyse(paths, seeds=None, **_kw):
    """CPU stand-in for `synergies_for_files` (the real one needs a GPU): one 'cycle' per file whose error
    depends on the seed, so that the best-of-restarts merge has something to decide."""
    import numpy as np

    from muscle_synergies_b200.pipeline import CycleSynergies, TrialSynergies
    from muscle_synergies_b200.segment import Cycle, Trecho

    for path in paths:
        if path.endswith("bad.csv"):
            yield path, ValueError("fewer than 40 transitions")
            continue
        base = sum(map(ord, os.path.basename(path)))
        errs = {int(s): ((base * 31 + int(s) * 17) % 101) / 100.0 for s in seeds}
        best = min(errs, key=lambda s: (errs[s], s))
        cyc = CycleSynergies(Trecho.FIRST, Cycle.FIRST, slice(None), {2: None}, {2: None}, {2: 7}, {2: errs[best]}, {2: best},
                             np.array([[1.0 - errs[best], 0.5]]), ["All signals", "m"])
        yield path, TrialSynergies([cyc])


def _sharded_worker(rank, world, port, paths, queue):
    import torch.distributed as dist

    from muscle_synergies_b200.pipeline import synergies_for_files_sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        merged = synergies_for_files_sharded(paths, analyse=_fake_analyse, n_restarts=6, random_state=10)
        queue.put((rank, [(r["file"], r.get("random_state"), r.get("reconstruction_err"), r.get("error")) for r in merged]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_files", [5, 1])
def test_sharded_pipeline_world_size_2_gloo(n_files):
    """End to end over two ranks: by file when there are enough files, by (rank, restart) when there are fewer
    files than ranks; one host-side gather; every rank ends with the same merged table as a single process."""
    import torch.multiprocessing as mp

    from muscle_synergies_b200.pipeline import synergies_for_files_sharded

    paths = [f"/nonexistent/t{i}.csv" for i in range(n_files)] + (["/nonexistent/bad.csv"] if n_files > 1 else [])
    single = synergies_for_files_sharded(paths, rank=0, world=1, analyse=_fake_analyse, n_restarts=6, random_state=10, gather=False)
    want = sorted((r["file"], r.get("random_state"), r.get("reconstruction_err"), r.get("error")) for r in single)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.SimpleQueue()
    mp.spawn(_sharded_worker, args=(2, port, paths, queue), nprocs=2, join=True)
    got = dict(queue.get() for _ in range(2))
    assert sorted(got) == [0, 1]
    for rank in (0, 1):
        assert sorted(got[rank], key=str) == sorted(want, key=str), rank


def test_cpulist_parsing():
    from muscle_synergies_b200.sharding import parse_cpulist

    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("5") == [5]
    assert parse_cpulist("") == []
