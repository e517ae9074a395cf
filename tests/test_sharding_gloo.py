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


def test_cpulist_parsing():
    from muscle_synergies_b200.sharding import parse_cpulist

    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("5") == [5]
    assert parse_cpulist("") == []
