"""Host-side multi-process logic on CPU (gloo, world_size 2): pair sharding, stream sharding with a
one-frame halo, and the single all-reduce of the accumulator vector (models/model_interface.py:124-137)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from keypoint_bench_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_shard_pairs_partitions_every_pair_once():
    for n, world in ((10, 2), (64, 8), (7, 3), (0, 4)):
        seen = sorted(i for r in range(world) for i in parallel.shard_pairs(n, r, world))
        assert seen == list(range(n))


def test_shard_stream_covers_every_consecutive_pair_once():
    for frames, world in ((100, 8), (9, 2), (3, 4), (1, 2), (0, 2)):
        pairs = []
        for r in range(world):
            lo, hi = parallel.shard_stream(frames, r, world)
            pairs += [(t - 1, t) for t in range(lo + 1, hi)]
        assert sorted(pairs) == [(t - 1, t) for t in range(1, frames)]


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    r, w = parallel.init('gloo')
    assert (r, w) == (rank, world)
    # every rank owns pairs i = rank mod world; per-pair values are i/10 (repeatability) and i (error)
    mine = list(parallel.shard_pairs(11, rank, world))
    acc = torch.tensor([sum(i / 10 for i in mine), float(len(mine)), float(sum(mine)), float(len(mine)),
                        float(sum(1000 - i for i in mine))], dtype=torch.float64)
    parallel.barrier()
    acc = parallel.reduce_counts(acc)
    res = parallel.finalize_repeatability(acc)
    if rank == 0:
        torch.save(res, out)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_reduce_matches_single_process(tmp_path):
    out = str(tmp_path / 'res.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    want_rep = sum(i / 10 for i in range(11)) / 11
    assert res['pairs'] == 11
    assert abs(res['repeatability'] - want_rep) < 1e-12
    assert abs(res['rep_mean_err'] - sum(range(11)) / 11) < 1e-12
    assert abs(res['num_feat'] - sum(1000 - i for i in range(11)) / 11) < 1e-12
