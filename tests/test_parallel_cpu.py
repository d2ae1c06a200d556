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


# ------------------------------------------------------------------------------------------------ accumulators

def _reference_style_means(per_pair):
    """What models/model_interface.py does with the per-pair results of tasks/repeatability.py: lists appended in
    test_step (:242-246), means in on_test_end (:124-133) -- np.mean of repeatability, NaN-filtered np.mean of
    mean_error, np.mean of num_feat."""
    import numpy as np
    rep = np.asarray([p['repeatability'] for p in per_pair], dtype=np.float64)
    err = np.asarray([p['mean_error'] for p in per_pair], dtype=np.float64)
    err = err[~np.isnan(err)]
    return float(np.mean(rep)), (float(np.mean(err)) if err.size else float('nan')), \
        float(np.mean([p['num_feat'] for p in per_pair]))


def _fake_results(seed, n):
    """Per-pair outcomes in the three shapes val_key_points produces (tasks/repeatability.py:61-92): a normal pair, a
    pair whose mutual matches are all beyond th (gt_num 0 -> mean_error NaN) and a pair with no covisible keypoints
    (the zeros dict: num_feat 0, repeatability 0, mean_error 0 -- which on_test_end does NOT filter)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    stats = np.zeros((n, 4)); num_feat = np.zeros(n, dtype=np.int32); empty = np.zeros(n, dtype=bool)
    per_pair = []
    for i in range(n):
        kind = i % 4
        if kind == 3:
            empty[i] = True
            per_pair.append({'num_feat': 0, 'repeatability': 0, 'mean_error': 0})
            continue
        nf = int(rng.integers(50, 1000))
        gt = 0 if kind == 2 else int(rng.integers(1, nf))
        s = float(rng.random() * 3 * gt)
        stats[i] = (gt, s, gt + 5, 0)
        num_feat[i] = nf
        per_pair.append({'num_feat': nf, 'repeatability': gt / nf, 'mean_error': s / gt if gt else float('nan')})
    res = {'stats': torch.from_numpy(stats), 'num_feat': torch.from_numpy(num_feat), 'empty': torch.from_numpy(empty)}
    return res, per_pair


def test_accumulate_repeatability_equals_the_reference_means():
    from keypoint_bench_b200 import pipeline
    res, per_pair = _fake_results(3, 23)
    fin = parallel.finalize_repeatability(pipeline.accumulate_repeatability(res))
    rep, err, nf = _reference_style_means(per_pair)
    assert abs(fin['repeatability'] - rep) < 1e-12 and abs(fin['rep_mean_err'] - err) < 1e-12
    assert abs(fin['num_feat'] - nf) < 1e-12 and fin['pairs'] == 23
    # every pair NaN -> the reference's np.mean of an empty array is NaN
    res, per_pair = _fake_results(4, 1)
    res['stats'][:, :2] = 0
    fin = parallel.finalize_repeatability(pipeline.accumulate_repeatability(res))
    assert fin['rep_mean_err'] != fin['rep_mean_err']


def _acc_worker(rank, world, port, out):
    from keypoint_bench_b200 import pipeline
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    parallel.init('gloo')
    res, _ = _fake_results(11, 30)
    mine = torch.tensor(list(parallel.shard_pairs(30, rank, world)))
    shard = {k: v[mine] for k, v in res.items()}
    acc = parallel.reduce_counts(pipeline.accumulate_repeatability(shard))
    # the stream / match accumulator: [sum of matches, pairs] over this rank's chunk of a 41-frame sequence
    lo, hi = parallel.shard_stream(41, rank, world)
    n_matches = torch.arange(40)[lo:hi - 1] * 3 + 100          # pair (t-1, t) has 100 + 3 (t-1) matches
    acc_m = parallel.reduce_counts(pipeline.accumulate_matches({'n_matches': n_matches}))
    if rank == 0:
        torch.save({'rep': parallel.finalize_repeatability(acc), 'matches': acc_m}, out)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_accumulators_equal_the_single_process_reference_means(tmp_path):
    out = str(tmp_path / 'acc.pt')
    mp.spawn(_acc_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    _, per_pair = _fake_results(11, 30)
    rep, err, nf = _reference_style_means(per_pair)
    assert abs(got['rep']['repeatability'] - rep) < 1e-12 and abs(got['rep']['rep_mean_err'] - err) < 1e-12
    assert abs(got['rep']['num_feat'] - nf) < 1e-12 and got['rep']['pairs'] == 30
    assert got['matches'].tolist() == [float(sum(100 + 3 * t for t in range(40))), 40.0]
