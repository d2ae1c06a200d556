"""Multi-GPU plumbing: pairs are independent, so work shards by pair with no data-path collective;
the only exchange is one all-reduce of the float64 accumulator vector at the end of a run
(models/model_interface.py:124-137 takes means of per-pair values).  One process per GPU,
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank_world() -> tuple[int, int, int]:
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def init(backend: str | None = None) -> tuple[int, int]:
    """Join the process group described by RANK/WORLD_SIZE/MASTER_* (no-op for a single process)."""
    rank, local, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        kw = {'device_id': torch.device('cuda', local)} if backend == 'nccl' else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world


def shard_pairs(n_pairs: int, rank: int, world: int) -> range:
    """Pair datasets: pair i -> rank i mod world (results are order-independent means)."""
    return range(rank, n_pairs, world)


def shard_stream(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Frame streams (KITTI-like): contiguous chunks with a one-frame halo so that every consecutive
    pair (t-1,t) is owned by exactly one rank and every frame is detected once per rank.
    Returns [first_frame, last_frame) of this rank; pairs are (t-1,t) for t in (first, last)."""
    n_pairs = max(n_frames - 1, 0)
    lo = rank * n_pairs // world
    hi = (rank + 1) * n_pairs // world
    return lo, (hi + 1 if hi > lo else lo)


def reduce_counts(vec: torch.Tensor) -> torch.Tensor:
    """Sum the accumulator vector over ranks (the single collective of a run)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def finalize_repeatability(acc: torch.Tensor) -> dict:
    a = acc.tolist()
    return {'repeatability': a[0] / a[1] if a[1] else 0.0, 'rep_mean_err': a[2] / a[3] if a[3] else float('nan'),
            'num_feat': a[4] / a[1] if a[1] else 0.0, 'pairs': int(a[1])}


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
