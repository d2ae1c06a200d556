"""Batched, device-resident composition of the hot path for a batch of image pairs.

This is what the reference's ``MInterface.test_step`` does per pair after the backbone
(models/model_interface.py:231-253 -> tasks/repeatability.py:95-122 / tasks/MHA.py:11-72), expressed
over a whole batch with no host synchronisation between stages:

    detection x2 -> warp x2 (covisibility) -> [descriptor sampling x2 -> mutual-NN matching]
                                           -> [val_key_points counting]

Images 0 and 1 of every pair are stacked along the batch (first P maps = image 0, last P = image 1)
so each stage is one launch for 2P maps.
"""
from __future__ import annotations

import torch

from . import ops
from .synth import FrameBatch, PairBatch, PathConfig  # noqa: F401  (re-exported: the batches are defined beside the inputs)


def extract_match(batch: PairBatch, cfg: PathConfig, algo: int = -1, covisible_only: bool = True, timer=None) -> dict:
    """detect -> (covisible) -> sample -> match for every pair.  Returns padded device tensors:
    kpts [2P,top_k,3] + n_kpts [2P]; kcov [2P,top_k,2] + n_cov [2P] (when covisible_only);
    matches [P,top_k,2] int32 (indices into the rows fed to the matcher) + n_matches [P].
    Rows beyond the per-map counts are unspecified (outputs are not zero-filled here)."""
    with ops.no_zero_fill():
        return _extract_match(batch, cfg, algo, covisible_only, timer)


def _extract_match(batch, cfg, algo, covisible_only, timer) -> dict:
    P = batch.pairs
    t = timer or (lambda name: None)
    t('detect')
    xyp, count, raster, path = ops.detect_batched(batch.score, cfg.extractor_params)
    out = {'kpts': xyp, 'n_kpts': count, 'raster': raster, 'path': path}
    pts, n_pts = xyp, count
    if covisible_only:                                        # tasks/MHA.py:33-34
        t('warp')
        kv, kw, ids, ids_out, nv = ops.warp_batched(xyp, count, batch.h33, batch.wh)
        out.update(kcov=kv, kwarp=kw, cov_ids=ids, n_cov=nv)
        pts, n_pts = kv, nv
    if batch.desc is not None:
        t('sample')
        d = ops.sample_batched(batch.desc, pts, n_pts)        # utils/matcher.py:221-226
        t('match')
        # the reference returns matched rows only (utils/matcher.py:227-233): no distances are requested.
        # (ops.sample_match_batched(fused=True) is the form in which the sampler writes the matcher's operand rows itself:
        # identical pairs, measured SLOWER on B200 -- 222 + 12 us against 157 + 53 us at cfg2, DESIGN.md section 5 -- so the
        # two calls stay the default)
        pairs, _, n_m = ops.match_batched(d[:P], d[P:], n_pts[:P], n_pts[P:], cfg.max_distance, cfg.cross_check,
                                          algo=algo, want_dist=False)
        out.update(desc=d, matches=pairs, n_matches=n_m)
    t(None)
    return out


def extract_match_stream(frames: FrameBatch, cfg: PathConfig, algo: int = -1) -> dict:
    """Stream form of the path: every frame is detected and sampled ONCE and matched against its predecessor.
    The reference's per-pair step (models/model_interface.py:217-228 keeps the previous frame's maps and calls
    ``detection`` on both maps of every pair) does the previous frame's extraction a second time; the results per
    pair are identical because extraction depends on the frame alone.  The stream tasks match ALL keypoints
    (no covisibility warp: tasks/FundamentalMatrix.py:53-57, tasks/visual_odometer.py:44-60).
    Returns kpts [F+1,top_k,3], n_kpts [F+1], desc [F+1,top_k,C], matches [F,top_k,2] (pair f = frames f, f+1),
    n_matches [F].  A rank's chunk carries a one-frame halo (parallel.shard_stream)."""
    with ops.no_zero_fill():
        xyp, count, raster, path = ops.detect_batched(frames.score, cfg.extractor_params)
        d = ops.sample_batched(frames.desc, xyp, count)
        pairs, _, n_m = ops.match_batched(d[:-1], d[1:], count[:-1], count[1:], cfg.max_distance, cfg.cross_check,
                                          algo=algo, want_dist=False)
    return {'kpts': xyp, 'n_kpts': count, 'raster': raster, 'path': path, 'desc': d, 'matches': pairs, 'n_matches': n_m}


def repeatability_counts(batch: PairBatch, cfg: PathConfig, th: float = 3.0, timer=None) -> dict:
    """detect x2 -> warp x2 -> val_key_points counting for every pair (tasks/repeatability.py:95-122).
    stats[P,4] float64 = (gt_num, sum of errors<=th, n mutual pairs, 0); num_feat[P] = min(n0,n1)."""
    P = batch.pairs
    t = timer or (lambda name: None)
    with ops.no_zero_fill():                                  # rows beyond the per-map counts are unspecified
        t('detect')
        xyp, count, raster, path = ops.detect_batched(batch.score, cfg.extractor_params)
        t('warp')
        kv, kw, ids, ids_out, nv = ops.warp_batched(xyp, count, batch.h33, batch.wh)
        t('repeat')
        stats, errors, _ = ops.repeat_batched(kv[:P], kw[:P], nv[:P], kv[P:], kw[P:], nv[P:], float(batch.resize),
                                              float(batch.resize), th, want_errors=True)
        t(None)
    num_feat = torch.minimum(count[:P], count[P:])
    empty = (nv[:P] == 0) | (nv[P:] == 0)                     # repeatability.py:61-67 -> zeros
    return {'stats': stats, 'errors': errors, 'num_feat': torch.where(empty, torch.zeros_like(num_feat), num_feat),
            'kpts': xyp, 'n_kpts': count, 'raster': raster, 'path': path, 'kcov': kv, 'kwarp': kw, 'n_cov': nv,
            'empty': empty}


def accumulate_repeatability(res: dict) -> torch.Tensor:
    """Per-batch contribution to the run accumulators of models/model_interface.py:124-133:
    float64 [sum repeatability_i, n_pairs, sum mean_error_i over non-NaN pairs, n_nonNaN, sum num_feat_i]."""
    st = res['stats']
    nf = res['num_feat'].to(torch.float64)
    gt = st[:, 0]
    rep = torch.where(nf > 0, gt / nf.clamp(min=1), torch.zeros_like(gt))
    has = gt > 0
    mean_err = torch.where(has, st[:, 1] / gt.clamp(min=1), torch.zeros_like(gt))
    # pairs with no covisible keypoints report mean_error 0 (not NaN) in the reference
    counted = has | res['empty']
    return torch.stack([rep.sum(), st.new_full((), float(st.shape[0])), mean_err.sum(), counted.to(torch.float64).sum(),
                        nf.sum()])


def accumulate_matches(res: dict) -> torch.Tensor:
    """float64 [sum matches, n_pairs] (the stream / AUC configs count matches per pair)."""
    n = res['n_matches'].to(torch.float64)
    return torch.stack([n.sum(), n.new_full((), float(n.numel()))])


def run_task(inputs, cfg: PathConfig, algo: int = -1, timer=None, th: float = 3.0):
    """One step of the path as the task named by ``cfg.task`` runs it in the reference, -> (results, float64
    accumulator contribution):

    'repeatability'  detect x2 -> warp x2 -> val_key_points counting       (tasks/repeatability.py:95-122)
    'mha'            detect x2 -> warp x2 -> sample + match the COVISIBLE keypoints (tasks/MHA.py:29-39; the host
                     cv2 RANSAC that follows is out of scope)
    'match'          detect x2 -> sample + match ALL keypoints, no warp     (tasks/AUC.py:115-120)
    'stream'         ``inputs`` is a FrameBatch: every frame extracted once and matched with its predecessor, all
                     keypoints (tasks/FundamentalMatrix.py:53-57 on models/model_interface.py:217-228's frame pairs)."""
    if cfg.task == 'repeatability':
        res = repeatability_counts(inputs, cfg, th, timer)
        return res, accumulate_repeatability(res)
    if cfg.task == 'stream':
        res = extract_match_stream(inputs, cfg, algo)
    elif cfg.task in ('mha', 'match'):
        res = extract_match(inputs, cfg, algo=algo, covisible_only=(cfg.task == 'mha'), timer=timer)
    else:
        raise ValueError(f'unknown task {cfg.task!r}')
    return res, accumulate_matches(res)


class GraphedStep:
    """One pipeline step captured in a CUDA graph: the ~15 kernel launches of a step cost more host time than
    device time when issued one by one from Python, so steady-state callers replay the captured launch
    sequence instead.  ``fn`` must read its inputs from fixed device tensors (e.g. a PairBatch whose tensors
    are refilled in place) and return a tuple/dict of device tensors, which keep their addresses."""

    def __init__(self, fn, warmup: int = 2):
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out


class StepsInFlight:
    """Keeps ``len(fns)`` steps in flight: one captured CUDA graph per slot, each replayed on its own stream over
    its own input batch and output buffers.  A step is a chain of dependent kernels of which several are
    latency-bound (one CTA per map in the sparse NMS resolve, the matcher's resolve / rescan / gate tails), so a
    other steps' bandwidth- and tensor-bound kernels fill the SMs those leave idle (cfg2: +10 % pairs/s with two
    slots, +14 % with three, measured).  ``launch(i)`` replays slot ``i % depth`` and returns that slot's outputs, valid
    until the slot is launched again; work queued with ``on_slot`` runs on the slot's stream after it."""

    def __init__(self, fns):
        self.slots = [GraphedStep(fn) for fn in fns]
        self.streams = [torch.cuda.Stream() for _ in fns]

    @property
    def depth(self) -> int:
        return len(self.slots)

    def fork(self) -> None:
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)

    def join(self) -> None:
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)

    def stream(self, i: int):
        return self.streams[i % len(self.streams)]

    def launch(self, i: int, before=None, after=None):
        k = i % len(self.slots)
        with torch.cuda.stream(self.streams[k]):
            if before is not None:
                before(k)
            out = self.slots[k]()
            if after is not None:
                after(k, out)
        return out
