"""Drop-in for the reference's ``utils/extracter.py`` (same names, arguments and results).

    fast_nms                          utils/extracter.py:6-100
    prob_map_to_positions_with_prob   utils/extracter.py:129-161
    remove_border_points              utils/extracter.py:164-190
    detection                         utils/extracter.py:193-221

All arithmetic runs in the sm_100a kernels behind the C ABI (include/kb_b200.h); these wrappers
only translate arguments, size the outputs (the one host sync the variable-length return value
forces) and keep the reference's conventions (batch item 0, in-place border zeroing, raster /
sorted output order).  Batched, sync-free siblings: ``keypoint_bench_b200.ops``.
"""
import torch

from .. import ops
from ._dev import like, to_cuda


def fast_nms(image_probs: torch.Tensor, nms_dist: int = 4, max_iter: int = -1, min_value: float = 0.0) -> torch.Tensor:
    """BxCxHxW -> BxCxHxW with suppressed pixels set to ``min_value`` (extracter.py:6-100)."""
    if nms_dist == 0:
        return image_probs                      # extracter.py:40-41: the input object itself
    out = ops.fast_nms_batched(to_cuda(image_probs), nms_dist, max_iter, min_value)
    return like(out, image_probs).to(image_probs.dtype)


def remove_border_points(image_nms: torch.Tensor, border_dist: int = 4) -> torch.Tensor:
    """In-place zeroing of ``border_dist`` rows / columns on every side (extracter.py:164-190).
    Inside ``detection`` this predicate is fused into the selection kernel instead."""
    if border_dist > 0:
        image_nms[..., :, :border_dist] = 0.0
        image_nms[..., :, -border_dist:] = 0.0
        image_nms[..., :border_dist, :] = 0.0
        image_nms[..., -border_dist:, :] = 0.0
    return image_nms


def prob_map_to_positions_with_prob(prob_map: torch.Tensor, threshold: float = 0.0) -> torch.Tensor:
    """Nx1xHxW -> K x 3 rows (x, y, p) of batch item 0, raster order (extracter.py:129-161)."""
    m = to_cuda(prob_map).squeeze(dim=1)[0:1]
    xyp, count, _, _ = ops.select_batched(m, border_dist=0, threshold=threshold, min_score=0.0, top_k=0)
    k = int(count[0].item())
    return like(xyp[0, :k], prob_map)


def detection(score_map: torch.Tensor, params: dict = None) -> torch.Tensor:
    """Bx1xHxW -> N x 3 (x, y, prob) of batch item 0 (extracter.py:193-221)."""
    s = to_cuda(score_map.detach())
    if params is not None and params['nms_dist'] == 0 and params['border_dist'] > 0:
        # extracter.py:214-215: with nms_dist == 0 the border zeroing lands on the caller's tensor
        remove_border_points(score_map, params['border_dist'])
    if s.dim() == 4 and s.shape[0] > 1:
        # Only batch item 0 is returned (extracter.py:161), but fast_nms runs on the whole batch with ONE stopping
        # rule (the maxima count summed over the batch, extracter.py:73-78).  On non-negative maps every item reaches
        # its own fixed point whatever the others do, so item 0 alone gives the same rows; with negative scores the
        # stopping round is observable and the batch is processed jointly by the round-faithful kernel.
        nms_dist = 4 if params is None else params['nms_dist']
        if nms_dist > 0 and bool((s < 0).any()):
            border, thr, top_k, min_score = (8, 0.0, 300, 0.0) if params is None else (
                params['border_dist'], params['threshold'], params['top_k'], params['min_score'])
            nms = ops.fast_nms_batched(s, nms_dist)
            if top_k <= 0:
                return like(torch.zeros(0, 3, dtype=torch.float32, device=s.device), score_map)
            cap = int(top_k)
            if top_k > ops.SORT_CAP:                          # see ops.detect_batched: a top_k that can never bind
                cap = ops.nms_keep_bound(s.shape[-2], s.shape[-1], int(nms_dist)) if thr >= 0 else s.shape[-2] * s.shape[-1]
                if top_k < cap:
                    raise ops._lib.KbError(f'top_k = {top_k}: sorting more than {ops.SORT_CAP} survivors is not implemented')
                top_k = 0
            xyp, count, _, _ = ops.select_batched(nms[0:1], int(border), float(thr), float(min_score), int(top_k), cap=cap)
            return like(xyp[0, :int(count[0].item())], score_map)
        s = s[0:1]
    xyp, count, _, _ = ops.detect_batched(s, params)
    n = int(count[0].item())
    return like(xyp[0, :n], score_map)
