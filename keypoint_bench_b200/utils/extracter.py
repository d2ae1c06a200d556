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
    xyp, count, _, _ = ops.detect_batched(s, params)
    n = int(count[0].item())
    return like(xyp[0, :n], score_map)
