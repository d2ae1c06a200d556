"""Drop-in for the brute-force half of the reference's ``utils/matcher.py``.

    brute_force_matcher   utils/matcher.py:206-234

The optical-flow matchers of that module (matcher.py:7-203) are a different algorithm and are out
of scope; asking for them raises.
"""
import math

import torch

from .. import ops
from ._dev import like, to_cuda


def sample_descriptors_at(desc_map: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """[b,c,h,w], [n,>=2] normalised -> [n,c] (matcher.py:221-226; batch item 0, no normalisation)."""
    d = to_cuda(desc_map)[0:1]
    p = to_cuda(pts)[None, :, :2].contiguous()
    return ops.sample_batched(d, p, None, normalize=False, coord_mode=0)[0]


def brute_force_matcher(pts0: torch.Tensor, pts1: torch.Tensor, desc_map_0: torch.Tensor, desc_map_1: torch.Tensor,
                        params=None):
    """(n,2+) / (m,2+) keypoints in [0,1] + (b,c,h,w) descriptor maps -> matched rows of pts0 / pts1.
    ``params``: {'metric': 'euclidean', 'max_distance': float, 'cross_check': bool} (matcher.py:228-230)."""
    metric = params['metric']
    if metric not in (None, 'euclidean'):
        raise ValueError(f"keypoint_bench_b200 brute_force_matcher supports metric='euclidean' only, got {metric!r}")
    if pts0.shape[0] == 0 or pts1.shape[0] == 0:
        raise ValueError('attempt to get argmin of an empty sequence')      # what numpy raises in the reference
    desc0 = sample_descriptors_at(desc_map_0, pts0)
    desc1 = sample_descriptors_at(desc_map_1, pts1)
    max_distance = params['max_distance']
    max_distance = math.inf if max_distance is None else float(max_distance)
    algo = int(params.get('algo', -1))       # -1: tcgen05 path when D <= 256, float64 SIMT otherwise
    pairs, _, count = ops.match_batched(desc0[None], desc1[None], None, None, max_distance, bool(params['cross_check']),
                                        algo=algo, want_dist=False)
    k = int(count[0].item())
    matches = pairs[0, :k].to(torch.int64)
    matches = like(matches, pts0)
    return pts0[matches[:, 0]], pts1[matches[:, 1]]


def optical_flow_tensor(*_args, **_kwargs):
    raise NotImplementedError('optical-flow matchers (utils/matcher.py:7-203) are outside the accelerated path')


optical_flow_cv = optical_flow_tensor
