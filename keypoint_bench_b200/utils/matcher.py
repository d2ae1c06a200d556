"""Drop-in for the reference's ``utils/matcher.py``.

    brute_force_matcher   utils/matcher.py:206-234
    OpticalFlow           utils/matcher.py:7-142      (tensor Lucas-Kanade tracker)
    optical_flow_tensor   utils/matcher.py:188-203

``optical_flow_cv`` (matcher.py:145-185) is a host call into OpenCV's ``calcOpticalFlowPyrLK`` on uint8 images;
it is not part of the accelerated path and raises.
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


class OpticalFlow(object):
    """Pyramidal Gauss-Newton patch tracker (matcher.py:7-142).  Same constructor dict and call signature; the
    whole coarse-to-fine schedule runs in one kernel per call instead of unfolding C*win^2-channel maps."""

    def __init__(self, params=None):
        if params is None:                                       # matcher.py:9-16
            params = {'distance': 3, 'win_size': 3, 'levels': 1, 'interation': 40, 'gray': False}
        self.distance = params['distance']
        self.win_size = params['win_size']
        self.levels = params['levels']
        self.interation = params['interation']
        self.gray = params['gray']

    def start_points(self, pts2_px: torch.Tensor, h: int, w: int) -> torch.Tensor:
        """matcher.py:54-61: a random unit offset of length ``distance``, clamped 10 px inside the image."""
        angle = torch.randn(pts2_px[0, :, 0].shape, device=pts2_px.device) * 6.28
        start = pts2_px + torch.stack([torch.cos(angle), torch.sin(angle)], dim=1) * self.distance
        start[0, :, 0] = torch.clamp(start[0, :, 0], min=10, max=w - 10)
        start[0, :, 1] = torch.clamp(start[0, :, 1], min=10, max=h - 10)
        return start

    def __call__(self, img1, img2, pts1, pts2, start=None):
        i1, i2 = to_cuda(img1), to_cuda(img2)
        n, c, h, w = i1.shape
        if n != 1 or c != (1 if self.gray else 3):
            # the reference's Sobel conv weights are [1,1,3,3] (gray) or [3,3,3,3] and its grid has batch 1
            raise RuntimeError(f'OpticalFlow expects one {"1" if self.gray else "3"}-channel image, got {tuple(i1.shape)}')
        scale = torch.tensor([w - 1, h - 1], dtype=torch.float32, device=i1.device)
        p1 = to_cuda(pts1).float().unsqueeze(0) * scale          # matcher.py:51-52
        p2 = to_cuda(pts2).float().unsqueeze(0) * scale
        if start is None:
            start = self.start_points(p2, h, w)
        out = ops.lk_track_batched(i1, i2, p1, to_cuda(start).float(), None, int(self.win_size), int(self.levels),
                                   int(self.interation))
        error = torch.clamp(torch.norm(out - p2, dim=2), max=8)   # matcher.py:88-90
        return like(out, pts1), like(error, pts1)


def optical_flow_tensor(pts0: torch.Tensor, pts1: torch.Tensor, img0: torch.Tensor, img1: torch.Tensor, params=None):
    """(n,2) keypoints in [0,1] of both images + (1,c,h,w) images -> tracked points [1,n,2] in PIXELS
    (matcher.py:188-203; the reference does not re-normalise)."""
    pts1_, _error = OpticalFlow(params)(img0, img1, pts0, pts1)
    return pts1_


def optical_flow_cv(*_args, **_kwargs):
    raise NotImplementedError('optical_flow_cv (utils/matcher.py:145-185) is a host OpenCV call outside the '
                              'accelerated path')
