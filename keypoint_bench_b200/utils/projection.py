"""Drop-in for the keypoint-warping half of the reference's ``utils/projection.py``.

    to_homogeneous     utils/projection.py:128-134
    warp_homography    utils/projection.py:137-167
    warp               utils/projection.py:185-192
    warp_se3           utils/projection.py:194-267 (with interpolate_depth :270-372)

The numpy depth-map utilities of that file (warp_depth, unproject_depth, ... :7-125, used by the MegaDepth
loader at training time) are dataset code and out of scope.
"""
import torch

from .. import ops
from ._dev import as_int, like, to_cuda


def to_homogeneous(kpts: torch.Tensor) -> torch.Tensor:
    """Nx2 -> Nx3 (projection.py:128-134)."""
    return torch.cat((kpts, kpts.new_ones([kpts.shape[0], 1])), dim=1)


def warp_homography(kpts0: torch.Tensor, params: dict):
    """Nx2 normalised keypoints -> (valid kpts, warped valid kpts, ids, ids_out) (projection.py:137-167)."""
    w, h = as_int(params['width']), as_int(params['height'])
    k = to_cuda(kpts0)
    if k.shape[0] == 0:
        e = kpts0.new_zeros((0, 2))
        i = torch.zeros(0, dtype=torch.int64, device=kpts0.device)
        return e, e.clone(), i, i.clone()
    hm = to_cuda(torch.as_tensor(params['homography_matrix'])).to(torch.float32).reshape(1, 9)
    wh = torch.tensor([[float(w), float(h)]], dtype=torch.float32, device=k.device)
    kv, kw, ids, ids_out, nv = ops.warp_batched(k[None, :, :2].contiguous(), None, hm, wh)
    a = int(nv[0].item())
    n = k.shape[0]
    return (like(kv[0, :a], kpts0), like(kw[0, :a], kpts0), like(ids[0, :a].to(torch.int64), kpts0),
            like(ids_out[0, :n - a].to(torch.int64), kpts0))


def warp(kpts0: torch.Tensor, params: dict):
    """Dispatch on params['mode'] (projection.py:185-192)."""
    mode = params['mode']
    if mode == 'homo':
        return warp_homography(kpts0[:, 0:2], params)
    if mode == 'se3':
        return warp_se3(kpts0[:, 0:2], params)
    raise ValueError('unknown mode!')


def warp_se3(kpts0: torch.Tensor, params: dict):
    """Depth-based covisibility (projection.py:194-267): -> (valid kpts0, their projections into view 1,
    ids of the valid ones, ids that are surely unmatched: projected outside view 1 or occluded)."""
    k = to_cuda(kpts0)
    if k.shape[0] == 0:
        e = kpts0.new_zeros((0, 2))
        i = torch.zeros(0, dtype=torch.int64, device=kpts0.device)
        return e, e.clone(), i, i.clone()
    t = lambda name: to_cuda(torch.as_tensor(params[name]))[None]      # noqa: E731
    kv, kw, ids, ids_out, nv, no = ops.warp_se3_batched(k[None, :, :2].contiguous(), None, t('depth0'), t('depth1'),
                                                        t('intrinsics0'), t('intrinsics1'), t('pose01'), t('bbox0'),
                                                        t('bbox1'))
    a, o = int(nv[0].item()), int(no[0].item())
    return (like(kv[0, :a], kpts0), like(kw[0, :a], kpts0), like(ids[0, :a].to(torch.int64), kpts0),
            like(ids_out[0, :o].to(torch.int64), kpts0))
