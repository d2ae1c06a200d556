"""Drop-in for the LightGlue-style extraction helpers of the reference's ``models/lightglue.py``
(SURVEY.md section 8(f) rank 3): the second NMS / top-k flavour of the code base.

    sample_descriptors   models/lightglue.py:24-41
    simple_nms           models/lightglue.py:904-920
    top_k_kps            models/lightglue.py:923-927
    extract_from_maps    models/lightglue.py:929-979 minus the network call (``scores, descriptors = net(image)``)
    extract              the same with the reference's ``(net, image, s)`` signature

The LightGlue matcher network itself is a backbone and stays out of scope.
"""
import torch

from .. import ops
from ._dev import like, to_cuda


def sample_descriptors(keypoints: torch.Tensor, descriptors: torch.Tensor, s: int = 8) -> torch.Tensor:
    """[b,n,2] pixel keypoints (x, y), [b,c,h,w] -> [b,c,n], L2-normalised over c (lightglue.py:24-41)."""
    d = to_cuda(descriptors)
    k = to_cuda(keypoints).float().contiguous()
    out = ops.sample_batched(d, k, None, normalize=True, coord_mode=1, s=int(s))        # [b,n,c]
    return like(out.transpose(1, 2), descriptors)


def simple_nms(scores: torch.Tensor, nms_radius: int) -> torch.Tensor:
    """Max-pool NMS with two recovery passes (lightglue.py:904-920); any leading dims, same shape out."""
    assert nms_radius >= 0
    return like(ops.simple_nms_batched(to_cuda(scores), int(nms_radius)), scores)


def top_k_kps(keypoints: torch.Tensor, scores: torch.Tensor, k: int):
    """lightglue.py:923-927 (``torch.topk`` is itself the reference's device code; ties in its order)."""
    if k >= len(keypoints):
        return keypoints, scores
    scores, indices = torch.topk(scores, k, dim=0, sorted=True)
    return keypoints[indices], scores


def extract_batched(scores: torch.Tensor, descriptors: torch.Tensor, s: int, detection_threshold: float = 0.0,
                    pad: int = 4, nms_radius: int = 5, max_num_kps: int = 1000):
    """Every map of the batch, no host synchronisation: -> keypoints [B,K,2] (x, y pixels), keypoint_scores [B,K],
    descriptors [B,K,C] (unit norm), count [B]; rows >= count[b] are zero.  Top-k ties are broken by raster index."""
    if detection_threshold < 0:
        raise ValueError('detection_threshold must be >= 0 (the border is marked with -1, lightglue.py:942-945)')
    sc = to_cuda(scores)
    w = sc.shape[-1]
    nms = ops.simple_nms_batched(sc, int(nms_radius))
    # border rows/cols -> excluded, `> threshold`, top-k by score when more than k remain, raster order otherwise
    xyp, count, raster, _ = ops.select_batched(nms, int(pad), float(detection_threshold), 0.0, int(max_num_kps))
    live = torch.arange(raster.shape[1], device=raster.device)[None, :] < count[:, None]
    # (rows beyond the count may be uninitialised under ops.no_zero_fill: select, never multiply)
    kps = torch.where(live[..., None], torch.stack([raster % w, raster // w], dim=-1).float(), 0.0)     # (h, w) -> (x, y), lightglue.py:970
    val = torch.where(live, xyp[..., 2], 0.0)
    desc = ops.sample_batched(to_cuda(descriptors), kps, count, normalize=True, coord_mode=1, s=int(s))
    return kps, val, desc, count


def extract_from_maps(scores: torch.Tensor, descriptors: torch.Tensor, s: int, detection_threshold: float = 0.0,
                      pad: int = 4, nms_radius: int = 5, max_num_kps: int = 1000) -> dict:
    """What the reference's ``extract`` returns for the network outputs ``scores`` [b,1,h,w] and ``descriptors``
    [b,c,h',w']: like the reference (``scores = scores[0]``, lightglue.py:938) only batch item 0 is used.
    -> {'keypoints': [1,n,2], 'keypoint_scores': [1,n], 'descriptors': [1,n,c]}."""
    kps, val, desc, count = extract_batched(scores[0:1], descriptors[0:1], s, detection_threshold, pad, nms_radius,
                                            max_num_kps)
    n = int(count[0].item())
    return {'keypoints': like(kps[:, :n], scores), 'keypoint_scores': like(val[:, :n], scores),
            'descriptors': like(desc[:, :n].contiguous(), scores)}


def extract(net, image: torch.Tensor, s: int) -> dict:
    """lightglue.py:929-979 with the reference's constants (threshold 0, pad 4, radius 5, 1000 keypoints)."""
    scores, descriptors = net(image)
    return extract_from_maps(scores, descriptors, s)
