"""Drop-in for the reference's ``tasks/MHA.py:11-72``: detect -> covisible -> match on the device,
``cv2.findHomography(RANSAC)`` on the host exactly as the reference calls it, corner error on the device."""
import numpy as np
import torch

from .. import ops
from ..utils._dev import as_int
from ..utils.extracter import detection
from ..utils.matcher import brute_force_matcher
from ..utils.projection import warp


def _covisible(score_map, warp_params, extractor_params):
    """Keypoints of one image that land inside the other one (detection + warp, MHA.py:31-36)."""
    inside, _, _, _ = warp(detection(score_map, extractor_params), warp_params)
    return inside


def mha(idx, img_0, score_map_0, desc_map_0, img_1, score_map_1, desc_map_1, warp01, warp10, params):
    """Same arguments and return value as the reference: one 0/1 flag per threshold of ``params['MHA_params']['th']``
    (all zero when either image has no covisible keypoint or RANSAC finds no homography)."""
    import cv2
    thresholds = params['MHA_params']['th']
    zeros = [0] * len(thresholds)
    cov0 = _covisible(score_map_0, warp01, params['extractor_params'])
    cov1 = _covisible(score_map_1, warp10, params['extractor_params'])
    if min(cov0.shape[0], cov1.shape[0]) == 0:
        return zeros
    matched0, matched1 = brute_force_matcher(cov0, cov1, desc_map_0, desc_map_1,
                                             params['matcher_params']['brute_force_params'])
    height, width = as_int(warp01['height']), as_int(warp01['width'])
    to_px = torch.tensor([width - 1, height - 1], dtype=torch.float32, device=matched0.device)
    src = (matched0[:, :2] * to_px).cpu().numpy()
    dst = (matched1[:, :2] * to_px).cpu().numpy()
    h_est, _ = cv2.findHomography(src, dst, cv2.RANSAC)    # host, where the reference calls it (MHA.py:45-47; raises on < 4 matches)
    if h_est is None:
        return zeros
    h_true = torch.as_tensor(warp01['homography_matrix']).detach().cpu().numpy().astype(np.float64)
    dev = torch.device('cuda', torch.cuda.current_device())
    # corner projection, rescaling to the network input size and thresholds: MHA.py:51-72, on the device in float64
    _, flags = ops.corner_error_batched(torch.as_tensor(h_est, dtype=torch.float64, device=dev)[None],
                                        torch.as_tensor(h_true, device=dev)[None], None, width, height,
                                        img_0.shape[2], img_0.shape[3], thresholds)
    return [float(v) for v in flags[0].cpu().tolist()]
