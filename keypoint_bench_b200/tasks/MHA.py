"""Drop-in for the reference's ``tasks/MHA.py:11-72``: detect -> covisible -> match on the device,
``cv2.findHomography(RANSAC)`` on the host exactly as the reference calls it, corner error on the device."""
import numpy as np
import torch

from .. import ops
from ..utils._dev import as_int
from ..utils.extracter import detection
from ..utils.matcher import brute_force_matcher
from ..utils.projection import warp


def mha(idx, img_0, score_map_0, desc_map_0, img_1, score_map_1, desc_map_1, warp01, warp10, params):
    import cv2
    th = params['MHA_params']['th']
    mha_result = [0 for _ in th]
    kps0 = detection(score_map_0, params['extractor_params'])
    kps1 = detection(score_map_1, params['extractor_params'])
    kps0_cov, _, _, _ = warp(kps0, warp01)
    kps1_cov, _, _, _ = warp(kps1, warp10)
    if kps0_cov.shape[0] == 0 or kps1_cov.shape[0] == 0:
        return mha_result
    m_pts0, m_pts1 = brute_force_matcher(kps0_cov, kps1_cov, desc_map_0, desc_map_1,
                                         params['matcher_params']['brute_force_params'])
    h, w = as_int(warp01['height']), as_int(warp01['width'])
    scale = torch.tensor([w - 1, h - 1], dtype=torch.float32, device=m_pts0.device)
    p0 = (m_pts0[:, 0:2] * scale).cpu().numpy()
    p1 = (m_pts1[:, 0:2] * scale).cpu().numpy()
    H, _ = cv2.findHomography(p0, p1, cv2.RANSAC)         # MHA.py:45-47 (raises on < 4 matches, as upstream)
    if H is None:
        return mha_result
    real_H = torch.as_tensor(warp01['homography_matrix']).detach().cpu().numpy().astype(np.float64)
    dev = torch.device('cuda', torch.cuda.current_device())
    _, flags = ops.corner_error_batched(torch.as_tensor(H, dtype=torch.float64, device=dev)[None],
                                        torch.as_tensor(real_H, device=dev)[None], None, w, h,
                                        img_0.shape[2], img_0.shape[3], th)
    return [float(v) for v in flags[0].cpu().tolist()]
