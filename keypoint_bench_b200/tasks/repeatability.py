"""Drop-in for the evaluation core of the reference's ``tasks/repeatability.py``.

    compute_keypoints_distance / mutual_argmin   tasks/repeatability.py:9-51   (fused on the device)
    val_key_points                               tasks/repeatability.py:54-92
    repeatability                                tasks/repeatability.py:95-122 (without the PNG dumps)
"""
import numpy as np
import torch

from .. import ops
from ..utils._dev import as_int, to_cuda
from ..utils.extracter import detection
from ..utils.projection import warp


def val_key_points(kps0, kps1, warp01, warp10, th: int = 3, return_pairs: bool = False):
    """Same dict as the reference: num_feat, repeatability, mean_error, errors (repeatability.py:54-92)."""
    num_feat = min(kps0.shape[0], kps1.shape[0])
    kps0_cov, kps01_cov, _, _ = warp(kps0, warp01)
    kps1_cov, kps10_cov, _, _ = warp(kps1, warp10)
    if kps0_cov.shape[0] == 0 or kps1_cov.shape[0] == 0:
        return {'num_feat': 0, 'repeatability': 0, 'mean_error': 0, 'errors': None}
    if 'resize' in warp01:                                   # repeatability.py:76-81
        s01, s10 = as_int(warp01['resize']), as_int(warp10['resize'])
    else:
        s01, s10 = as_int(warp01['width']), as_int(warp10['width'])
    a, b = kps0_cov.shape[0], kps1_cov.shape[0]
    stats, errors, pairs = ops.repeat_batched(to_cuda(kps0_cov)[None], to_cuda(kps01_cov)[None], None,
                                              to_cuda(kps1_cov)[None], to_cuda(kps10_cov)[None], None,
                                              float(s01), float(s10), float(th), True,
                                              pair_cap=(a + b) * 4 if return_pairs else 0)
    st = stats[0].cpu().numpy()
    gt_num = int(st[0])
    mean_error = np.float32(st[1] / st[0]) if gt_num else np.float32('nan')    # mean of an empty array
    out = {'num_feat': num_feat,
           'repeatability': torch.tensor(gt_num) / num_feat,
           'mean_error': mean_error,
           'errors': errors[0].to(kps0.device)}
    if return_pairs:
        n_pairs = int(st[2])
        out['pairs'] = pairs[0, :n_pairs].cpu().numpy()
        out['gt_num'] = gt_num
    return out


def repeatability(idx, img_0, score_map_0, img_1, score_map_1, warp01, warp10, params):
    """detection x2 + val_key_points (repeatability.py:95-122); the PNG overlays are not produced."""
    kps0 = detection(score_map_0, params['extractor_params'])
    kps1 = detection(score_map_1, params['extractor_params'])
    return val_key_points(kps0, kps1, warp01, warp10, th=params['repeatability_params']['th'])
