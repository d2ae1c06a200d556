"""Shared comparison rules of the GPU parity tests (defined beside the oracle, oracle/compare.py)."""
import numpy as np

from oracle import ref_ops
from oracle.compare import check_detection_rows, exact_pairs_or_near_tie, explain_repeat_pair_diffs  # noqa: F401


def oracle_dist_mutual(k0, k1, warp01, warp10):
    """dist_mutual of tasks/repeatability.py:69-73 (float32, 99999 on the index diagonal) from the oracle's warps."""
    k0c, k01c, _, _ = ref_ops.warp(k0, warp01)
    k1c, k10c, _, _ = ref_ops.warp(k1, warp10)
    d01 = ref_ops.keypoint_distance(k0c, k10c)
    d10 = ref_ops.keypoint_distance(k1c, k01c)
    dm = ((d01 + d10.T) / np.float32(2)).astype(np.float32)
    n = min(dm.shape)
    dm[np.arange(n), np.arange(n)] = np.float32(99999)
    return dm
