"""Comparison rules of the parity checks (canonical comparison rules).

TEST INFRASTRUCTURE: like everything under oracle/, this is only imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / parity spot check -- never by the product package."""
import numpy as np
from scipy.spatial.distance import cdist


def exact_pairs_or_near_tie(got, d0, d1, maxd, cc, dist=None):
    """Match pairs must equal the float64 restatement of skimage.match_descriptors (utils/matcher.py:227-230) except
    rows / columns whose best and second-best distances are within 1e-5 relative, or pairs whose distance is within
    1e-5 relative of ``max_distance`` (north_star).  ``dist`` = a precomputed float64 cdist(d0, d1).
    Returns the number of tolerated differences; raises AssertionError on any other difference."""
    D = cdist(d0, d1) if dist is None else dist
    i0 = np.arange(D.shape[0])
    i1 = np.argmin(D, axis=1)
    if cc:
        keep = i0 == np.argmin(D, axis=0)[i1]
        i0, i1 = i0[keep], i1[keep]
    if maxd < np.inf:
        keep = D[i0, i1] < maxd
        i0, i1 = i0[keep], i1[keep]
    want = np.column_stack((i0, i1)).astype(np.int64)
    if np.array_equal(got, want):
        return 0
    s = np.partition(D, 1, axis=1)[:, :2] if D.shape[1] > 1 else np.concatenate([D, D], axis=1)
    t = np.partition(D, 1, axis=0)[:2] if D.shape[0] > 1 else np.concatenate([D, D], axis=0)
    amb_rows = set(np.flatnonzero((s[:, 1] - s[:, 0]) <= 1e-5 * s[:, 1]).tolist())
    amb_cols = set(np.flatnonzero((t[1] - t[0]) <= 1e-5 * t[1]).tolist())
    diff = set(map(tuple, got.tolist())) ^ set(map(tuple, want.tolist()))
    for i, j in diff:
        assert i in amb_rows or j in amb_cols or abs(D[i, j] - maxd) <= 1e-5 * maxd, (i, j, D[i, j])
    return len(diff)


def check_detection_rows(pts, want, raster=None, want_raster=None):
    """Keypoint rows: verbatim (including order) wherever the score is unique, same multiset of scores otherwise
    (the reference's argsort leaves the order of equal scores unspecified, utils/extracter.py:217-218); raster
    indices in canonical order (score desc, raster asc)."""
    assert pts.shape == want.shape, (pts.shape, want.shape)
    uniq, cnt = np.unique(want[:, 2], return_counts=True)
    single = uniq[cnt == 1]
    tf = np.isin(want[:, 2], single) & np.isin(pts[:, 2], single)
    assert np.array_equal(pts[tf], want[tf])
    assert np.array_equal(np.sort(pts[:, 2]), np.sort(want[:, 2]))
    if want_raster is not None:
        assert np.array_equal(raster, want_raster)


def explain_repeat_pair_diffs(got_pairs, want_pairs, dm):
    """tasks/repeatability.py:18-32 compares v = (-dm) - min(-dm) with ==: with the 99999 diagonal present v is
    quantised to 2^-7, so a 1-ulp difference in a warped coordinate (the reference's einsum vs three rounded
    products) can move an entry across a bucket edge.  Every pair in the symmetric difference must be such an
    edge case: its quantised value within one bucket of both its row and its column maximum."""
    neg = (-dm).astype(np.float32)
    v = (neg - neg.min()).astype(np.float32)
    row, col = v.max(axis=1), v.max(axis=0)
    bucket = np.float32(2.0 ** -7)
    diff = set(map(tuple, got_pairs)) ^ set(map(tuple, want_pairs))
    for i, j in diff:
        assert row[i] - v[i, j] <= bucket and col[j] - v[i, j] <= bucket, (i, j, float(v[i, j]), float(row[i]), float(col[j]))
    return len(diff)
