"""CPU restatement of the reference's post-network hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference file:line it follows (paths relative to the
upstream repo root).  Nothing in here is imported by the product package; see
``oracle/__init__.py`` for the rules and for how these restatements are pinned
against the reference's own outputs.

Two flavours exist for the NMS because the reference's algorithm is both the
semantics to match and the CPU cost to report:

* ``nms_rounds_im2col``   -- the same im2col / argmax / col2im rounds the reference
  executes (this is what ``bench.py`` times as the CPU baseline);
* ``nms_rounds_separable`` -- the same rounds, restated with separable running
  maxima so the test-suite finishes in seconds;
* ``nms_greedy``          -- the closed form the rounds converge to on
  non-negative maps (greedy NMS in the order score desc, raster asc).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from scipy.spatial.distance import cdist

# --------------------------------------------------------------------------------------
# Stage 1: score-map NMS   (utils/extracter.py:6-100)
# --------------------------------------------------------------------------------------


def nms_rounds_im2col(score: torch.Tensor, nms_dist: int = 4, max_iter: int = -1,
                      min_value: float = 0.0) -> torch.Tensor:
    """Round-for-round restatement of ``fast_nms`` (utils/extracter.py:6-100).

    Per round: every pixel's (2r+1)^2 zero-padded neighbourhood is materialised
    (extracter.py:54-66); a pixel is a local maximum iff the FIRST maximal entry
    of that neighbourhood is the centre (extracter.py:69-70); the loop stops when
    the number of maxima did not change (extracter.py:73-78, tested BEFORE
    suppressing); otherwise every pixel with a different maximum within
    Chebyshev distance r is overwritten with ``min_value`` (extracter.py:81-96).
    ``nms_dist == 0`` hands the input back untouched (extracter.py:40-41).
    """
    if nms_dist == 0:
        return score
    r = int(nms_dist)
    k = 2 * r + 1
    centre = (k * k) // 2
    b, _, h, w = score.shape
    seen = None
    rounds = 0
    while rounds != max_iter:
        cols = F.unfold(score, kernel_size=k, padding=r).reshape(b, k * k, h, w)
        is_max = cols.argmax(dim=1, keepdim=True) == centre
        n_max = int(is_max.sum())
        if seen is not None and n_max == seen:
            break
        seen = n_max
        spread = is_max.to(score.dtype).expand(-1, k * k, -1, -1).reshape(b, k * k, h * w).clone()
        spread[:, centre] = 0.0
        hit = F.fold(spread, output_size=(h, w), kernel_size=k, padding=r)
        score = score.masked_fill(hit > 0.0, min_value)
        rounds += 1
    return score


def _window_sides(v: np.ndarray, r: int):
    """For each pixel: max over the window entries that precede / follow the centre in
    raster order of the (2r+1)^2 zero-padded window (the argmax tie rule of
    extracter.py:69 -- first maximal index wins)."""
    h, w = v.shape[-2:]
    lead = v.shape[:-2]
    p = np.zeros(lead + (h + 2 * r, w + 2 * r), dtype=v.dtype)
    p[..., r:r + h, r:r + w] = v
    # horizontal full-width (2r+1) max of every padded row, at image columns
    hfull = p[..., :, 0:w].copy()
    for d in range(1, 2 * r + 1):
        np.maximum(hfull, p[..., :, d:d + w], out=hfull)
    above = hfull[..., 0:h, :].copy()            # rows y-r .. y-1 (padded rows 0..r-1 offset)
    for d in range(1, r):
        np.maximum(above, hfull[..., d:d + h, :], out=above)
    below = hfull[..., r + 1:r + 1 + h, :].copy()  # rows y+1 .. y+r
    for d in range(1, r):
        np.maximum(below, hfull[..., r + 1 + d:r + 1 + d + h, :], out=below)
    mid = p[..., r:r + h, :]
    left = mid[..., :, 0:w].copy()                # cols x-r .. x-1
    for d in range(1, r):
        np.maximum(left, mid[..., :, d:d + w], out=left)
    right = mid[..., :, r + 1:r + 1 + w].copy()   # cols x+1 .. x+r
    for d in range(1, r):
        np.maximum(right, mid[..., :, r + 1 + d:r + 1 + d + w], out=right)
    return np.maximum(above, left), np.maximum(below, right)


def _box_count(mask: np.ndarray, r: int) -> np.ndarray:
    h, w = mask.shape[-2:]
    lead = mask.shape[:-2]
    p = np.zeros(lead + (h + 2 * r, w + 2 * r), dtype=np.int32)
    p[..., r:r + h, r:r + w] = mask
    rows = np.zeros(lead + (h + 2 * r, w), dtype=np.int32)
    for d in range(2 * r + 1):
        rows += p[..., :, d:d + w]
    out = np.zeros(lead + (h, w), dtype=np.int32)
    for d in range(2 * r + 1):
        out += rows[..., d:d + h, :]
    return out


def nms_rounds_separable(score: np.ndarray, nms_dist: int = 4, max_iter: int = -1,
                         min_value: float = 0.0, return_rounds: bool = False):
    """Same rounds as ``nms_rounds_im2col`` (utils/extracter.py:49-98) without the
    169x im2col blow-up: centre is the first maximal entry of its window iff it is
    strictly greater than every earlier entry and >= every later entry
    (extracter.py:69-70); suppression = "another maximum within Chebyshev r"
    (extracter.py:81-96).  ``score``: [..., H, W] float32 (leading dims = batch,
    counted jointly like extracter.py:73)."""
    v = np.array(score, dtype=np.float32, copy=True)
    if nms_dist == 0:
        return (v, 0) if return_rounds else v
    r = int(nms_dist)
    seen = None
    rounds = 0
    while rounds != max_iter:
        before, after = _window_sides(v, r)
        is_max = (v > before) & (v >= after)
        n_max = int(is_max.sum())
        if seen is not None and n_max == seen:
            break
        seen = n_max
        others = _box_count(is_max, r) - is_max.astype(np.int32)
        v[others > 0] = np.float32(min_value)
        rounds += 1
    return (v, rounds) if return_rounds else v


def nms_greedy(score: np.ndarray, nms_dist: int) -> np.ndarray:
    """Closed form of the fixed point of utils/extracter.py:49-98 for maps whose
    entries are all >= 0: greedy NMS over the total order (score desc, raster asc),
    Chebyshev radius ``nms_dist``; suppressed pixels read 0.  [H, W] only."""
    v = np.asarray(score, dtype=np.float32)
    h, w = v.shape
    r = int(nms_dist)
    if r == 0:
        return v.copy()
    flat = v.ravel()
    cand = np.flatnonzero(flat > 0)
    order = cand[np.lexsort((cand, -flat[cand].astype(np.float64)))]
    dead = np.zeros((h, w), dtype=bool)
    out = np.zeros_like(v)
    for idx in order.tolist():
        y, x = divmod(idx, w)
        if dead[y, x]:
            continue
        out[y, x] = v[y, x]
        dead[max(0, y - r):y + r + 1, max(0, x - r):x + r + 1] = True
    return out


# --------------------------------------------------------------------------------------
# Stage 1b: border removal, positions, detection  (utils/extracter.py:129-221)
# --------------------------------------------------------------------------------------


def clear_border(nms_map: np.ndarray, border_dist: int = 4) -> np.ndarray:
    """utils/extracter.py:164-190 -- zero ``border_dist`` rows/columns on each side
    (the reference does it in place; the restatement returns a copy)."""
    out = np.array(nms_map, copy=True)
    b = int(border_dist)
    if b > 0:
        out[..., :, :b] = 0.0
        out[..., :, -b:] = 0.0
        out[..., :b, :] = 0.0
        out[..., -b:, :] = 0.0
    return out


def positions_with_prob(prob_map: np.ndarray, threshold: float = 0.0):
    """utils/extracter.py:129-161 -- raster-order ``prob > threshold`` of batch item 0,
    returned as rows (x, y, p) with x=(col+0.5)/W, y=(row+0.5)/H in float32 (one add,
    one IEEE divide: extracter.py:149,158).  Also returns the raster index of each row."""
    m = np.asarray(prob_map, dtype=np.float32)
    while m.ndim > 2:
        m = m[0]
    h, w = m.shape
    rows, cols = np.nonzero(m > np.float32(threshold))
    x = (cols.astype(np.float32) + np.float32(0.5)) / np.float32(w)
    y = (rows.astype(np.float32) + np.float32(0.5)) / np.float32(h)
    p = m[rows, cols]
    return np.stack([x, y, p], axis=1).astype(np.float32).reshape(-1, 3), (rows * w + cols).astype(np.int64)


def canonical_order(p: np.ndarray, raster: np.ndarray) -> np.ndarray:
    """Canonical total order used wherever the reference's own order is
    implementation-defined (unstable ``argsort`` at extracter.py:218): score
    descending, raster index ascending."""
    return np.lexsort((raster, -p.astype(np.float64)))


def detection(score_map, params: dict | None = None, nms: str = "separable"):
    """utils/extracter.py:193-221.  Returns (pts[N,3] float32, raster[N] int64).

    ``K > top_k`` -> rows sorted by score descending, truncated (extracter.py:217-218;
    ties broken canonically, see ``canonical_order``); otherwise raster order.
    ``min_score`` is applied after the truncation (extracter.py:219-220)."""
    if params is None:
        nms_dist, threshold, border_dist, top_k, min_score = 4, 0.0, 8, 300, 0.0
    else:
        nms_dist = params['nms_dist']
        threshold = params['threshold']
        border_dist = params['border_dist']
        top_k = params['top_k']
        min_score = params['min_score']
    s = score_map.detach().cpu().numpy() if isinstance(score_map, torch.Tensor) else np.asarray(score_map)
    s = s.astype(np.float32)
    if nms == "im2col":
        t = torch.from_numpy(s.reshape((-1, 1) + s.shape[-2:]))
        kept = nms_rounds_im2col(t, nms_dist).numpy()
    elif nms == "greedy":
        kept = nms_greedy(s.reshape((-1,) + s.shape[-2:])[0], nms_dist)
    else:
        kept = nms_rounds_separable(s.reshape((-1,) + s.shape[-2:]), nms_dist)
    kept = clear_border(kept, border_dist)
    pts, raster = positions_with_prob(kept, threshold)
    if pts.shape[0] > top_k:
        sel = canonical_order(pts[:, 2], raster)[:top_k]
        pts, raster = pts[sel], raster[sel]
    if min_score > 0:
        keep = pts[:, 2] > np.float32(min_score)
        pts, raster = pts[keep], raster[keep]
    return pts, raster


# --------------------------------------------------------------------------------------
# Stage 2: bilinear descriptor sampling  (utils/matcher.py:221-226, models/lightglue.py:24-41)
# --------------------------------------------------------------------------------------


def bilinear_gather(desc_map: np.ndarray, px: np.ndarray, py: np.ndarray) -> np.ndarray:
    """Bilinear tap of a [C,h,w] map at float pixel coordinates (zero padding outside),
    i.e. what ``grid_sample(mode='bilinear', padding_mode='zeros')`` evaluates once the
    grid is un-normalised.  Returns [n, C] float32."""
    c, h, w = desc_map.shape
    px = px.astype(np.float32)
    py = py.astype(np.float32)
    x0 = np.floor(px)
    y0 = np.floor(py)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - px) * (y1 - py)
    w_ne = (px - x0) * (y1 - py)
    w_sw = (x1 - px) * (py - y0)
    w_se = (px - x0) * (py - y0)
    out = np.zeros((px.shape[0], c), dtype=np.float32)
    for xs, ys, ws in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
        xi = xs.astype(np.int64)
        yi = ys.astype(np.int64)
        ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
        xi = np.clip(xi, 0, w - 1)
        yi = np.clip(yi, 0, h - 1)
        tap = desc_map[:, yi, xi].T  # [n, C]
        out += np.where(ok[:, None], tap * ws[:, None].astype(np.float32), np.float32(0))
    return out


def sample_brute_force(desc_map: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """utils/matcher.py:221-226 -- g=(pts-0.5)*2, ``grid_sample(align_corners=True)`` so
    the tap lands at ((g+1)/2)*(size-1) of the DESCRIPTOR map; no normalisation.
    ``desc_map`` [1,C,h,w] or [C,h,w]; ``pts`` [n,>=2] normalised (x,y).  -> [n,C]."""
    d = np.asarray(desc_map, dtype=np.float32)
    if d.ndim == 4:
        d = d[0]
    _, h, w = d.shape
    g = (np.asarray(pts, dtype=np.float32)[:, :2] - np.float32(0.5)) * np.float32(2)
    px = ((g[:, 0] + np.float32(1)) / np.float32(2)) * np.float32(w - 1)
    py = ((g[:, 1] + np.float32(1)) / np.float32(2)) * np.float32(h - 1)
    return bilinear_gather(d, px, py)


def sample_lightglue(desc_map: np.ndarray, kpts_px: np.ndarray, s: int = 8) -> np.ndarray:
    """models/lightglue.py:24-41 -- pixel keypoints -> (k - s/2 + 0.5)/(size*s - s/2 - 0.5)*2-1,
    bilinear ``grid_sample(align_corners=True)``, then x / max(||x||_2, 1e-12).  -> [n,C]."""
    d = np.asarray(desc_map, dtype=np.float32)
    if d.ndim == 4:
        d = d[0]
    _, h, w = d.shape
    k = np.asarray(kpts_px, dtype=np.float32)[:, :2] - np.float32(s / 2) + np.float32(0.5)
    k = k / np.array([w * s - s / 2 - 0.5, h * s - s / 2 - 0.5], dtype=np.float32)
    g = k * np.float32(2) - np.float32(1)
    px = ((g[:, 0] + np.float32(1)) / np.float32(2)) * np.float32(w - 1)
    py = ((g[:, 1] + np.float32(1)) / np.float32(2)) * np.float32(h - 1)
    out = bilinear_gather(d, px, py)
    nrm = np.sqrt((out.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
    return out / np.maximum(nrm, np.float32(1e-12))[:, None]


# --------------------------------------------------------------------------------------
# LightGlue-style extraction  (models/lightglue.py:904-979)
# --------------------------------------------------------------------------------------


def simple_nms(scores: np.ndarray, nms_radius: int) -> np.ndarray:
    """models/lightglue.py:904-920 -- max-pool NMS: maxima of the (2r+1)^2 window (ties all kept),
    then two rounds in which pixels outside the dilated maxima may become maxima of the
    zero-suppressed map.  ``max_pool2d`` pads with -inf.  [H,W] or [B,H,W] -> same shape."""
    from scipy.ndimage import maximum_filter
    assert nms_radius >= 0                                          # :906
    s = np.asarray(scores, dtype=np.float32)
    k = 2 * nms_radius + 1
    size = (1,) * (s.ndim - 2) + (k, k)

    def pool(x):                                                    # :908-911
        return maximum_filter(x, size=size, mode='constant', cval=-np.inf)

    keep = s == pool(s)                                             # :914
    for _ in range(2):                                              # :915-919
        near = pool(keep.astype(np.float32)) > 0
        rest = np.where(near, np.float32(0), s)
        keep = keep | ((rest == pool(rest)) & ~near)
    return np.where(keep, s, np.float32(0))                         # :920


def lightglue_extract(scores: np.ndarray, desc_map: np.ndarray, s: int, detection_threshold: float = 0.0,
                      pad: int = 4, nms_radius: int = 5, max_num_kps: int | None = 1000):
    """models/lightglue.py:929-979 after the network call, for ONE image (the reference indexes
    ``scores[0]``, :938): simple_nms; the ``pad`` border rows/cols become -1 (:942-945); pixels
    ``> detection_threshold`` in raster order (:948-955); if more than ``max_num_kps``, the best by
    score (``torch.topk``, sorted; tie order canonicalised here as raster asc) (:923-927, 958-967);
    keypoints as float (x, y) pixels (:970); descriptors via ``sample_descriptors`` (:972-975).
    -> keypoints [n,2], keypoint_scores [n], descriptors [n,C], raster [n]."""
    sc = simple_nms(np.asarray(scores, dtype=np.float32).reshape(scores.shape[-2:]), nms_radius).copy()
    h, w = sc.shape
    if pad > 0:
        sc[:pad] = -1
        sc[:, :pad] = -1
        sc[-pad:] = -1
        sc[:, -pad:] = -1
    ys, xs = np.nonzero(sc > np.float32(detection_threshold))
    val = sc[ys, xs]
    raster = ys.astype(np.int64) * w + xs
    if max_num_kps is not None and max_num_kps < val.shape[0]:
        order = canonical_order(val, raster)[:max_num_kps]
        ys, xs, val, raster = ys[order], xs[order], val[order], raster[order]
    kps = np.stack([xs, ys], axis=1).astype(np.float32)
    desc = sample_lightglue(desc_map, kps, s) if kps.shape[0] else np.zeros((0, np.asarray(desc_map).shape[-3]), np.float32)
    return kps, val, desc, raster


# --------------------------------------------------------------------------------------
# Tensor Lucas-Kanade matcher  (utils/matcher.py:7-142, 188-203)
# --------------------------------------------------------------------------------------


def lk_pyramid(img: np.ndarray, levels: int):
    """matcher.py:38-47 -- level j >= 1 is ``avg_pool2d(img, kernel=2j, stride=2j)`` of the ORIGINAL
    image (kernel 2, 4, 6 ... while the coordinates are scaled by 2^j -- the mismatch from level 3 on
    is the reference's).  img [C,H,W] -> list of [C,Hj,Wj]."""
    img = np.asarray(img, dtype=np.float32)
    out = [img]
    c, h, w = img.shape
    for j in range(1, levels):
        k = 2 * j
        hh, ww = h // k, w // k
        acc = np.zeros((c, hh, ww), dtype=np.float32)
        for a in range(k):                      # row-major running sum like the CPU pooling loop
            for b in range(k):
                acc += img[:, a:a + hh * k:k, b:b + ww * k:k]
        out.append(acc / np.float32(k * k))
    return out


def lk_sobel(img: np.ndarray):
    """matcher.py:22-35, 104-109 -- per-channel 3x3 cross-correlation with [[1,0,-1],[2,0,-2],[1,0,-1]]
    (dx) and its transpose (dy), zero padding 1.  [C,H,W] -> (dx, dy)."""
    p = np.pad(np.asarray(img, dtype=np.float32), ((0, 0), (1, 1), (1, 1)))
    tl, tc, tr = p[:, :-2, :-2], p[:, :-2, 1:-1], p[:, :-2, 2:]
    ml, mr = p[:, 1:-1, :-2], p[:, 1:-1, 2:]
    bl, bc, br = p[:, 2:, :-2], p[:, 2:, 1:-1], p[:, 2:, 2:]
    two = np.float32(2)
    dx = (tl - tr) + two * (ml - mr) + (bl - br)
    dy = (tl + two * tc + tr) - (bl + two * bc + br)
    return dx.astype(np.float32), dy.astype(np.float32)


def lk_patch_sample(img: np.ndarray, pts_px: np.ndarray, win: int) -> np.ndarray:
    """``grid_sample(unfold(img, win, padding=win//2).view(1,-1,H,W), pts/(W-1,H-1)*2-1, align_corners=True)``
    (matcher.py:113-128, 134) without materialising the unfold: entry (c,u,v) of the unfolded map at pixel
    (y,x) is img[c, y+u-p, x+v-p] (0 outside), and the bilinear tap adds the four integer corners that lie
    INSIDE the map, so a corner outside contributes nothing even when its shifted pixel is inside.
    img [C,H,W], pts [n,2] pixels -> [n, C*win*win] (channel order c, u, v)."""
    c, h, w = img.shape
    pad = win // 2
    pts = np.asarray(pts_px, dtype=np.float32)
    g = pts / np.array([w - 1, h - 1], dtype=np.float32) * np.float32(2) - np.float32(1)
    ix = ((g[:, 0] + np.float32(1)) / np.float32(2)) * np.float32(w - 1)
    iy = ((g[:, 1] + np.float32(1)) / np.float32(2)) * np.float32(h - 1)
    x0, y0 = np.floor(ix), np.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    wts = ((x1 - ix) * (y1 - iy), (ix - x0) * (y1 - iy), (x1 - ix) * (iy - y0), (ix - x0) * (iy - y0))
    corners = ((x0, y0), (x1, y0), (x0, y1), (x1, y1))
    padded = np.pad(img, ((0, 0), (pad, pad), (pad, pad)))
    n = pts.shape[0]
    out = np.zeros((n, c, win, win), dtype=np.float32)
    du = np.arange(win)
    for (xs, ys), wt in zip(corners, wts):
        ok = (xs >= 0) & (xs <= w - 1) & (ys >= 0) & (ys <= h - 1) & np.isfinite(xs) & np.isfinite(ys)
        xi = np.where(ok, xs, 0).astype(np.int64)
        yi = np.where(ok, ys, 0).astype(np.int64)
        rows = (yi[:, None] + du[None, :])                       # padded row of (u): y + u - pad + pad
        cols = (xi[:, None] + du[None, :])
        tap = padded[:, rows[:, :, None], cols[:, None, :]]       # [C, n, win, win]
        tap = np.transpose(tap, (1, 0, 2, 3))
        out += np.where(ok[:, None, None, None], tap * wt.astype(np.float32)[:, None, None, None], np.float32(0))
    return out.reshape(n, c * win * win)


def lk_level(img1: np.ndarray, img2: np.ndarray, pts1: np.ndarray, pts2: np.ndarray, win: int, iterations: int):
    """matcher.py:94-142 -- ``iterations`` Gauss-Newton steps on one pyramid level.  The update is the
    reference's ``einsum('bik,bk->bk', G_inverse, b)`` (:139): component b of the right-hand side times the
    ROW SUM of the inverse (not a matrix-vector product) -- part of the contract.  Points whose 2x2 normal
    matrix has det <= 1e-6 do not move (:137).  Returns the new points [n,2] float32."""
    dx2, dy2 = lk_sobel(img2)
    ref_patch = lk_patch_sample(img1, pts1, win)
    p = np.array(pts2, dtype=np.float32, copy=True)
    for _ in range(iterations):
        d_i = ref_patch - lk_patch_sample(img2, p, win)
        jx = lk_patch_sample(dx2, p, win)
        jy = lk_patch_sample(dy2, p, win)
        gxx = (jx * jx).sum(1, dtype=np.float32)
        gxy = (jx * jy).sum(1, dtype=np.float32)
        gyy = (jy * jy).sum(1, dtype=np.float32)
        bx = (d_i * jx).sum(1, dtype=np.float32)
        by = (d_i * jy).sum(1, dtype=np.float32)
        det = gxx.astype(np.float64) * gyy - gxy.astype(np.float64) * gxy
        ok = det.astype(np.float32) > np.float32(1e-6)
        safe = np.where(ok, det, 1.0)
        i00, i01 = (gyy / safe).astype(np.float32), (-gxy / safe).astype(np.float32)
        i11 = (gxx / safe).astype(np.float32)
        step = np.stack([i00 * bx + i01 * bx, i01 * by + i11 * by], axis=1).astype(np.float32)
        p = np.where(ok[:, None], p - step, p).astype(np.float32)
    return p


def lk_init_points(pts1_norm: np.ndarray, h: int, w: int, distance: float, angle: np.ndarray) -> np.ndarray:
    """matcher.py:52-61 -- start points: ``pts1*(W-1,H-1) + (cos, sin)(angle)*distance`` clamped to
    x in [10, W-10], y in [10, H-10]; ``angle`` is the reference's ``randn(n)*6.28`` draw."""
    p = np.asarray(pts1_norm, dtype=np.float32)[:, :2] * np.array([w - 1, h - 1], dtype=np.float32)
    a = np.asarray(angle, dtype=np.float32)
    p = p + np.stack([np.cos(a), np.sin(a)], axis=1).astype(np.float32) * np.float32(distance)
    p[:, 0] = np.clip(p[:, 0], 10, w - 10)
    p[:, 1] = np.clip(p[:, 1], 10, h - 10)
    return p.astype(np.float32)


def lk_track(img0: np.ndarray, img1: np.ndarray, pts0_px: np.ndarray, init_px: np.ndarray, win: int, levels: int,
             iterations: int) -> np.ndarray:
    """``OpticalFlow.__call__`` (matcher.py:49-92) after the random start has been drawn: coarse-to-fine over
    the pyramid, coordinates divided by 2^level going down and multiplied back coming up.  img [C,H,W],
    points in PIXELS of the full-resolution image -> tracked points [n,2] in pixels (what the reference
    returns: ``optical_flow_tensor`` does not re-normalise, :201-203)."""
    pyr0, pyr1 = lk_pyramid(img0, levels), lk_pyramid(img1, levels)
    p0 = np.asarray(pts0_px, dtype=np.float32)
    cur = np.asarray(init_px, dtype=np.float32)
    for i in range(levels):
        lvl = levels - i - 1
        sc = np.float32(2 ** lvl)
        cur = lk_level(pyr0[lvl], pyr1[lvl], p0 / sc, cur / sc, win, iterations) * sc
    return cur


# --------------------------------------------------------------------------------------
# Stage 3: brute-force mutual-NN matching  (utils/matcher.py:227-234 -> skimage, absent)
# --------------------------------------------------------------------------------------


def match_descriptors(d0: np.ndarray, d1: np.ndarray, metric=None, p=2, max_distance=np.inf,
                      cross_check=True, max_ratio=1.0) -> np.ndarray:
    """Published algorithm of ``skimage.feature.match_descriptors`` (scikit-image,
    unpinned at requirements.txt:13; call site utils/matcher.py:227-230): float64
    ``cdist``; row argmin (first of ties); optional cross-check against the column
    argmin; strict ``< max_distance`` gate; optional Lowe ratio.  -> int64 [k,2] sorted
    by the first index."""
    if d0.shape[1] != d1.shape[1]:
        raise ValueError("Descriptor length must equal.")
    dist = cdist(d0, d1, metric='euclidean' if metric is None else metric)
    i0 = np.arange(d0.shape[0])
    i1 = np.argmin(dist, axis=1)
    if cross_check:
        back = np.argmin(dist, axis=0)
        keep = i0 == back[i1]
        i0, i1 = i0[keep], i1[keep]
    if max_distance < np.inf:
        keep = dist[i0, i1] < max_distance
        i0, i1 = i0[keep], i1[keep]
    if max_ratio < 1.0:
        best = dist[i0, i1]
        dist[i0, i1] = np.inf
        second = np.min(dist[i0], axis=1)
        second[second == 0] = np.finfo(np.double).eps
        keep = best / second < max_ratio
        i0, i1 = i0[keep], i1[keep]
    return np.column_stack((i0, i1)).astype(np.int64)


def brute_force_matcher(pts0, pts1, desc_map_0, desc_map_1, params: dict):
    """utils/matcher.py:206-234 -- sample both descriptor maps at the keypoints, match,
    return the FULL matched rows of pts0/pts1 plus the index pairs."""
    pts0 = np.asarray(pts0, dtype=np.float32)
    pts1 = np.asarray(pts1, dtype=np.float32)
    d0 = sample_brute_force(desc_map_0, pts0)
    d1 = sample_brute_force(desc_map_1, pts1)
    pairs = match_descriptors(d0, d1, metric=params['metric'], max_distance=params['max_distance'],
                              cross_check=params['cross_check'])
    return pts0[pairs[:, 0]], pts1[pairs[:, 1]], pairs


# --------------------------------------------------------------------------------------
# Stage 4: homography projection + repeatability / MHA counting
# (utils/projection.py:128-192, tasks/repeatability.py:9-92, tasks/MHA.py:40-72)
# --------------------------------------------------------------------------------------


def _as_int(v):
    if isinstance(v, torch.Tensor):
        return int(v.item())
    return int(np.asarray(v).item()) if not isinstance(v, (int, np.integer)) else int(v)


def warp_homography(kpts, params: dict):
    """utils/projection.py:137-167 -- p = kpts*[w-1,h-1]; q = H [p,1]; q/=q_z;
    valid iff 0<=x<=w-1 and 0<=y<=h-1; returns (kpts_valid, warped_valid) divided back
    by [w-1,h-1], the valid ids and the invalid ids.  float32 arithmetic."""
    k = np.asarray(kpts, dtype=np.float32)[:, :2]
    hm = params['homography_matrix']
    hm = hm.detach().cpu().numpy() if isinstance(hm, torch.Tensor) else np.asarray(hm)
    hm = hm.astype(np.float32)
    w, h = _as_int(params['width']), _as_int(params['height'])
    scale = np.array([w - 1, h - 1], dtype=np.float32)
    p = k * scale
    x, y = p[:, 0], p[:, 1]
    one = np.float32(1)
    q = np.stack([hm[i, 0] * x + hm[i, 1] * y + hm[i, 2] * one for i in range(3)], axis=1).astype(np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        uv = q[:, :2] / q[:, 2:]
    valid = (uv[:, 0] >= 0) & (uv[:, 0] <= w - 1) & (uv[:, 1] >= 0) & (uv[:, 1] <= h - 1)
    ids = np.flatnonzero(valid).astype(np.int64)
    ids_out = np.flatnonzero(~valid).astype(np.int64)
    return (p[valid] / scale).astype(np.float32), (uv[valid] / scale).astype(np.float32), ids, ids_out


def _np32(v):
    return (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)).astype(np.float32)


def interpolate_depth(pos: np.ndarray, depth: np.ndarray):
    """utils/projection.py:270-372 -- pos [N,2] (x,y) in pixels.  Corners floor/ceil of (i=y, j=x) must lie
    inside a 10-pixel border (:289-301), all four corner depths must be > 0 (:321-324), bilinear weights
    from the floor corner (:346-357).  Returns (depth, pos_valid, ids, ids_valid_corners, ids_valid_depth)."""
    border = 10
    pos = np.asarray(pos, dtype=np.float32)
    h, w = depth.shape
    i, j = pos[:, 1], pos[:, 0]
    with np.errstate(invalid='ignore'):
        i0, j0 = np.floor(i).astype(np.int64), np.floor(j).astype(np.int64)
        i1, j1 = np.ceil(i).astype(np.int64), np.ceil(j).astype(np.int64)
    valid_corners = (i0 >= border) & (j0 >= border) & (j1 < w - border) & (i1 < h - border)
    ids_c = np.flatnonzero(valid_corners).astype(np.int64)
    i0c, j0c, i1c, j1c = i0[ids_c], j0[ids_c], i1[ids_c], j1[ids_c]
    vd = (depth[i0c, j0c] > 0) & (depth[i0c, j1c] > 0) & (depth[i1c, j0c] > 0) & (depth[i1c, j1c] > 0)
    ids = ids_c[vd]
    i0v, j0v, i1v, j1v = i0c[vd], j0c[vd], i1c[vd], j1c[vd]
    di = (i[ids] - i0v.astype(np.float32)).astype(np.float32)
    dj = (j[ids] - j0v.astype(np.float32)).astype(np.float32)
    one = np.float32(1)
    w_tl, w_tr = (one - di) * (one - dj), (one - di) * dj
    w_bl, w_br = di * (one - dj), di * dj
    z = (w_tl * depth[i0v, j0v] + w_tr * depth[i0v, j1v] + w_bl * depth[i1v, j0v] + w_br * depth[i1v, j1v]).astype(np.float32)
    return z, pos[ids], ids, ids_c, ids


def warp_se3(kpts, params: dict):
    """utils/projection.py:194-267 -- depth-based covisibility: keypoints (normalised x,y) are scaled by
    (w0,h0), given a depth by interpolate_depth (view 0), unprojected with K0 (crop offset bbox0 + 0.5, COLMAP
    convention), moved by pose01, projected with K1, and checked against view 1's interpolated depth
    (|z_proj - z_interp| < 0.05).  Returns (kpts0_valid, kpts01_valid, ids_valid, ids_out) with
    ids_out = [projected outside view 1's valid-corner area] ++ [occluded]."""
    k = np.asarray(kpts, dtype=np.float32)[:, :2]
    depth0, depth1 = _np32(params['depth0']), _np32(params['depth1'])
    k0m, k1m, pose = _np32(params['intrinsics0']), _np32(params['intrinsics1']), _np32(params['pose01'])
    bbox0, bbox1 = _np32(params['bbox0']), _np32(params['bbox1'])
    wh0 = np.array([depth0.shape[1], depth0.shape[0]], dtype=np.float32)
    wh1 = np.array([depth1.shape[1], depth1.shape[0]], dtype=np.float32)
    px = k * wh0
    z0, k0v, ids0, _, _ = interpolate_depth(px, depth0)
    bk = k0v + bbox0[[1, 0]][None, :] + np.float32(0.5)
    duv1 = np.concatenate([bk * z0[:, None], z0[:, None]], axis=1).astype(np.float32)
    kinv = np.linalg.inv(k0m.astype(np.float64)).astype(np.float32)
    p3 = (duv1 @ kinv.T).astype(np.float32)
    p3h = np.concatenate([p3, np.ones((p3.shape[0], 1), np.float32)], axis=1)
    p31 = (p3h @ pose.T)[:, :3].astype(np.float32)
    zuv = (p31 @ k1m.T).astype(np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        uv = (zuv / zuv[:, 2:3])[:, :2]
    z01 = zuv[:, 2]
    uv01 = (uv - bbox1[[1, 0]][None, :] - np.float32(0.5)).astype(np.float32)
    z01i, k01v, ids01, ids01_c, _ = interpolate_depth(uv01, depth1)
    outside = np.ones(ids0.shape[0], bool)
    outside[ids01_c] = False
    ids_outside = ids0[outside]
    ids_valid = ids0[ids01]
    k0v = k0v[ids01]
    inlier = np.abs(z01[ids01] - z01i) < 0.05
    ids_occ = ids_valid[~inlier]
    return ((k0v[inlier] / wh0).astype(np.float32), (k01v[inlier] / wh1).astype(np.float32), ids_valid[inlier],
            np.concatenate([ids_outside, ids_occ]))


def warp(kpts, params: dict):
    """utils/projection.py:185-192 -- dispatch on params['mode']."""
    if params['mode'] == 'homo':
        return warp_homography(np.asarray(kpts)[:, 0:2], params)
    if params['mode'] == 'se3':
        return warp_se3(np.asarray(kpts)[:, 0:2], params)
    raise ValueError('unknown mode!')


def keypoint_distance(k0: np.ndarray, k1: np.ndarray) -> np.ndarray:
    """tasks/repeatability.py:39-51 -- [M,N] Euclidean distances of normalised points (float32)."""
    d = k0[:, None, :].astype(np.float32) - k1[None, :, :].astype(np.float32)
    return np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32)


def mutual_argmin(value: np.ndarray):
    """tasks/repeatability.py:9-36 -- v = (-value) - min(-value) in float32 (this rounding
    step merges nearby distances once the 99999 diagonal is present); every entry equal to
    both its row max and its column max is returned, raster order."""
    neg = (-value).astype(np.float32)
    v = (neg - neg.min()).astype(np.float32)
    row = v.max(axis=1, keepdims=True)
    col = v.max(axis=0, keepdims=True)
    return np.nonzero((v == row) & (v == col))


def val_key_points(kps0, kps1, warp01: dict, warp10: dict, th: float = 3):
    """tasks/repeatability.py:54-92."""
    kps0 = np.asarray(kps0, dtype=np.float32)
    kps1 = np.asarray(kps1, dtype=np.float32)
    num_feat = min(kps0.shape[0], kps1.shape[0])
    k0c, k01c, _, _ = warp(kps0, warp01)
    k1c, k10c, _, _ = warp(kps1, warp10)
    if k0c.shape[0] == 0 or k1c.shape[0] == 0:
        return {'num_feat': 0, 'repeatability': 0, 'mean_error': 0, 'errors': None, 'gt_num': 0}
    d01 = keypoint_distance(k0c, k10c)
    d10 = keypoint_distance(k1c, k01c)
    dm = ((d01 + d10.T) / np.float32(2)).astype(np.float32)
    n = min(dm.shape)
    dm[np.arange(n), np.arange(n)] = np.float32(99999)
    ii, jj = mutual_argmin(dm)
    scale = np.float32(_as_int(warp01['resize']) if 'resize' in warp01 else _as_int(warp01['width']))
    scale10 = np.float32(_as_int(warp10['resize']) if 'resize' in warp01 else _as_int(warp10['width']))
    dist = dm[ii, jj] * scale
    good = dist <= th
    gt_num = int(good.sum())
    with np.errstate(invalid='ignore'):
        mean_error = float(dist[good].mean()) if gt_num else float('nan')
    errors = (dm * scale10).min(axis=1)
    return {'num_feat': num_feat, 'repeatability': gt_num / num_feat if num_feat else 0,
            'mean_error': mean_error, 'errors': errors, 'gt_num': gt_num,
            'pairs': np.stack([ii, jj], axis=1)}


def corner_error_flags(h_est: np.ndarray, h_real: np.ndarray, w: int, h: int, resize_h: int,
                       resize_w: int, th=(3, 5, 7)):
    """tasks/MHA.py:51-72 -- project the four (x/y-swapped, as in the reference) corners with
    the real and the estimated homography, rescale, mean L2, one 0/1 flag per threshold."""
    corners = np.array([[0, 0, 1], [h - 1, 0, 1], [0, w - 1, 1], [h - 1, w - 1, 1]])
    real = corners @ np.asarray(h_real).T
    real = real[:, :2] / real[:, 2:]
    est = corners @ np.asarray(h_est).T
    est = est[:, :2] / est[:, 2:]
    sc = np.array([resize_h / h, resize_w / w])
    mean_dist = np.mean(np.linalg.norm(real * sc - est * sc, axis=1))
    return [float(mean_dist <= t) for t in th], float(mean_dist)


def mha_pair(score0, desc0, score1, desc1, warp01, warp10, params, resize_hw):
    """tasks/MHA.py:11-72 minus the PNG/IO: detect -> covisible -> match -> pixels ->
    cv2.findHomography(RANSAC) on the host -> corner flags."""
    import cv2
    th = params['MHA_params']['th']
    flags = [0 for _ in th]
    k0, _ = detection(score0, params['extractor_params'])
    k1, _ = detection(score1, params['extractor_params'])
    k0c, _, _, _ = warp(k0, warp01)
    k1c, _, _, _ = warp(k1, warp10)
    if k0c.shape[0] == 0 or k1c.shape[0] == 0:
        return flags, None
    m0, m1, pairs = brute_force_matcher(k0c, k1c, desc0, desc1, params['matcher_params']['brute_force_params'])
    h, w = _as_int(warp01['height']), _as_int(warp01['width'])
    sc = np.array([w - 1, h - 1], dtype=np.float32)
    if m0.shape[0] < 4:
        return flags, pairs
    hm, _ = cv2.findHomography(m0[:, :2] * sc, m1[:, :2] * sc, cv2.RANSAC)
    if hm is None:
        return flags, pairs
    hr = warp01['homography_matrix']
    hr = hr.detach().cpu().numpy() if isinstance(hr, torch.Tensor) else np.asarray(hr)
    fl, _ = corner_error_flags(hm, hr, w, h, resize_hw[0], resize_hw[1], th)
    return fl, pairs
