"""Mint golden fixtures from the UNMODIFIED reference (run in the authoring container only).

    python -m oracle.make_golden            # writes tests/golden/*.npz and oracle/REFCHECK.log

The reference has no tests or known-answer vectors for this path, so the
oracle is pinned against the reference's own outputs: each fixture stores the reference
result for a seeded input (inputs are regenerated from the seed by
``keypoint_bench_b200.synth`` -- iid kinds are bit-reproducible -- or stored when small).
While writing the fixtures the script also asserts that the restatements in
``oracle/ref_ops.py`` reproduce the reference on every case; the outcome is logged.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from keypoint_bench_b200 import synth  # noqa: E402
from oracle import _refimport, ref_ops  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
LOG = []


def log(msg):
    print(msg, flush=True)
    LOG.append(msg)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


NMS_CASES = [
    # (kind, h, w, seed, nms_dist, min_value, max_iter)
    ('uniform', 48, 64, 1, 4, 0.0, -1),
    ('uniform', 61, 83, 2, 6, 0.0, -1),
    ('uniform', 40, 56, 3, 2, 0.0, -1),
    ('uniform', 40, 56, 3, 8, 0.0, -1),
    ('uniform', 40, 56, 3, 0, 0.0, -1),
    ('uniform', 33, 47, 4, 1, 0.0, -1),
    ('ties', 48, 64, 5, 4, 0.0, -1),
    ('ties', 57, 70, 6, 6, 0.0, -1),
    ('ramp', 24, 160, 7, 6, 0.0, -1),
    ('negative', 48, 64, 8, 4, 0.0, -1),
    ('negative', 40, 56, 9, 2, 0.0, -1),
    ('mixed', 48, 64, 10, 4, 0.0, -1),
    ('mixed', 64, 96, 11, 6, 0.0, -1),
    ('relu', 48, 64, 12, 4, 0.0, -1),
    ('uniform', 48, 64, 13, 4, 0.0, 2),      # max_iter cap
    ('uniform', 48, 64, 14, 4, -1.0, -1),    # non-zero suppression value
    ('uniform', 48, 64, 15, 4, 0.25, -1),    # suppression value above some scores
    ('alike', 64, 96, 16, 6, 0.0, -1),
    ('uniform', 5, 7, 17, 6, 0.0, -1),       # window larger than the image
    ('uniform', 1, 40, 18, 3, 0.0, -1),
]

DETECT_CASES = [
    # (tag, kind, h, w, seed, params)
    ('cfg1_480x640', 'uniform', 480, 640, synth.pair_seed(1, 0),
     dict(nms_dist=6, threshold=0, border_dist=8, top_k=1000, min_score=0.0)),
    ('cfg3_480x640_r4_k4096', 'uniform', 480, 640, synth.pair_seed(3, 0),
     dict(nms_dist=4, threshold=0, border_dist=8, top_k=4096, min_score=0.0)),
    ('cfg5_376x1241', 'uniform', 376, 1241, synth.pair_seed(5, 0),
     dict(nms_dist=6, threshold=0, border_dist=8, top_k=1000, min_score=0.0)),
    ('raster_order_k_le_topk', 'uniform', 120, 160, 77,
     dict(nms_dist=6, threshold=0, border_dist=8, top_k=1000, min_score=0.0)),
    ('threshold_minscore', 'uniform', 120, 160, 78,
     dict(nms_dist=4, threshold=0.5, border_dist=4, top_k=100, min_score=0.995)),
    ('no_nms_border0', 'uniform', 24, 32, 79,
     dict(nms_dist=0, threshold=0.9, border_dist=0, top_k=50, min_score=0.0)),
    ('mixed_sign', 'mixed', 96, 128, 80,
     dict(nms_dist=4, threshold=0, border_dist=8, top_k=60, min_score=0.0)),
    ('ramp', 'ramp', 32, 200, 81,
     dict(nms_dist=6, threshold=0, border_dist=8, top_k=1000, min_score=0.0)),
]


def gen_nms(ref):
    out = {}
    for i, (kind, h, w, seed, r, mv, mi) in enumerate(NMS_CASES):
        s = synth.score_map(kind, h, w, seed)
        got = ref.extracter.fast_nms(s.clone(), nms_dist=r, max_iter=mi, min_value=mv).numpy()
        mine = ref_ops.nms_rounds_separable(s.numpy(), r, mi, mv)
        im2 = ref_ops.nms_rounds_im2col(s.clone(), r, mi, mv).numpy()
        assert np.array_equal(got, mine), f'nms separable restatement differs on case {i}'
        assert np.array_equal(got, im2), f'nms im2col restatement differs on case {i}'
        if kind in ('uniform', 'ties', 'ramp', 'relu', 'alike') and mv == 0.0 and mi == -1 and r > 0:
            gr = ref_ops.nms_greedy(s.numpy()[0, 0], r)
            assert np.array_equal(got[0, 0], gr), f'greedy closed form differs on case {i}'
        out[f'in_{i}'] = s.numpy()[0, 0]
        out[f'out_{i}'] = got[0, 0]
        log(f'nms case {i} {kind} {h}x{w} r={r} min_value={mv} max_iter={mi}: '
            f'kept={int((got != mv).sum()) if r else -1} restatements equal')
    out['cases'] = np.array([repr(c) for c in NMS_CASES])
    np.savez_compressed(os.path.join(GOLD, 'ref_nms.npz'), **out)


def gen_detect(ref):
    out = {}
    for tag, kind, h, w, seed, params in DETECT_CASES:
        s = synth.score_map(kind, h, w, seed)
        t0 = time.time()
        got = ref.extracter.detection(s.clone(), dict(params)).numpy()
        dt = time.time() - t0
        mine, raster = ref_ops.detection(s.clone(), dict(params))
        tie_free = len(np.unique(got[:, 2])) == got.shape[0]
        if tie_free:
            assert np.array_equal(got, mine), f'detection restatement differs on {tag}'
        else:  # canonical comparison (SURVEY 8(c) rule ii)
            assert np.array_equal(np.sort(got[:, 2]), np.sort(mine[:, 2])), tag
        out[f'{tag}__pts'] = got
        out[f'{tag}__raster'] = raster
        out[f'{tag}__insha'] = np.array(sha(s.numpy()))
        log(f'detection {tag}: N={got.shape[0]} tie_free={tie_free} reference_cpu_s={dt:.2f} restatement equal')
    np.savez_compressed(os.path.join(GOLD, 'ref_detect.npz'), **out)


def gen_match(ref):
    out = {}
    cases = [('small32', 32, 12, 16, 150, 140, 5.0, True), ('sp256', 256, 60, 80, 300, 280, 5.0, True),
             ('nocross', 64, 20, 24, 200, 220, 1.2, False), ('tight', 64, 20, 24, 200, 220, 0.9, True)]
    for tag, c, h, w, n, m, maxd, cc in cases:
        g = torch.Generator().manual_seed(abs(hash(tag)) % 10000 if False else sum(map(ord, tag)))
        d0 = torch.nn.functional.normalize(torch.randn(1, c, h, w, generator=g), dim=1)
        d1 = d0 + 0.08 * torch.randn(1, c, h, w, generator=g)
        p0 = torch.rand(n, 3, generator=g)
        p1 = torch.cat([p0[: m // 2, :2] + 0.002 * torch.randn(m // 2, 2, generator=g),
                        torch.rand(m - m // 2, 2, generator=g)], dim=0)
        p1 = torch.cat([p1.clamp(0, 1), torch.rand(m, 1, generator=g)], dim=1)
        params = {'metric': 'euclidean', 'max_distance': maxd, 'cross_check': cc}
        r0, r1 = ref.matcher.brute_force_matcher(p0, p1, d0, d1, params)
        # the sampled descriptors exactly as utils/matcher.py:221-226 computes them
        s0 = torch.nn.functional.grid_sample(d0, ((p0[:, :2] - 0.5) * 2)[None, None], align_corners=True)[0, :, 0].T
        s1 = torch.nn.functional.grid_sample(d1, ((p1[:, :2] - 0.5) * 2)[None, None], align_corners=True)[0, :, 0].T
        pairs = ref_ops.match_descriptors(s0.numpy(), s1.numpy(), metric='euclidean', max_distance=maxd, cross_check=cc)
        m0, m1, mypairs = ref_ops.brute_force_matcher(p0.numpy(), p1.numpy(), d0.numpy(), d1.numpy(), params)
        assert np.allclose(ref_ops.sample_brute_force(d0.numpy(), p0.numpy()), s0.numpy(), atol=1e-6), tag
        assert np.array_equal(pairs, mypairs), f'match restatement differs on {tag}'
        assert np.array_equal(r0.numpy(), m0) and np.array_equal(r1.numpy(), m1), tag
        small = c * h * w <= 32 * 12 * 16 * 4
        out[f'{tag}__desc0'] = d0.numpy() if small else np.zeros(0, np.float32)
        out[f'{tag}__desc1'] = d1.numpy() if small else np.zeros(0, np.float32)
        out[f'{tag}__seed'] = np.array(sum(map(ord, tag)))
        out[f'{tag}__shape'] = np.array([c, h, w, n, m])
        out[f'{tag}__maxd_cc'] = np.array([maxd, float(cc)])
        out[f'{tag}__p0'] = p0.numpy()
        out[f'{tag}__p1'] = p1.numpy()
        out[f'{tag}__s0'] = s0.numpy().astype(np.float32)
        out[f'{tag}__s1'] = s1.numpy().astype(np.float32)
        out[f'{tag}__pairs'] = pairs
        out[f'{tag}__r0'] = r0.numpy()
        out[f'{tag}__r1'] = r1.numpy()
        log(f'match {tag}: n={n} m={m} D={c} matches={pairs.shape[0]} restatement equal')
    np.savez_compressed(os.path.join(GOLD, 'ref_match.npz'), **out)


def gen_eval(ref):
    out = {}
    cfg = synth.CONFIGS['cfg1']
    for tag, kind, h, w, seed in [('rep_480x640', 'uniform', 480, 640, synth.pair_seed(1, 1)),
                                  ('rep_small', 'uniform', 120, 160, 91)]:
        hm = synth.homography(seed + 7)
        s0 = synth.score_map(kind, h, w, seed)
        s1 = synth.warp_map(s0, hm, 'nearest')
        w01, w10 = synth.warp_params(hm, h, w)
        k0 = ref.extracter.detection(s0.clone(), cfg.extractor_params)
        k1 = ref.extracter.detection(s1.clone(), cfg.extractor_params)
        a, b, ids, ids_out = ref.projection.warp(k0, w01)
        ma, mb, mids, mids_out = ref_ops.warp(k0.numpy(), w01)
        assert np.array_equal(ids.numpy(), mids) and np.array_equal(ids_out.numpy(), mids_out), tag
        assert np.allclose(a.numpy(), ma, rtol=1e-5, atol=1e-6) and np.allclose(b.numpy(), mb, rtol=1e-5, atol=1e-6), tag
        res = ref.repeatability.val_key_points(k0, k1, w01, w10, th=3)
        mine = ref_ops.val_key_points(k0.numpy(), k1.numpy(), w01, w10, th=3)
        gt_ref = int(round(float(res['repeatability']) * res['num_feat']))
        log(f'val_key_points {tag}: n0={k0.shape[0]} n1={k1.shape[0]} num_feat={res["num_feat"]} '
            f'repeatability={float(res["repeatability"]):.6f} (restated {mine["repeatability"]:.6f}) '
            f'mean_error={float(res["mean_error"]):.6f} (restated {mine["mean_error"]:.6f})')
        # the reference's own mutual pairs (val_key_points returns only counts): its functions, its arithmetic
        rk0c, rk01c, _, _ = ref.projection.warp(k0, w01)
        rk1c, rk10c, _, _ = ref.projection.warp(k1, w10)
        rdm = (ref.repeatability.compute_keypoints_distance(rk0c, rk10c) +
               ref.repeatability.compute_keypoints_distance(rk1c, rk01c).t()) / 2
        for q in range(min(rdm.shape)):
            rdm[q, q] = 99999                                       # repeatability.py:72-73
        rii, rjj = ref.repeatability.mutual_argmin(rdm)
        ref_pairs = np.stack([rii.numpy(), rjj.numpy()], axis=1).astype(np.int64)
        # the restatement may differ from the reference only on pairs at a 2^-7 bucket edge (1-ulp warp coordinates)
        from oracle.compare import explain_repeat_pair_diffs
        dm = ((ref_ops.keypoint_distance(ma, ref_ops.warp(k1.numpy(), w10)[1]) +
               ref_ops.keypoint_distance(ref_ops.warp(k1.numpy(), w10)[0], mb).T) / np.float32(2)).astype(np.float32)
        dm[np.arange(min(dm.shape)), np.arange(min(dm.shape))] = np.float32(99999)
        n_edge = explain_repeat_pair_diffs(map(tuple, ref_pairs.tolist()), map(tuple, mine['pairs'].tolist()), dm)
        log(f'val_key_points {tag}: {ref_pairs.shape[0]} mutual pairs in the reference, {n_edge} differ from the restatement (bucket-edge cases)')
        assert abs(gt_ref - mine['gt_num']) <= n_edge, tag
        out[f'{tag}__pairs'] = ref_pairs
        # errors are normalised distances x resize(512): 1e-5 in normalised units = 5.12e-3 here
        assert np.allclose(res['errors'].numpy(), mine['errors'], rtol=1e-5, atol=1e-5 * 512), tag
        out[f'{tag}__seed'] = np.array(seed)
        out[f'{tag}__hw'] = np.array([h, w])
        out[f'{tag}__H'] = hm.numpy()
        out[f'{tag}__s1sha'] = np.array(sha(s1.numpy()))
        out[f'{tag}__k0'] = k0.numpy()
        out[f'{tag}__k1'] = k1.numpy()
        out[f'{tag}__warp_valid'] = a.numpy()
        out[f'{tag}__warp_proj'] = b.numpy()
        out[f'{tag}__ids'] = ids.numpy()
        out[f'{tag}__ids_out'] = ids_out.numpy()
        out[f'{tag}__num_feat'] = np.array(res['num_feat'])
        out[f'{tag}__gt_num'] = np.array(gt_ref)
        out[f'{tag}__repeatability'] = np.array(float(res['repeatability']))
        out[f'{tag}__mean_error'] = np.array(float(res['mean_error']))
        out[f'{tag}__errors'] = res['errors'].numpy()
    # MHA end to end on a reduced SuperPoint-shaped pair (tasks/MHA.py:11-72)
    h, w, c = 240, 320, 64
    seed = 4242
    hm = synth.homography(seed + 7)
    s0 = synth.score_map('uniform', h, w, seed)
    s1 = synth.warp_map(s0, hm, 'nearest')
    d0 = synth.desc_map(c, h // 8, w // 8, seed + 13, True)
    g = torch.Generator().manual_seed(seed + 17)
    d1 = synth.warp_map(d0, synth.rescale_homography(hm, 8), 'bilinear') + 0.05 * torch.randn(d0.shape, generator=g)
    w01, w10 = synth.warp_params(hm, h, w)
    params = {'extractor_params': dict(nms_dist=6, threshold=0, border_dist=8, top_k=300, min_score=0.0),
              'matcher_params': {'brute_force_params': {'metric': 'euclidean', 'max_distance': 5, 'cross_check': True}},
              'MHA_params': {'th': [3, 5, 7]}}
    img = torch.zeros(1, 3, h, w)
    flags = ref.mha.mha(0, img, s0, d0, img, s1, d1, w01, w10, params)
    myflags, pairs = ref_ops.mha_pair(s0, d0.numpy(), s1, d1.numpy(), w01, w10, params, (h, w))
    assert list(flags) == list(myflags), (flags, myflags)
    log(f'mha reduced pair: flags={flags} matches={0 if pairs is None else pairs.shape[0]} restatement equal')
    out['mha__seed'] = np.array(seed)
    out['mha__hwc'] = np.array([h, w, c])
    out['mha__d0'] = d0.numpy()
    out['mha__d1'] = d1.numpy()
    out['mha__s1sha'] = np.array(sha(s1.numpy()))
    out['mha__flags'] = np.array(flags)
    out['mha__pairs'] = pairs
    np.savez_compressed(os.path.join(GOLD, 'ref_eval.npz'), **out)


SE3_CASES = [('se3_240x320', 240, 320, 31, 600), ('se3_480x640', 480, 640, 32, 1000)]


def gen_se3(ref):
    """warp(mode='se3') (utils/projection.py:194-267) on synthetic two-view depth scenes."""
    out = {}
    for tag, h, w, seed, n in SE3_CASES:
        params = synth.se3_scene(h, w, seed)
        g = torch.Generator().manual_seed(seed + 1)
        kp = torch.rand(n, 3, generator=g)
        a, b, ids, ids_out = ref.projection.warp(kp, params)
        ma, mb, mids, mids_out = ref_ops.warp(kp.numpy(), params)
        assert np.array_equal(ids.numpy(), mids) and np.array_equal(ids_out.numpy(), mids_out), tag
        assert np.allclose(a.numpy(), ma, rtol=1e-5, atol=1e-6) and np.allclose(b.numpy(), mb, rtol=1e-5, atol=1e-5), tag
        log(f'warp se3 {tag}: n={n} valid={ids.shape[0]} out={ids_out.shape[0]} '
            f'(no-depth {n - ids.shape[0] - ids_out.shape[0]}) restatement equal')
        out[f'{tag}__kp'] = kp.numpy()
        out[f'{tag}__valid'] = a.numpy()
        out[f'{tag}__proj'] = b.numpy()
        out[f'{tag}__ids'] = ids.numpy()
        out[f'{tag}__ids_out'] = ids_out.numpy()
        out[f'{tag}__depth_sha'] = np.array(sha(params['depth0'].numpy()) + sha(params['depth1'].numpy()))
    np.savez_compressed(os.path.join(GOLD, 'ref_se3.npz'), **out)


LG_CASES = [  # tag, score kind, H, W, seed, desc C, s (cell size of the descriptor map)
    ('lg_uniform', 'uniform', 120, 160, 61, 64, 8),
    ('lg_alike', 'alike', 96, 128, 62, 32, 8),
    ('lg_ties', 'ties', 64, 80, 63, 16, 1),
    ('lg_relu_few', 'relu', 72, 88, 64, 16, 2),
    ('lg_topk', 'uniform', 480, 640, 65, 32, 8),      # > 1000 candidates: torch.topk binds
]


def gen_lightglue():
    """simple_nms / extract of the LightGlue-style extractor (models/lightglue.py:904-979)."""
    lg = _refimport.load_lightglue_extract()
    out = {}
    for tag, kind, h, w, seed, c, s in LG_CASES:
        sc = synth.score_map(kind, h, w, seed)                       # [1,1,H,W]
        g = torch.Generator().manual_seed(seed + 7)
        dm = torch.randn(1, c, h // s, w // s, generator=g)
        for r in (0, 2, 5):
            want = lg.simple_nms(sc[0], r).numpy()
            mine = ref_ops.simple_nms(sc[0].numpy(), r)
            assert np.array_equal(want, mine), (tag, r)
            if h * w <= 40000 or r == 5:
                out[f'{tag}__nms{r}'] = want[0]
        feats = lg.extract(lambda img: (sc.clone(), dm.clone()), torch.zeros(1, 3, h, w), s)
        kp, val, desc, raster = ref_ops.lightglue_extract(sc.numpy(), dm.numpy(), s)
        rkp, rval, rdesc = feats['keypoints'][0].numpy(), feats['keypoint_scores'][0].numpy(), feats['descriptors'][0].numpy()
        assert rkp.shape == kp.shape and rdesc.shape == desc.shape, tag
        # torch.topk orders ties arbitrarily: compare as sets of (x, y, score), verbatim where scores are unique
        rr = (rkp[:, 1].astype(np.int64) * w + rkp[:, 0].astype(np.int64))
        assert np.array_equal(np.sort(rr), np.sort(raster)), tag
        uniq, cnt = np.unique(rval, return_counts=True)
        single = np.isin(rval, uniq[cnt == 1])
        assert np.array_equal(rkp[single], kp[np.isin(val, uniq[cnt == 1])]), tag
        o_r, o_m = np.argsort(rr), np.argsort(raster)
        assert np.array_equal(rval[o_r], val[o_m]), tag
        assert np.allclose(rdesc[o_r], desc[o_m], rtol=1e-5, atol=1e-6), tag
        log(f'lightglue extract {tag}: {h}x{w} C={c} s={s} -> n={kp.shape[0]} '
            f'(ties {int((~single).sum())}) simple_nms r=0/2/5 bit-equal, extract restatement equal')
        out[f'{tag}__kp'] = rkp
        out[f'{tag}__val'] = rval
        out[f'{tag}__desc'] = rdesc
        out[f'{tag}__dm'] = dm.numpy()
    np.savez_compressed(os.path.join(GOLD, 'ref_lightglue.npz'), **out)


LK_CASES = [  # tag, C, H, W, win_size, levels, iterations, distance, seed, n
    ('lk_cfg', 3, 96, 128, 21, 3, 40, 10, 1, 150),        # the shipped YAML settings (config_fund.yaml:72-77)
    ('lk_default', 3, 64, 80, 3, 1, 40, 3, 2, 100),       # the class defaults (matcher.py:9-16)
    ('lk_gray', 1, 72, 96, 9, 2, 10, 5, 3, 100),
    ('lk_onestep', 3, 96, 128, 21, 3, 1, 10, 4, 150),     # a single step per level: no convergence to hide behind
]


def gen_lk(ref):
    """optical_flow_tensor (utils/matcher.py:188-203) with the random start replayed from the same torch seed."""
    out = {}
    for tag, c, h, w, win, levels, iters, dist, seed, n in LK_CASES:
        img0, img1 = synth.lk_scene(c, h, w, seed)
        g = torch.Generator().manual_seed(seed + 50)
        pts = torch.rand(n, 2, generator=g)
        pts[:4] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.02, 0.97], [0.5, 0.001]])      # corners / borders
        params = {'distance': dist, 'win_size': win, 'levels': levels, 'interation': iters, 'gray': c == 1}
        torch.manual_seed(seed + 99)
        t0 = time.time()
        want = ref.matcher.optical_flow_tensor(pts, pts, img0, img1, params)[0].numpy()
        dt = time.time() - t0
        torch.manual_seed(seed + 99)
        angle = (torch.randn(n) * 6.28).numpy()                                           # matcher.py:55
        init = ref_ops.lk_init_points(pts.numpy(), h, w, dist, angle)
        p0 = pts.numpy() * np.array([w - 1, h - 1], np.float32)
        mine = ref_ops.lk_track(img0[0].numpy(), img1[0].numpy(), p0, init, win, levels, iters)
        err = np.abs(mine - want).max()
        assert err < 1e-3, (tag, err)
        log(f'optical_flow_tensor {tag}: C={c} {h}x{w} win={win} levels={levels} it={iters} n={n} '
            f'reference_cpu_s={dt:.2f} max |restated - reference| = {err:.2e} px')
        out[f'{tag}__img0'] = img0.numpy()
        out[f'{tag}__img1'] = img1.numpy()
        out[f'{tag}__pts'] = pts.numpy()
        out[f'{tag}__init'] = init
        out[f'{tag}__out'] = want
    np.savez_compressed(os.path.join(GOLD, 'ref_lk.npz'), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = _refimport.load()
    log(f'reference root: {_refimport.REFERENCE_ROOT}; torch {torch.__version__}; numpy {np.__version__}')
    gen_nms(ref)
    gen_detect(ref)
    gen_match(ref)
    gen_eval(ref)
    gen_se3(ref)
    gen_lightglue()
    gen_lk(ref)
    with open(os.path.join(ROOT, 'oracle', 'REFCHECK.log'), 'w') as f:
        f.write('\n'.join(LOG) + '\n')


if __name__ == '__main__':
    main()
