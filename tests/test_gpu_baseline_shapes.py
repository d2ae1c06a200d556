"""GPU parity on BASELINE.json's own shapes: the matcher at cfg3's 4096 x 4096 x 64, the full-resolution gather
sampler at cfg4 / cfg5 map sizes, and the task-faithful pipeline of every config (what bench.py times) against the
CPU oracle on two pairs at full size -- exact keypoint rows, exact pair sets (modulo the 1e-5 near-tie rule of
north_star), exact counts, and the run accumulators as models/model_interface.py:124-137 computes them."""
import math

import numpy as np
import pytest
import torch
from scipy.spatial.distance import cdist

from helpers import check_detection_rows, exact_pairs_or_near_tie
from keypoint_bench_b200 import synth
from oracle import ref_ops

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def ops():
    from keypoint_bench_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------------ matcher, cfg3 shape

@pytest.fixture(scope='module')
def cfg3_descriptors():
    gen = torch.Generator().manual_seed(4096)
    n = m = 4096
    a = torch.nn.functional.normalize(torch.randn(n, 64, generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(m, 64, generator=gen), dim=1)
    b[:2500] = a[torch.randperm(n, generator=gen)[:2500]] + 0.04 * torch.randn(2500, 64, generator=gen)
    b[7] = b[3]                                              # duplicated rows: first of ties (np.argmin)
    a[11] = a[5]
    return a, b, cdist(a.numpy(), b.numpy())


@pytest.mark.parametrize('algo', [0, 1])
@pytest.mark.parametrize('maxd,cc', [(5.0, True), (math.inf, True), (5.0, False), (math.inf, False), (0.45, True)])
def test_matcher_cfg3_shape_4096x4096x64(cfg3_descriptors, algo, maxd, cc):
    """utils/matcher.py:227-233 at XFeat's top_k = 4096, D = 64 (sixteen column tiles per row tile, the shape where
    the tensor-core kernel is epilogue-bound)."""
    a, b, D = cfg3_descriptors
    for want_dist in (True, False):
        pairs, dist, count = ops().match_batched(a[None].to(DEV), b[None].to(DEV), None, None, maxd, cc, algo=algo,
                                                 want_dist=want_dist)
        got = pairs[0, :int(count[0])].cpu().numpy().astype(np.int64)
        assert got.shape[0] > 1500
        exact_pairs_or_near_tie(got, a.numpy(), b.numpy(), maxd, cc, dist=D)
        if want_dist:
            assert np.allclose(dist[0, :got.shape[0]].cpu().numpy(), D[got[:, 0], got[:, 1]], rtol=1e-12, atol=1e-12)


def test_matcher_cfg3_shape_batched_ragged():
    """Two pairs of the cfg3 batch with ragged keypoint counts (the batched call bench.py makes)."""
    gen = torch.Generator().manual_seed(77)
    a = torch.nn.functional.normalize(torch.randn(2, 4096, 64, generator=gen), dim=2)
    b = torch.nn.functional.normalize(torch.randn(2, 4096, 64, generator=gen), dim=2)
    b[:, :3000] = a[:, 500:3500] + 0.05 * torch.randn(2, 3000, 64, generator=gen)
    n0 = torch.tensor([4096, 3777], dtype=torch.int32)
    n1 = torch.tensor([3901, 4096], dtype=torch.int32)
    pairs, _, count = ops().match_batched(a.to(DEV), b.to(DEV), n0.to(DEV), n1.to(DEV), 5.0, True, want_dist=False)
    for i in range(2):
        got = pairs[i, :int(count[i])].cpu().numpy().astype(np.int64)
        exact_pairs_or_near_tie(got, a[i, :n0[i]].numpy(), b[i, :n1[i]].numpy(), 5.0, True)


# ------------------------------------------------------------------------------------------------ gather sampler, cfg4 / cfg5 shapes

@pytest.mark.parametrize('c,h,w,n,normalized', [(128, 1024, 1024, 2048, True), (64, 376, 1241, 1000, False)])
def test_gather_sampler_full_resolution_maps(c, h, w, n, normalized):
    """utils/matcher.py:221-226 on full-resolution descriptor maps (DISK [1,128,1024,1024] x 2048 keypoints, ALIKE
    [1,64,376,1241] x 1000): the NCHW gather path of kb_sample_desc, with the keypoints detection() would hand over
    (score order, i.e. spatially random) plus points on the corners / outside, and a ragged second map."""
    gen = torch.Generator(device=DEV).manual_seed(c + h)
    d = synth.desc_map(c, h, w, 31, normalized, DEV)
    d = torch.cat([d, d.flip(3) * 0.5])
    params = dict(nms_dist=6, threshold=0.0, border_dist=8, top_k=n, min_score=0.0)
    xyp, count, _, _ = ops().detect_batched(torch.rand(2, 1, h, w, generator=gen, device=DEV), params)
    assert int(count.min()) == n
    pts = xyp.clone()
    pts[0, 0, :2] = torch.tensor([0.0, 0.0]); pts[0, 1, :2] = torch.tensor([1.0, 1.0])
    pts[0, 2, :2] = torch.tensor([1.0, 0.0]); pts[0, 3, :2] = torch.tensor([-0.2, 0.4]); pts[0, 4, :2] = torch.tensor([0.3, 1.3])
    cnt = torch.tensor([n, n - 313], dtype=torch.int32, device=DEV)
    out = ops().sample_batched(d, pts, cnt).cpu().numpy()
    dc, pc = d.cpu().numpy(), pts.cpu().numpy()
    for b in range(2):
        k = int(cnt[b])
        want = ref_ops.sample_brute_force(dc[b], pc[b, :k])
        assert np.allclose(out[b, :k], want, rtol=1e-5, atol=1e-5), (b, float(np.abs(out[b, :k] - want).max()))
    assert not out[1, n - 313:].any()                         # rows beyond the count stay zero
    # lightglue mode (normalised rows) on the same map: models/lightglue.py:24-41 with s = 1 (DISK)
    kp_px = (pts[:1, :, :2] * torch.tensor([w, h], device=DEV)).contiguous()
    outn = ops().sample_batched(d[:1], kp_px, None, normalize=True, coord_mode=1, s=1)[0].cpu().numpy()
    wantn = ref_ops.sample_lightglue(dc[0], kp_px[0].cpu().numpy(), 1)
    assert np.allclose(outn, wantn, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ task pipelines, full shape

def _np(t):
    return t.detach().cpu().numpy()


def _oracle_keypoints(score_cpu, cfg):
    return [ref_ops.detection(score_cpu[i:i + 1], cfg.extractor_params, nms='greedy') for i in range(score_cpu.shape[0])]


def _check_kpts(res, want, i):
    n = int(res['n_kpts'][i])
    check_detection_rows(_np(res['kpts'][i, :n]), want[i][0], _np(res['raster'][i, :n]).astype(np.int64), want[i][1])
    assert int(res['path'][i]) == 1                           # the sparse exact path certified the map


@pytest.mark.parametrize('name', ['cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg5'])
def test_task_pipeline_full_shape_against_oracle(name):
    from keypoint_bench_b200 import parallel, pipeline
    cfg = synth.CONFIGS[name]
    P = 2
    if cfg.task == 'stream':
        frames = synth.make_frames(cfg, 99, 4, P + 1, 40, DEV)            # frames 4..6 of a 40-frame sequence
        res, acc = pipeline.run_task(frames, cfg)
        score, desc = frames.score.cpu(), _np(frames.desc)
        want = _oracle_keypoints(score, cfg)
        for i in range(P + 1):
            _check_kpts(res, want, i)
        total = 0
        for f in range(P):
            k0, k1 = want[f][0], want[f + 1][0]
            got = _np(res['matches'][f, :int(res['n_matches'][f])]).astype(np.int64)
            d0, d1 = ref_ops.sample_brute_force(desc[f], k0), ref_ops.sample_brute_force(desc[f + 1], k1)
            assert exact_pairs_or_near_tie(got, d0, d1, cfg.max_distance, cfg.cross_check) <= 2
            assert got.shape[0] > 100
            total += got.shape[0]
        assert acc.tolist() == [float(total), float(P)]
        return
    batch, hms = synth.make_batch(cfg, int(name[3:]), P, 0, DEV)
    res, acc = pipeline.run_task(batch, cfg)
    score = batch.score.cpu()
    want = _oracle_keypoints(score, cfg)
    for i in range(2 * P):
        _check_kpts(res, want, i)
    if cfg.task == 'repeatability':
        reps, errs, feats = [], [], []                         # the lists of models/model_interface.py:242-246
        for i in range(P):
            w01, w10 = synth.warp_params(hms[i], cfg.height, cfg.width)
            ora = ref_ops.val_key_points(want[i][0], want[P + i][0], w01, w10, th=3)
            st = _np(res['stats'][i])
            assert int(res['num_feat'][i]) == ora['num_feat']
            assert int(st[0]) == ora['gt_num'] and int(st[2]) == ora['pairs'].shape[0], (st, ora['gt_num'])
            assert ora['gt_num'] > 300
            assert abs(st[1] / st[0] - ora['mean_error']) < 1e-5
            a = int(res['n_cov'][i])
            assert np.array_equal(_np(res['errors'][i, :a]), ora['errors'])
            reps.append(ora['repeatability']); errs.append(ora['mean_error']); feats.append(ora['num_feat'])
        fin = parallel.finalize_repeatability(acc)
        e = np.asarray(errs)
        assert abs(fin['repeatability'] - np.mean(reps)) < 1e-12                 # model_interface.py:124-133
        assert abs(fin['rep_mean_err'] - np.mean(e[~np.isnan(e)])) < 1e-5
        assert abs(fin['num_feat'] - np.mean(feats)) < 1e-12 and fin['pairs'] == P
        return
    desc = _np(batch.desc)
    total = 0
    for i in range(P):
        k0, k1 = want[i][0], want[P + i][0]
        if cfg.task == 'mha':                                  # tasks/MHA.py:33-34: covisible keypoints, two columns
            w01, w10 = synth.warp_params(hms[i], cfg.height, cfg.width)
            k0, _, ids0, _ = ref_ops.warp(k0, w01)
            k1, _, ids1, _ = ref_ops.warp(k1, w10)
            for j, (kc, ids) in ((i, (k0, ids0)), (P + i, (k1, ids1))):
                n = int(res['n_cov'][j])
                assert n == kc.shape[0]
                assert np.array_equal(_np(res['kcov'][j, :n]), kc) and np.array_equal(_np(res['cov_ids'][j, :n]), ids)
        else:
            assert 'kcov' not in res                           # tasks/AUC.py:115-120 matches every keypoint
        got = _np(res['matches'][i, :int(res['n_matches'][i])]).astype(np.int64)
        d0, d1 = ref_ops.sample_brute_force(desc[i], k0), ref_ops.sample_brute_force(desc[P + i], k1)
        n0 = k0.shape[0]
        assert np.allclose(_np(res['desc'][i, :n0]), d0, rtol=1e-5, atol=1e-5)
        assert exact_pairs_or_near_tie(got, d0, d1, cfg.max_distance, cfg.cross_check) <= 2
        assert got.shape[0] > 300
        total += got.shape[0]
    assert acc.tolist() == [float(total), float(P)]
