"""GPU parity: the sm_100a kernels (through the C ABI) against the CPU oracle and the fixtures minted
from the reference.  Bit-exact for keypoint sets / indices / match pairs, 1e-5 for float outputs
(the tolerance BASELINE.json's north_star states)."""
import ast
import math

import numpy as np
import pytest
import torch

from helpers import explain_repeat_pair_diffs, oracle_dist_mutual
from keypoint_bench_b200 import synth
from oracle import ref_ops

pytestmark = pytest.mark.gpu

DEV = 'cuda'
LK_TOL = 2e-5       # px (measured on B200: <= 1.53e-5 against the reference fixtures, 7.6e-6 against the oracle); tracked positions of the Lucas-Kanade matcher (see test_lk_tracker_matches_reference_fixtures)


def ops():
    from keypoint_bench_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------------ NMS

def test_fast_nms_matches_reference_fixtures(golden):
    g = golden('ref_nms.npz')
    cases = [ast.literal_eval(str(c)) for c in g['cases']]
    from keypoint_bench_b200.utils.extracter import fast_nms
    for i, (kind, h, w, seed, r, mv, mi) in enumerate(cases):
        s = torch.from_numpy(g[f'in_{i}'])[None, None].to(DEV)
        out = fast_nms(s, nms_dist=r, max_iter=mi, min_value=mv)
        assert out.shape == s.shape
        assert np.array_equal(out.cpu().numpy()[0, 0], g[f'out_{i}']), f'case {i} {kind} r={r} mv={mv} mi={mi}'
    s = torch.rand(1, 1, 8, 8, device=DEV)
    assert fast_nms(s, nms_dist=0) is s                      # extracter.py:40-41


@pytest.mark.parametrize('kind,h,w,r', [('uniform', 480, 640, 6), ('uniform', 376, 1241, 6), ('alike', 240, 320, 6),
                                        ('ties', 200, 333, 4), ('mixed', 130, 170, 5), ('negative', 90, 110, 3),
                                        ('ramp', 64, 640, 6), ('relu', 150, 150, 8)])
def test_fast_nms_matches_oracle_rounds(kind, h, w, r):
    s = synth.score_map(kind, h, w, 100 + r)
    want, rounds = ref_ops.nms_rounds_separable(s.numpy()[0, 0][None], r, return_rounds=True)
    out, got_rounds = ops().fast_nms_batched(s.to(DEV), r, return_rounds=True)
    assert np.array_equal(out.cpu().numpy()[0, 0], want[0])
    assert int(got_rounds.item()) == rounds


def test_fast_nms_batch_is_joint_like_the_reference():
    # the reference counts maxima over the whole batch (extracter.py:73); mixed-sign maps make the
    # stopping round observable, so the batch must be processed jointly
    s = torch.cat([synth.score_map('mixed', 60, 80, 1), synth.score_map('negative', 60, 80, 2),
                   synth.score_map('uniform', 60, 80, 3)], dim=0)
    want = ref_ops.nms_rounds_separable(s.numpy()[:, 0], 4)
    out = ops().fast_nms_batched(s.to(DEV), 4)
    assert np.array_equal(out.cpu().numpy()[:, 0], want)


# ------------------------------------------------------------------------------------------------ detection

# round 1 of the sparse path has three kernels with identical output (kb_round1_packed.cu for large batches, the tiled
# round1_kernel otherwise, kb_round1_stream.cu as a tested alternative): phases bit 3 / 4 / 5 force one of them
# (include/kb_b200.h)
ROUND1 = {'tiled': 7 | 8, 'stream': 7 | 16, 'packed': 7 | 32}


@pytest.fixture(params=['tiled', 'stream', 'packed'])
def round1(request):
    return ROUND1[request.param]


def _check_detection(pts, want, raster=None, want_raster=None):
    assert pts.shape == want.shape
    uniq, cnt = np.unique(want[:, 2], return_counts=True)
    single = uniq[cnt == 1]
    tf = np.isin(want[:, 2], single) & np.isin(pts[:, 2], single)
    assert np.array_equal(pts[tf], want[tf])                  # verbatim incl. order where scores are unique
    assert np.array_equal(np.sort(pts[:, 2]), np.sort(want[:, 2]))
    if want_raster is not None:
        assert np.array_equal(raster, want_raster)             # canonical order (score desc, raster asc)


def test_detection_matches_reference_fixtures(golden):
    from keypoint_bench_b200.utils.extracter import detection
    from oracle.make_golden import DETECT_CASES
    g = golden('ref_detect.npz')
    for tag, kind, h, w, seed, params in DETECT_CASES:
        s = synth.score_map(kind, h, w, seed).to(DEV)
        pts = detection(s, dict(params)).cpu().numpy()
        _check_detection(pts, g[f'{tag}__pts'])
        xyp, count, raster, _ = ops().detect_batched(s, dict(params))
        n = int(count[0])
        assert np.array_equal(raster[0, :n].cpu().numpy().astype(np.int64), g[f'{tag}__raster']), tag


def test_detection_1024_and_batched_against_oracle(round1):
    cfg = synth.CONFIGS['cfg4']
    s = synth.score_map('uniform', 1024, 1024, 9)
    want, want_r = ref_ops.detection(s, cfg.extractor_params, nms='greedy')
    xyp, count, raster, _ = ops().detect_batched(s.to(DEV), cfg.extractor_params, phases=round1)
    n = int(count[0])
    _check_detection(xyp[0, :n].cpu().numpy(), want, raster[0, :n].cpu().numpy().astype(np.int64), want_r)
    # a batch of different maps is processed independently per map
    maps = [synth.score_map(k, 120, 160, 40 + i) for i, k in enumerate(['uniform', 'ties', 'alike', 'relu', 'uniform'])]
    params = dict(nms_dist=4, threshold=0.0, border_dist=8, top_k=50, min_score=0.0)
    xyp, count, raster, _ = ops().detect_batched(torch.cat(maps, 0).to(DEV), params, phases=round1)
    for i, m in enumerate(maps):
        want, want_r = ref_ops.detection(m, params)
        n = int(count[i])
        assert n == want.shape[0]
        assert np.array_equal(raster[i, :n].cpu().numpy().astype(np.int64), want_r)
        assert np.array_equal(xyp[i, :n].cpu().numpy(), want)


def test_detection_accepts_any_top_k_and_batch_like_the_reference():
    """utils/extracter.py:193-221 takes any top_k (0 -> no rows; one the NMS can never exceed -> raster order), returns
    batch item 0 only, and runs fast_nms on the whole batch with one stopping rule (observable with negative scores)."""
    from keypoint_bench_b200.utils.extracter import detection
    s = synth.score_map('uniform', 90, 120, 17)
    base = dict(nms_dist=4, threshold=0.0, border_dist=8, min_score=0.0)
    assert detection(s.to(DEV), dict(base, top_k=0)).shape == (0, 3)
    for top_k in (10 ** 6, 90 * 120):
        want, _ = ref_ops.detection(s, dict(base, top_k=top_k))
        got = detection(s.to(DEV), dict(base, top_k=top_k)).cpu().numpy()
        assert np.array_equal(got, want)
    want, _ = ref_ops.detection(s, dict(base, nms_dist=0, top_k=10 ** 6))
    assert np.array_equal(detection(s.clone().to(DEV), dict(base, nms_dist=0, top_k=10 ** 6)).cpu().numpy(), want)
    with pytest.raises(Exception):
        detection(torch.rand(1, 1, 700, 900, device=DEV), dict(base, nms_dist=1, top_k=9000))     # 8192 < top_k < keep bound
    # batch of two maps with negative scores: the joint stopping rule decides item 0's rounds
    # (seeds chosen so that item 0 processed alone stops at a DIFFERENT round than inside the batch: 8 pixels differ)
    bt = torch.cat([synth.score_map('mixed', 40, 56, 2), synth.score_map('negative', 40, 56, 102)], dim=0)
    params = dict(base, top_k=50, threshold=-2.0)
    kept = ref_ops.nms_rounds_separable(bt.numpy()[:, 0], 4)[0]
    assert not np.array_equal(kept, ref_ops.nms_rounds_separable(bt.numpy()[:1, 0], 4)[0])
    pts, raster = ref_ops.positions_with_prob(ref_ops.clear_border(kept, 8)[None], -2.0)
    sel = ref_ops.canonical_order(pts[:, 2], raster)[:50]
    got = detection(bt.to(DEV), params).cpu().numpy()
    assert np.array_equal(got, pts[sel])
    # non-negative batch: item 0 alone
    b2 = torch.cat([s, synth.score_map('uniform', 90, 120, 18)])
    want, _ = ref_ops.detection(s, dict(base, top_k=100))
    assert np.array_equal(detection(b2.to(DEV), dict(base, top_k=100)).cpu().numpy(), want)
    # more maps than one kb_detect call takes
    many = torch.rand(2100, 1, 24, 24, device=DEV)
    xyp, count, raster, path = ops().detect_batched(many, dict(base, border_dist=2, top_k=5))
    assert xyp.shape == (2100, 5, 3) and int(count.min()) >= 1
    w5, _ = ref_ops.detection(many[2099:].cpu(), dict(base, border_dist=2, top_k=5))
    assert np.array_equal(xyp[2099, :int(count[2099])].cpu().numpy(), w5)


def test_positions_and_border_dropins():
    from keypoint_bench_b200.utils.extracter import prob_map_to_positions_with_prob, remove_border_points
    s = synth.score_map('uniform', 50, 70, 5)
    want, _ = ref_ops.positions_with_prob(s.numpy(), 0.7)
    got = prob_map_to_positions_with_prob(s.to(DEV), 0.7).cpu().numpy()
    assert np.array_equal(got, want)
    t = s.clone().to(DEV)
    assert remove_border_points(t, 5) is t
    assert np.array_equal(t.cpu().numpy(), ref_ops.clear_border(s.numpy(), 5))
    assert prob_map_to_positions_with_prob(torch.zeros(1, 1, 9, 9, device=DEV)).shape == (0, 3)


# ------------------------------------------------------------------------------------------------ sampling

def test_sampling_matches_reference_fixture_and_oracle(golden):
    from keypoint_bench_b200.utils.matcher import sample_descriptors_at
    g = golden('ref_match.npz')
    d0 = torch.from_numpy(g['small32__desc0']).to(DEV)
    p0 = torch.from_numpy(g['small32__p0']).to(DEV)
    got = sample_descriptors_at(d0, p0).cpu().numpy()
    assert np.allclose(got, g['small32__s0'], rtol=1e-5, atol=1e-5)
    # points on / outside the border, non-multiple-of-32 channel count, batch > 1 with ragged counts
    gen = torch.Generator().manual_seed(3)
    d = torch.randn(2, 70, 13, 17, generator=gen)
    p = torch.rand(2, 90, 3, generator=gen) * 1.2 - 0.1
    p[0, 0, :2] = torch.tensor([0.0, 0.0]); p[0, 1, :2] = torch.tensor([1.0, 1.0]); p[0, 2, :2] = torch.tensor([1.0, 0.0])
    cnt = torch.tensor([90, 37], dtype=torch.int32)
    out = ops().sample_batched(d.to(DEV), p.to(DEV), cnt.to(DEV)).cpu().numpy()
    for b in range(2):
        n = int(cnt[b])
        want = torch.nn.functional.grid_sample(d[b:b + 1], ((p[b, :n, :2] - 0.5) * 2)[None, None],
                                               align_corners=True)[0, :, 0].T.numpy()
        assert np.allclose(out[b, :n], want, rtol=1e-5, atol=1e-5)
        assert np.allclose(out[b, :n], ref_ops.sample_brute_force(d[b].numpy(), p[b, :n].numpy()), rtol=1e-5, atol=1e-5)


def test_lightglue_sampling_mode():
    gen = torch.Generator().manual_seed(8)
    d = torch.randn(1, 256, 30, 40, generator=gen)
    kp = torch.rand(1, 200, 2, generator=gen) * torch.tensor([320.0, 240.0])
    out = ops().sample_batched(d.to(DEV), kp.to(DEV), None, normalize=True, coord_mode=1, s=8)[0].cpu().numpy()
    want = ref_ops.sample_lightglue(d.numpy(), kp[0].numpy(), 8)
    assert np.allclose(out, want, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ matching

def _exact_pairs_or_near_tie(got, d0, d1, maxd, cc):
    """Pairs must equal the float64 oracle except rows/cols whose best and second-best distances are
    within 1e-5 relative (north_star)."""
    want = ref_ops.match_descriptors(d0, d1, metric='euclidean', max_distance=maxd, cross_check=cc)
    if np.array_equal(got, want):
        return
    from scipy.spatial.distance import cdist
    D = cdist(d0, d1)
    s = np.sort(D, axis=1)
    t = np.sort(D, axis=0)
    amb_rows = set(np.flatnonzero((s[:, 1] - s[:, 0]) <= 1e-5 * s[:, 1]).tolist())
    amb_cols = set(np.flatnonzero((t[1] - t[0]) <= 1e-5 * t[1]).tolist())
    diff = set(map(tuple, got.tolist())) ^ set(map(tuple, want.tolist()))
    for i, j in diff:
        assert i in amb_rows or j in amb_cols or abs(D[i, j] - maxd) <= 1e-5 * maxd, (i, j)


@pytest.mark.parametrize('algo', [0, 1])
def test_matcher_matches_reference_fixtures(golden, algo):
    """ref_match.npz was minted by the reference's own brute_force_matcher (utils/matcher.py:206-234) -- but the matching
    inside it is skimage.feature.match_descriptors, a third-party function that is neither in the reference tree nor
    installable here, so the fixture ran with the oracle's restatement of its published algorithm injected
    (oracle/_refimport.py): for the matching half this fixture is UNPINNED UPSTREAM; the sampling half and
    the row gather around it are the reference's."""
    g = golden('ref_match.npz')
    for tag in ('small32', 'sp256', 'nocross', 'tight'):
        maxd, cc = g[f'{tag}__maxd_cc']
        s0, s1 = g[f'{tag}__s0'], g[f'{tag}__s1']
        pairs, dist, count = ops().match_batched(torch.from_numpy(s0)[None].to(DEV), torch.from_numpy(s1)[None].to(DEV),
                                                 None, None, float(maxd), bool(cc), algo=algo)
        k = int(count[0])
        got = pairs[0, :k].cpu().numpy().astype(np.int64)
        assert np.array_equal(got, g[f'{tag}__pairs']), tag
        from scipy.spatial.distance import cdist
        D = cdist(s0, s1)
        assert np.allclose(dist[0, :k].cpu().numpy(), D[got[:, 0], got[:, 1]], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('algo', [0, 1])
@pytest.mark.parametrize('n,m,dim', [(1000, 1000, 256), (1000, 977, 64), (2048, 2048, 128), (1, 5, 32), (130, 1, 64),
                                     (257, 511, 48)])
def test_matcher_against_oracle_sizes(algo, n, m, dim):
    gen = torch.Generator().manual_seed(n * 7 + m)
    a = torch.nn.functional.normalize(torch.randn(n, dim, generator=gen), dim=1)
    b = torch.nn.functional.normalize(torch.randn(m, dim, generator=gen), dim=1)
    k = min(n, m) // 2
    b[:k] = a[:k] + 0.05 * torch.randn(k, dim, generator=gen)
    for maxd, cc in ((5.0, True), (0.9, True), (math.inf, False)):
        pairs, dist, count = ops().match_batched(a[None].to(DEV), b[None].to(DEV), None, None, maxd, cc, algo=algo)
        got = pairs[0, :int(count[0])].cpu().numpy().astype(np.int64)
        _exact_pairs_or_near_tie(got, a.numpy(), b.numpy(), maxd, cc)


@pytest.mark.parametrize('algo', [0, 1])
@pytest.mark.parametrize('dim,scale', [(256, 1.0), (64, 1.0), (128, 37.5)])
def test_matcher_without_distances_gives_same_pairs(algo, dim, scale):
    """``want_dist=False`` (what the pipeline uses: matcher.py:227-233 keeps only the index pairs) lets the tensor-core
    path decide ``d < max_distance`` from its certified score and fall back to float64 only inside the error band;
    the pairs must be identical, including for a max_distance that equals one pair's float64 distance exactly
    (the comparison is strict)."""
    gen = torch.Generator().manual_seed(dim)
    n, m = 900, 1000
    a = torch.nn.functional.normalize(torch.randn(2, n, dim, generator=gen), dim=2) * scale
    b = torch.nn.functional.normalize(torch.randn(2, m, dim, generator=gen), dim=2) * scale
    b[:, :600] = a[:, :600] + scale * torch.linspace(0.0, 0.12, 600)[None, :, None] * torch.randn(2, 600, dim, generator=gen)
    full, dist, cnt = ops().match_batched(a.to(DEV), b.to(DEV), None, None, math.inf, True, algo=algo)
    d_sorted = np.sort(dist[0, :int(cnt[0])].cpu().numpy())
    gates = [float(d_sorted[len(d_sorted) // 2]), float(np.nextafter(d_sorted[len(d_sorted) // 2], np.inf)),
             float(d_sorted[len(d_sorted) // 3]) * (1 + 1e-7), 0.9 * scale, 5.0 * scale, 1e-3 * scale]
    for maxd in gates:
        for cc in (True, False):
            p_ref, _, c_ref = ops().match_batched(a.to(DEV), b.to(DEV), None, None, maxd, cc, algo=algo)
            p_got, d_got, c_got = ops().match_batched(a.to(DEV), b.to(DEV), None, None, maxd, cc, algo=algo,
                                                      want_dist=False)
            assert d_got is None
            assert torch.equal(c_ref, c_got), (maxd, cc)
            for i in range(2):
                k = int(c_ref[i])
                assert torch.equal(p_ref[i, :k], p_got[i, :k]), (maxd, cc, i)
            if algo == 1:
                _exact_pairs_or_near_tie(p_got[0, :int(c_got[0])].cpu().numpy().astype(np.int64), a[0].numpy(),
                                         b[0].numpy(), maxd, cc)


@pytest.mark.parametrize('algo', [0, 1])
@pytest.mark.parametrize('seed', range(12))
def test_matcher_randomised_against_oracle(algo, seed):
    """Random shapes (descriptor lengths that are not multiples of 4 / 64, single rows, ragged counts), magnitudes,
    duplicated rows on either side and gates: pairs equal the float64 oracle (or differ only on 1e-5 near-ties)."""
    rng = np.random.default_rng(1000 + seed)
    b = int(rng.integers(1, 4))
    n, m = int(rng.integers(1, 700)), int(rng.integers(1, 700))
    dim = int(rng.choice([3, 17, 32, 64, 100, 128, 130, 255, 256]))
    scale = float(rng.choice([1.0, 1.0, 0.01, 40.0]))
    gen = torch.Generator().manual_seed(seed)
    a = torch.randn(b, n, dim, generator=gen) * scale
    d = torch.randn(b, m, dim, generator=gen) * scale
    k = min(n, m) // 2
    d[:, :k] = a[:, :k] + 0.05 * scale * torch.randn(b, k, dim, generator=gen)
    if m > 5:
        d[:, 4] = d[:, 1]                                    # duplicated database rows: first of ties wins
    if n > 9:
        a[:, 8] = a[:, 2]
    n0 = torch.tensor(rng.integers(1, n + 1, size=b), dtype=torch.int32)
    n1 = torch.tensor(rng.integers(1, m + 1, size=b), dtype=torch.int32)
    typical = float(np.sqrt(2 * dim)) * scale                # distance between unrelated rows
    maxd = float(rng.choice([math.inf, 5.0 * typical, 0.6 * typical, 0.1 * typical]))
    cc = bool(rng.integers(0, 2))
    want_dist = bool(rng.integers(0, 2))
    from keypoint_bench_b200 import _lib
    with ops().debug_knob(_lib.KB_KNOB_TC_BF16X3, seed % 3):       # operand split: automatic / bf16 x 3 / fp16 x 2
        pairs, dist, count = ops().match_batched(a.to(DEV), d.to(DEV), n0.to(DEV), n1.to(DEV), maxd, cc, algo=algo,
                                                 want_dist=want_dist)
    for i in range(b):
        ai, di = a[i, :n0[i]].numpy(), d[i, :n1[i]].numpy()
        got = pairs[i, :int(count[i])].cpu().numpy().astype(np.int64)
        _exact_pairs_or_near_tie(got, ai, di, maxd, cc)
        if want_dist and got.shape[0]:
            from scipy.spatial.distance import cdist
            D = cdist(ai, di)
            assert np.allclose(dist[i, :got.shape[0]].cpu().numpy(), D[got[:, 0], got[:, 1]], rtol=1e-12, atol=1e-12)


def test_matcher_cta_pair_kernel_gives_same_pairs():
    """kb_debug_knob(KB_KNOB_TC_CLUSTER, 2) selects the cta_group::2 variant of the tensor-core search (clusters of two CTAs, M = 256
    MMAs issued by the leader, TMA loads completing on the leader's barriers): same pairs, same top-3 records."""
    gen = torch.Generator().manual_seed(77)
    cases = [(3, 1000, 1000, 256), (2, 700, 1300, 64), (1, 129, 300, 128), (2, 2048, 2048, 128)]
    for b, n, m, dim in cases:
        a = torch.nn.functional.normalize(torch.randn(b, n, dim, generator=gen), dim=2)
        d = torch.nn.functional.normalize(torch.randn(b, m, dim, generator=gen), dim=2)
        k = min(n, m) // 2
        d[:, :k] = a[:, :k] + 0.05 * torch.randn(b, k, dim, generator=gen)
        n0 = torch.tensor([n - 37 * i for i in range(b)], dtype=torch.int32)
        n1 = torch.tensor([m - 91 * i for i in range(b)], dtype=torch.int32)
        out = {}
        from keypoint_bench_b200 import _lib
        for mode in ('1', '2'):
            with ops().debug_knob(_lib.KB_KNOB_TC_CLUSTER, int(mode)):
                p, _, c, ws = ops().match_batched(a.to(DEV), d.to(DEV), n0.to(DEV), n1.to(DEV), 0.9, True, algo=1,
                                                  return_ws=True)
            best = ops().match_tc_debug(ws, b, n, m, dim)['res0'][0][:int(n0[0])]
            out[mode] = (p.clone(), c.clone(), best.clone())
        assert torch.equal(out['1'][1], out['2'][1]), (b, n, m, dim)
        assert torch.allclose(out['1'][2], out['2'][2], rtol=1e-6, atol=1e-6)        # the per-row best scores themselves
        for i in range(b):
            kk = int(out['1'][1][i])
            assert torch.equal(out['1'][0][i, :kk], out['2'][0][i, :kk]), (b, n, m, dim, i)
        want = ref_ops.match_descriptors(a[0, :n0[0]].numpy(), d[0, :n1[0]].numpy(), max_distance=0.9, cross_check=True)
        _exact_pairs_or_near_tie(out['2'][0][0, :int(out['2'][1][0])].cpu().numpy().astype(np.int64), a[0, :n0[0]].numpy(),
                                 d[0, :n1[0]].numpy(), 0.9, True)
        assert want.shape[0] > 0


@pytest.mark.parametrize('b,n,m,dim,dups', [(3, 1000, 977, 256, 0), (2, 700, 1300, 64, 40), (2, 2048, 2048, 128, 5),
                                            (1, 4096, 4096, 64, 0), (2, 130, 90, 32, 20), (1, 33, 1, 16, 0)])
def test_matcher_one_pass_cross_check_equals_two_pass(b, n, m, dim, dups):
    """KB_KNOB_TC_ONE_PASS decides the cross-check of the tensor-core path from column-group maxima of ONE Gram pass
    (redux.sync in the epilogue, exact rescan of the columns that are too close to call) instead of the second Gram with
    rows and columns swapped (the default).  Same pairs -- also with duplicated rows, which tie exactly on a column --
    and the exact float64 oracle agrees."""
    from keypoint_bench_b200 import _lib
    gen = torch.Generator().manual_seed(b * 1000 + n + dim)
    a = torch.nn.functional.normalize(torch.randn(b, n, dim, generator=gen), dim=2)
    d = torch.nn.functional.normalize(torch.randn(b, m, dim, generator=gen), dim=2)
    k = min(n, m) // 2
    d[:, :k] = a[:, :k] + 0.05 * torch.randn(b, k, dim, generator=gen)
    for t in range(dups):                                       # row 2t+1 duplicates row 2t: both are column t's best
        a[:, 2 * t + 1] = a[:, 2 * t]
    n0 = torch.tensor([n - 13 * i for i in range(b)], dtype=torch.int32)
    n1 = torch.tensor([max(m - 29 * i, 1) for i in range(b)], dtype=torch.int32)
    for maxd, cc in ((math.inf, True), (0.9, True)):
        out = {}
        for two_pass in (0, 1):
            with ops().debug_knob(_lib.KB_KNOB_TC_ONE_PASS, 1 - two_pass):
                p, _, c, ws = ops().match_batched(a.to(DEV), d.to(DEV), n0.to(DEV), n1.to(DEV), maxd, cc, algo=1,
                                                  return_ws=True, want_dist=False)
                out[two_pass] = (p.clone(), c.clone(), int(ops().match_tc_debug(ws, b, n, m, dim)['n_col_rescan'][0]))
        assert torch.equal(out[0][1], out[1][1]), (maxd, out[0][1], out[1][1])
        for i in range(b):
            kk = int(out[0][1][i])
            assert torch.equal(out[0][0][i, :kk], out[1][0][i, :kk]), (maxd, i)
        assert out[1][2] == 0                                   # the two-pass path never queues a column
        assert out[0][2] <= b * (2 * dups + 8) + int(n0.sum()) // 50, out[0][2]   # the tied columns and the rare near-ties (< 2 % of the rows) are rescanned
        if dups:
            assert out[0][2] >= dups
        _exact_pairs_or_near_tie(out[0][0][0, :int(out[0][1][0])].cpu().numpy().astype(np.int64), a[0, :n0[0]].numpy(),
                                 d[0, :n1[0]].numpy(), maxd, cc)


@pytest.mark.parametrize('scale', [1e-6, 3e-4, 300.0, 5e4])
def test_matcher_fp16_operands_at_extreme_magnitudes(scale):
    """The default operand split for D > 64 is fp16: descriptors whose components underflow fp16's normal range (the
    absolute rounding term of the bound) or overflow it (|row|^2 above the limit: every row of the pair goes to the
    exact rescan) must still give the float64 oracle's pairs."""
    gen = torch.Generator().manual_seed(int(scale * 7) % 1000 + 3)
    n, m, dim = 260, 240, 128
    a = torch.randn(2, n, dim, generator=gen) * scale
    d = torch.randn(2, m, dim, generator=gen) * scale
    d[:, :150] = a[:, :150] + 0.05 * scale * torch.randn(2, 150, dim, generator=gen)
    typical = float(np.sqrt(2 * dim)) * scale
    for maxd, cc in ((math.inf, True), (0.7 * typical, True), (math.inf, False)):
        pairs, _, count = ops().match_batched(a.to(DEV), d.to(DEV), None, None, maxd, cc, algo=1, want_dist=False)
        for i in range(2):
            _exact_pairs_or_near_tie(pairs[i, :int(count[i])].cpu().numpy().astype(np.int64), a[i].numpy(), d[i].numpy(), maxd, cc)


@pytest.mark.parametrize('algo', [0, 1])
def test_matcher_ragged_batch_and_ties(algo):
    gen = torch.Generator().manual_seed(11)
    a = torch.randn(3, 300, 64, generator=gen)
    b = torch.randn(3, 280, 64, generator=gen)
    b[1, 7] = b[1, 3]                      # duplicate columns: first of ties must win (np.argmin)
    a[2, 9] = a[2, 4]
    n0 = torch.tensor([300, 123, 64], dtype=torch.int32)
    n1 = torch.tensor([280, 200, 1], dtype=torch.int32)
    pairs, dist, count = ops().match_batched(a.to(DEV), b.to(DEV), n0.to(DEV), n1.to(DEV), math.inf, True, algo=algo)
    for i in range(3):
        want = ref_ops.match_descriptors(a[i, :n0[i]].numpy(), b[i, :n1[i]].numpy(), cross_check=True)
        got = pairs[i, :int(count[i])].cpu().numpy().astype(np.int64)
        assert np.array_equal(got, want), i


def test_brute_force_matcher_dropin(golden):
    from keypoint_bench_b200.utils.matcher import brute_force_matcher
    g = golden('ref_match.npz')
    tag = 'small32'
    maxd, cc = g[f'{tag}__maxd_cc']
    params = {'metric': 'euclidean', 'max_distance': float(maxd), 'cross_check': bool(cc)}
    p0, p1 = torch.from_numpy(g[f'{tag}__p0']).to(DEV), torch.from_numpy(g[f'{tag}__p1']).to(DEV)
    r0, r1 = brute_force_matcher(p0, p1, torch.from_numpy(g[f'{tag}__desc0']).to(DEV),
                                 torch.from_numpy(g[f'{tag}__desc1']).to(DEV), params)
    assert np.array_equal(r0.cpu().numpy(), g[f'{tag}__r0'])
    assert np.array_equal(r1.cpu().numpy(), g[f'{tag}__r1'])
    with pytest.raises(ValueError):
        brute_force_matcher(p0[:0], p1, torch.from_numpy(g[f'{tag}__desc0']).to(DEV),
                            torch.from_numpy(g[f'{tag}__desc1']).to(DEV), params)
    with pytest.raises(ValueError):
        brute_force_matcher(p0, p1, torch.from_numpy(g[f'{tag}__desc0']).to(DEV),
                            torch.from_numpy(g[f'{tag}__desc1']).to(DEV), dict(params, metric='hamming'))


# ------------------------------------------------------------------------------------------------ evaluation

def test_warp_and_val_key_points_match_reference_fixtures(golden):
    from keypoint_bench_b200.tasks.repeatability import val_key_points
    from keypoint_bench_b200.utils.projection import warp
    g = golden('ref_eval.npz')
    for tag in ('rep_480x640', 'rep_small'):
        h, w = [int(v) for v in g[f'{tag}__hw']]
        hm = torch.from_numpy(g[f'{tag}__H'])
        w01, w10 = synth.warp_params(hm, h, w)
        k0 = torch.from_numpy(g[f'{tag}__k0']).to(DEV)
        k1 = torch.from_numpy(g[f'{tag}__k1']).to(DEV)
        a, b, ids, ids_out = warp(k0, w01)
        assert ids.dtype == torch.int64
        assert np.array_equal(ids.cpu().numpy(), g[f'{tag}__ids'])
        assert np.array_equal(ids_out.cpu().numpy(), g[f'{tag}__ids_out'])
        assert np.allclose(a.cpu().numpy(), g[f'{tag}__warp_valid'], rtol=1e-5, atol=1e-6)
        assert np.allclose(b.cpu().numpy(), g[f'{tag}__warp_proj'], rtol=1e-5, atol=1e-6)
        res = val_key_points(k0, k1, w01, w10, th=3, return_pairs=True)
        assert res['num_feat'] == int(g[f'{tag}__num_feat'])
        # (a) against the oracle, whose warp uses the same three rounded products as the kernel: EXACT counts and pairs
        ora = ref_ops.val_key_points(g[f'{tag}__k0'], g[f'{tag}__k1'], w01, w10, th=3)
        got_pairs = set(map(tuple, res['pairs'].tolist()))
        want_pairs = set(map(tuple, ora['pairs'].tolist()))
        assert got_pairs == want_pairs and res['gt_num'] == ora['gt_num']
        assert np.array_equal(res['errors'].cpu().numpy(), ora['errors'])
        assert abs(float(res['mean_error']) - ora['mean_error']) < 1e-5
        # (b) against the fixture minted from the reference itself, whose torch.einsum rounds the warped coordinates
        # differently by up to 1 ulp: mutual_argmin quantises distances to 2^-7 once the 99999 diagonal is present
        # (repeatability.py:18-32), so such an ulp can move an entry across a bucket edge.  Every pair that differs must
        # be such an edge case (within one bucket of both its row and its column maximum), and gt_num can differ by at
        # most the number of such pairs.
        ref_pairs = set(map(tuple, g[f'{tag}__pairs'].tolist())) if f'{tag}__pairs' in g.files else None
        dm = oracle_dist_mutual(g[f'{tag}__k0'], g[f'{tag}__k1'], w01, w10)
        n_edge = explain_repeat_pair_diffs(got_pairs, ref_pairs, dm) if ref_pairs is not None else 1
        assert abs(res['gt_num'] - int(g[f'{tag}__gt_num'])) <= n_edge, (res['gt_num'], int(g[f'{tag}__gt_num']), n_edge)
        assert abs(float(res['mean_error']) - float(g[f'{tag}__mean_error'])) < 1e-3
        assert np.allclose(res['errors'].cpu().numpy(), g[f'{tag}__errors'], rtol=1e-5, atol=1e-5 * 512)


def test_val_key_points_empty_and_unknown_mode():
    from keypoint_bench_b200.tasks.repeatability import val_key_points
    from keypoint_bench_b200.utils.projection import warp
    hm = torch.tensor([[1.0, 0, 5000.0], [0, 1.0, 0], [0, 0, 1.0]])       # pushes everything out of view
    w01, w10 = synth.warp_params(hm, 100, 100)
    k = torch.rand(20, 3, device=DEV)
    res = val_key_points(k, k, w01, w10)
    assert res == {'num_feat': 0, 'repeatability': 0, 'mean_error': 0, 'errors': None}
    with pytest.raises(ValueError):
        warp(k, {'mode': 'bogus'})


def test_corner_error_and_mha(golden):
    from keypoint_bench_b200.tasks.MHA import mha
    g = golden('ref_eval.npz')
    h, w, c = [int(v) for v in g['mha__hwc']]
    seed = int(g['mha__seed'])
    hm = synth.homography(seed + 7)
    s0 = synth.score_map('uniform', h, w, seed)
    s1 = synth.warp_map(s0, hm, 'nearest')
    w01, w10 = synth.warp_params(hm, h, w)
    params = {'extractor_params': dict(nms_dist=6, threshold=0, border_dist=8, top_k=300, min_score=0.0),
              'matcher_params': {'brute_force_params': {'metric': 'euclidean', 'max_distance': 5, 'cross_check': True}},
              'MHA_params': {'th': [3, 5, 7]}}
    img = torch.zeros(1, 3, h, w)
    flags = mha(0, img, s0.to(DEV), torch.from_numpy(g['mha__d0']).to(DEV), img, s1.to(DEV),
                torch.from_numpy(g['mha__d1']).to(DEV), w01, w10, params)
    assert list(flags) == list(g['mha__flags'])
    # corner error kernel against the float64 restatement
    rng = np.random.default_rng(0)
    he = np.stack([hm.numpy().astype(np.float64) + 1e-3 * rng.standard_normal((3, 3)) * [[1, 1, 50], [1, 1, 50], [1e-3, 1e-3, 0]]
                   for _ in range(5)])
    md, fl = ops().corner_error_batched(torch.from_numpy(he).to(DEV), hm.double()[None].expand(5, 3, 3).to(DEV), None,
                                        w, h, h, w, [3, 5, 7])
    for i in range(5):
        f, m = ref_ops.corner_error_flags(he[i], hm.numpy().astype(np.float64), w, h, h, w, (3, 5, 7))
        assert abs(m - float(md[i])) < 1e-9 and f == fl[i].cpu().tolist()


# ------------------------------------------------------------------------------------------------ sparse path

def test_detect_paths_sparse_and_fallback(round1):
    """uniform maps are certified by the sparse path (path 1); maps with negative scores, heavy ties or
    a top_k the candidate budget cannot reach fall back to the round-faithful kernel (path 2); both
    must equal the oracle."""
    params = dict(nms_dist=6, threshold=0.0, border_dist=8, top_k=1000, min_score=0.0)
    kinds = ['uniform', 'mixed', 'ties', 'uniform', 'alike', 'relu', 'negative', 'ramp']
    maps = [synth.score_map(k, 480, 640, 300 + i) for i, k in enumerate(kinds)]
    xyp, count, raster, path = ops().detect_batched(torch.cat(maps, 0).to(DEV), params, phases=round1)
    path = path.cpu().tolist()
    assert path[0] == 1 and path[3] == 1, path
    assert path[1] == 2 and path[6] == 2, path
    for i, m in enumerate(maps):
        want, want_r = ref_ops.detection(m, params)
        n = int(count[i])
        assert n == want.shape[0], (kinds[i], n, want.shape[0])
        assert np.array_equal(raster[i, :n].cpu().numpy().astype(np.int64), want_r), kinds[i]
        assert np.array_equal(xyp[i, :n].cpu().numpy(), want), kinds[i]


@pytest.mark.parametrize('seed', range(24))
def test_detect_sparse_randomised_against_greedy_oracle(seed, round1):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(40, 300)), int(rng.integers(40, 400))
    params = dict(nms_dist=int(rng.integers(1, 9)), threshold=float(rng.choice([0.0, 0.0, 0.3, 0.9])),
                  border_dist=int(rng.integers(0, 12)), top_k=int(rng.choice([1, 7, 50, 300, 2000])),
                  min_score=float(rng.choice([0.0, 0.0, 0.5, 0.97])))
    kind = ['uniform', 'relu', 'alike', 'ties'][seed % 4]
    m = synth.score_map(kind, h, w, 500 + seed)
    want, want_r = ref_ops.detection(m, params)
    xyp, count, raster, path = ops().detect_batched(m.to(DEV), params, phases=round1)
    n = int(count[0])
    assert n == want.shape[0], (params, kind, h, w, int(path[0]))
    assert np.array_equal(raster[0, :n].cpu().numpy().astype(np.int64), want_r), (params, kind, int(path[0]))
    assert np.array_equal(xyp[0, :n].cpu().numpy(), want)


@pytest.mark.parametrize('h,w,r,top_k,nmaps', [(480, 640, 6, 1000, 96), (376, 1241, 6, 1000, 57), (480, 640, 4, 4096, 80),
                                               (203, 517, 3, 300, 151)])
def test_detect_large_batch_stream_equals_tiled_and_oracle(h, w, r, top_k, nmaps):
    """A batch large enough for the automatic choice to be the packed streaming round-1 kernel (CTAs walk several bands,
    bands cross map boundaries, an odd batch leaves the last pair half empty): automatic == forced packed == forced fp32
    streaming == forced tiled, bit for bit, and a few maps of the batch equal the greedy oracle."""
    params = dict(nms_dist=r, threshold=0.0, border_dist=8, top_k=top_k, min_score=0.0)
    gen = torch.Generator(device=DEV).manual_seed(4242 + h)
    s = torch.rand(nmaps, 1, h, w, generator=gen, device=DEV)
    s[3] = torch.floor(s[3] * 8) / 8                          # a tie-heavy map in the middle of the batch
    s[5, 0, h // 2, :] = 0.999                                # a row of equal values: first-of-ties rule
    s[6, 0, :, w // 3] = 0.9995                               # a column of equal values
    s[7] -= 0.5                                               # negative scores -> flagged for the round-faithful kernel
    outs = {}
    for name, ph in (('auto', 7), ('tiled', 7 | 8), ('stream', 7 | 16), ('packed', 7 | 32)):
        with torch.no_grad():
            xyp, count, raster, path = ops().detect_batched(s, params, phases=ph)
        outs[name] = (xyp.cpu().numpy(), count.cpu().numpy(), raster.cpu().numpy(), path.cpu().numpy())
    for name in ('tiled', 'stream', 'packed'):
        assert np.array_equal(outs['auto'][1], outs[name][1]), name
        assert np.array_equal(outs['auto'][3], outs[name][3]), name
        for b in range(nmaps):
            n = int(outs['auto'][1][b])
            assert np.array_equal(outs['auto'][2][b, :n], outs[name][2][b, :n]), (name, b)
            assert np.array_equal(outs['auto'][0][b, :n], outs[name][0][b, :n]), (name, b)
    assert outs['auto'][3][0] == 1 and outs['auto'][3][7] == 2, outs['auto'][3][:8]
    for b in (0, 3, 5, 6, 7, nmaps - 1):
        want, want_r = ref_ops.detection(s[b:b + 1].cpu(), params, nms='greedy' if b != 7 else 'separable')
        n = int(outs['auto'][1][b])
        assert n == want.shape[0], (b, n, want.shape[0])
        assert np.array_equal(outs['auto'][2][b, :n].astype(np.int64), want_r), b
        assert np.array_equal(outs['auto'][0][b, :n], want), b


@pytest.mark.parametrize('scale', [1e-12, 3e-7, 1.0, 7e4, 1e12])
def test_detect_packed_kernel_is_scale_free(scale):
    """The packed round-1 kernel decides on a 16-bit image of (score - tau) * qscale, qscale from the largest sampled score:
    maps whose scores live far from 1 (fp16 would underflow to 0 or overflow to inf without it) give the tiled kernel's
    rows bit for bit, and the greedy oracle's on one map."""
    params = dict(nms_dist=6, threshold=0.0, border_dist=8, top_k=400, min_score=0.0)
    gen = torch.Generator(device=DEV).manual_seed(99)
    s = torch.rand(6, 1, 240, 320, generator=gen, device=DEV) * scale
    s[1, 0, 100:140, 100:160] = 0.0                          # an exact-zero region
    s[2] = s[2] * 1e-3                                        # a map three decades below its pair partner
    out = {}
    for name, ph in (('tiled', 7 | 8), ('packed', 7 | 32)):
        xyp, count, raster, path = ops().detect_batched(s, params, phases=ph)
        out[name] = (xyp.cpu().numpy(), count.cpu().numpy(), raster.cpu().numpy(), path.cpu().numpy())
    assert np.array_equal(out['tiled'][1], out['packed'][1])
    assert (out['packed'][3] == 1).all()
    for b in range(6):
        n = int(out['tiled'][1][b])
        assert np.array_equal(out['tiled'][2][b, :n], out['packed'][2][b, :n]), b
        assert np.array_equal(out['tiled'][0][b, :n], out['packed'][0][b, :n]), b
    want, want_r = ref_ops.detection(s[2:3].cpu(), params, nms='greedy')
    n = int(out['packed'][1][2])
    assert np.array_equal(out['packed'][2][2, :n].astype(np.int64), want_r)
    assert np.array_equal(out['packed'][0][2, :n], want)


# ------------------------------------------------------------------------------------------------ tensor-core matcher internals

def test_matcher_auto_falls_back_beyond_256_dims():
    gen = torch.Generator().manual_seed(5)
    a = torch.randn(1, 150, 300, generator=gen)
    b = torch.randn(1, 170, 300, generator=gen)
    pairs, dist, count = ops().match_batched(a.to(DEV), b.to(DEV), None, None, math.inf, True)      # algo = -1
    want = ref_ops.match_descriptors(a[0].numpy(), b[0].numpy(), cross_check=True)
    assert np.array_equal(pairs[0, :int(count[0])].cpu().numpy().astype(np.int64), want)


@pytest.mark.parametrize('bf16x3', [2, 1])
@pytest.mark.parametrize('n,m,dim,normed', [(1000, 977, 256, True), (700, 900, 64, False), (300, 300, 128, True)])
def test_tensor_core_scores_stay_inside_the_certification_bound(n, m, dim, normed, bf16x3):
    """The tensor-core Gram scores t = x.y - |y|^2/2 -- two fp16 products (default) or three bf16 products
    (KB_KNOB_TC_BF16X3) -- must lie within the a-priori bound the resolver certifies with (tc_err_bound,
    kb_match_tc.cu), measured against float64; both operand splits give the same pairs."""
    from keypoint_bench_b200 import _lib
    gen = torch.Generator().manual_seed(n + dim)
    a = torch.randn(2, n, dim, generator=gen) * (1.0 if normed else 3.0)
    b = torch.randn(2, m, dim, generator=gen) * (1.0 if normed else 3.0)
    if normed:
        a, b = torch.nn.functional.normalize(a, dim=2), torch.nn.functional.normalize(b, dim=2)
    a, b = a.to(DEV), b.to(DEV)
    with ops().debug_knob(_lib.KB_KNOB_TC_BF16X3, bf16x3):
        pairs, dist, count, ws = ops().match_batched(a, b, None, None, math.inf, True, algo=1, return_ws=True)
    p0, _, c0 = ops().match_batched(a, b, None, None, math.inf, True, algo=0)
    assert torch.equal(count, c0) and all(torch.equal(pairs[i, :int(c0[i])], p0[i, :int(c0[i])]) for i in range(2))
    dbg = ops().match_tc_debug(ws, 2, n, m, dim)
    best, second, idx = dbg['res0']
    a64, b64 = a.double(), b.double()
    idx_l = idx.long().reshape(2, n)
    assert int(idx_l.min()) >= 0 and int(idx_l.max()) < m
    yb = torch.gather(b64, 1, idx_l[..., None].expand(2, n, dim))
    t_exact = (a64 * yb).sum(-1) - 0.5 * (yb * yb).sum(-1)
    err = (best.reshape(2, n).double() - t_exact).abs()
    nbmax = b64.norm(dim=2).max(dim=1).values[:, None]
    if bf16x3 == 1:
        bound = 6.2e-5 * a64.norm(dim=2) * nbmax + 3.1e-5 * nbmax * nbmax
    else:
        dp = (dim + 63) // 64 * 64
        bound = 2.8e-4 * a64.norm(dim=2) * nbmax + 3.1e-5 * nbmax * nbmax + 3.0e-8 * dp ** 0.5 * (a64.norm(dim=2) + nbmax)
    ratio = float((err / bound).max())
    print(f'tensor-core score error / bound ({"bf16 x 3" if bf16x3 == 1 else "fp16 x 2"}, D={dim}): {ratio:.3f}')
    assert ratio < 0.5                                # observed: bf16 x 3 ~0.03; the bound keeps a wide margin
    # and the reported best really is the maximum of the exact scores up to that bound
    t_all = a64 @ b64.transpose(1, 2) - 0.5 * (b64 * b64).sum(-1)[:, None, :]
    assert bool(((t_all.max(dim=2).values - t_exact) <= 2 * bound).all())


def test_sampling_plane_staged_path_cfg2_shape():
    """[.,256,60,80] maps with ~1000 keypoints take the plane-staged kernel; compare with the oracle and
    with torch's own CUDA grid_sample (tolerance 1e-5, north_star)."""
    gen = torch.Generator().manual_seed(21)
    d = torch.randn(3, 256, 60, 80, generator=gen)
    p = torch.rand(3, 1000, 3, generator=gen)
    p[0, :5, :2] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [1.0, 0.0], [-0.01, 0.5], [0.5, 1.02]])
    cnt = torch.tensor([1000, 873, 1], dtype=torch.int32)
    out = ops().sample_batched(d.to(DEV), p.to(DEV), cnt.to(DEV))
    for b in range(3):
        n = int(cnt[b])
        grid = ((p[b, :n, :2] - 0.5) * 2)[None, None].to(DEV)
        want = torch.nn.functional.grid_sample(d[b:b + 1].to(DEV), grid, align_corners=True)[0, :, 0].T
        assert torch.allclose(out[b, :n], want, rtol=1e-5, atol=1e-5)
        ref = ref_ops.sample_brute_force(d[b].numpy(), p[b, :n].numpy())
        assert np.allclose(out[b, :n].cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    assert float(out[1, 873:].abs().max()) == 0.0          # rows beyond the count stay untouched (zero)
    # the eight-channel staged kernel (the default here) and the four-channel one give the same bits, also normalised
    from keypoint_bench_b200 import _lib
    for norm in (False, True):
        a = ops().sample_batched(d.to(DEV), p.to(DEV), cnt.to(DEV), normalize=norm)
        with ops().debug_knob(_lib.KB_KNOB_SAMPLE_4CH, 1):
            b4 = ops().sample_batched(d.to(DEV), p.to(DEV), cnt.to(DEV), normalize=norm)
        assert torch.equal(a, b4), norm


def test_sampling_staged_kernels_agree_beyond_1024_keypoints():
    """cfg3's shape ([.,64,60,80] maps, 4096 keypoints) is beyond the eight-channel staged kernel's 1024 keypoints per map
    (one CTA per 1024 keypoints was measured slower there: 147 vs 135 us, the planes are re-staged per CTA) and takes the
    four-channel kernel with or without the knob: ragged counts against the oracle, rows beyond the counts untouched."""
    from keypoint_bench_b200 import _lib
    gen = torch.Generator().manual_seed(33)
    d = torch.randn(4, 64, 60, 80, generator=gen)
    p = torch.rand(4, 4096, 3, generator=gen)
    cnt = torch.tensor([4096, 1024, 1025, 3000], dtype=torch.int32)
    a = ops().sample_batched(d.to(DEV), p.to(DEV), cnt.to(DEV))
    with ops().debug_knob(_lib.KB_KNOB_SAMPLE_4CH, 1):
        b4 = ops().sample_batched(d.to(DEV), p.to(DEV), cnt.to(DEV))
    assert torch.equal(a, b4)
    for b in (0, 2):
        n = int(cnt[b])
        ref = ref_ops.sample_brute_force(d[b].numpy(), p[b, :n].numpy())
        assert np.allclose(a[b, :n].cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    assert float(a[1, 1024:].abs().max()) == 0.0 and float(a[2, 1025:].abs().max()) == 0.0


@pytest.mark.parametrize('pairs,c,h,w,n,seed', [(3, 256, 60, 80, 1000, 5), (2, 64, 30, 40, 517, 6), (1, 128, 16, 24, 1024, 7),
                                                (5, 192, 15, 20, 300, 8)])
def test_fused_sampling_and_operand_preparation_equals_the_two_calls(pairs, c, h, w, n, seed):
    """kb_sample_desc_operands + kb_match_mnn(phases 6 | 8): the sampler writes the matcher's operand rows (fp16 halves
    for C > 64, bf16 for C = 64) and partial norms itself.  Same float32 rows bit for bit, same pairs and distances
    as sampling and matching separately, and the pairs equal the float64 oracle on the sampled rows (ragged counts
    incl. an empty map, duplicated keypoints = exact ties)."""
    gen = torch.Generator().manual_seed(seed)
    d = torch.randn(2 * pairs, c, h, w, generator=gen)
    pt = torch.rand(2 * pairs, n, 3, generator=gen)
    pt[0, 10:20, :2] = pt[0, 0:10, :2]                     # duplicated keypoints: identical descriptor rows
    pt[pairs, :4, :2] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [-0.01, 0.5], [0.5, 1.02]])
    cnt = torch.randint(n // 2, n + 1, (2 * pairs,), generator=gen).to(torch.int32)
    cnt[0] = n
    if pairs > 2:
        cnt[pairs + 1] = 0                                  # an empty image 1
    dd, pp, nn = d.to(DEV), pt.to(DEV), cnt.to(DEV)
    for md in (math.inf, 0.9 * math.sqrt(2.0 * c)):
        o = ops()
        f_d, f_pairs, f_dist, f_cnt = o.sample_match_batched(dd, pp, nn, pairs, md, True, want_dist=True, fused=True)
        u_d, u_pairs, u_dist, u_cnt = o.sample_match_batched(dd, pp, nn, pairs, md, True, want_dist=True, fused=False)
        assert torch.equal(f_cnt, u_cnt)
        for b in range(2 * pairs):
            k = int(cnt[b])
            assert torch.equal(f_d[b, :k], u_d[b, :k]), b
        for b in range(pairs):
            k = int(f_cnt[b])
            assert torch.equal(f_pairs[b, :k], u_pairs[b, :k]), b
            assert torch.equal(f_dist[b, :k], u_dist[b, :k]), b
            n0, n1 = int(cnt[b]), int(cnt[pairs + b])
            if n0 and n1:
                _exact_pairs_or_near_tie(f_pairs[b, :k].cpu().numpy().astype(np.int64), u_d[b, :n0].cpu().numpy(),
                                         u_d[pairs + b, :n1].cpu().numpy(), md, True)
            else:
                assert k == 0
        # without distances (what the pipeline asks for)
        g_d, g_pairs, _, g_cnt = o.sample_match_batched(dd, pp, nn, pairs, md, True, want_dist=False, fused=True)
        assert torch.equal(g_cnt, f_cnt)
        for b in range(pairs):
            k = int(f_cnt[b])
            assert torch.equal(g_pairs[b, :k], f_pairs[b, :k]), b


# ------------------------------------------------------------------------------------------------ full-size configs

@pytest.mark.parametrize('kind,h,w,r,top_k,want_path', [
    ('alike', 480, 640, 6, 1000, 1),        # smooth, ALIKE-like map: blobs collapse to round-1 maxima
    ('uniform', 480, 640, 4, 4096, 1),      # cfg3: top_k binds only with every listed candidate on chip
    ('uniform', 376, 1241, 6, 1000, 1),     # cfg5: odd width, rows not 16-byte aligned (scalar staging)
    ('relu', 480, 640, 6, 1000, 1),         # exact zeros (KeyNet-like)
    ('ties', 480, 640, 6, 1000, None),      # eight distinct values: whichever path, result must be exact
    ('alike', 240, 320, 3, 300, None),      # odd radius (unaligned halo)
])
def test_detection_full_size_configs_against_greedy_oracle(kind, h, w, r, top_k, want_path, round1):
    params = dict(nms_dist=r, threshold=0.0, border_dist=8, top_k=top_k, min_score=0.0)
    m = synth.score_map(kind, h, w, 77 + r)
    want, want_r = ref_ops.detection(m, params, nms='greedy')
    xyp, count, raster, path = ops().detect_batched(m.to(DEV), params, phases=round1)
    n = int(count[0])
    assert n == want.shape[0], (kind, n, want.shape[0], int(path[0]))
    assert np.array_equal(raster[0, :n].cpu().numpy().astype(np.int64), want_r), (kind, int(path[0]))
    assert np.array_equal(xyp[0, :n].cpu().numpy(), want)
    if want_path is not None:
        assert int(path[0]) == want_path, (kind, int(path[0]))


def test_pipeline_matches_per_pair_dropins_and_graph_replay():
    """The batched pipeline (what bench.py times), its CUDA-graph replay and the per-pair drop-ins agree."""
    from keypoint_bench_b200 import pipeline
    from keypoint_bench_b200.utils.extracter import detection
    from keypoint_bench_b200.utils.matcher import brute_force_matcher
    from keypoint_bench_b200.utils.projection import warp
    cfg = synth.PathConfig('t', 120, 160, 64, 8, True, 4, 200)
    pairs = [synth.make_pair(cfg, 7, i) for i in range(3)]
    P = len(pairs)
    score = torch.cat([p['score0'] for p in pairs] + [p['score1'] for p in pairs]).to(DEV)
    desc = torch.cat([p['desc0'] for p in pairs] + [p['desc1'] for p in pairs]).to(DEV)
    h01 = torch.stack([p['H'] for p in pairs]).reshape(P, 9)
    h10 = torch.stack([torch.linalg.inv(p['H'].double()).float() for p in pairs]).reshape(P, 9)
    wh = torch.tensor([[160.0, 120.0]]).expand(2 * P, 2).contiguous()
    batch = pipeline.PairBatch(score, desc, torch.cat([h01, h10]).to(DEV), wh.to(DEV))
    res = pipeline.extract_match(batch, cfg)
    graphed = pipeline.GraphedStep(lambda: pipeline.extract_match(batch, cfg))
    res_g = graphed()
    torch.cuda.synchronize()
    for i, p in enumerate(pairs):
        k0 = detection(p['score0'].to(DEV), cfg.extractor_params)
        k1 = detection(p['score1'].to(DEV), cfg.extractor_params)
        k0c, _, _, _ = warp(k0, p['warp01'])
        k1c, _, _, _ = warp(k1, p['warp10'])
        m0, m1 = brute_force_matcher(k0c, k1c, p['desc0'].to(DEV), p['desc1'].to(DEV), cfg.matcher_params)
        for r in (res, res_g):
            n = int(r['n_matches'][i])
            assert n == m0.shape[0]
            idx = r['matches'][i, :n].long()
            assert torch.equal(r['kcov'][i][idx[:, 0]], m0[:, :2]) and torch.equal(r['kcov'][P + i][idx[:, 1]], m1[:, :2])


# ------------------------------------------------------------------------------------------------ warp(mode='se3')

def test_warp_se3_matches_reference_fixtures_and_oracle(golden):
    """utils/projection.py:194-267: ids exact, coordinates 1e-5 (the reference's own torch.inverse /
    einsum roundings are not reproducible bit for bit)."""
    from keypoint_bench_b200.utils.projection import warp
    from oracle.make_golden import SE3_CASES
    g = golden('ref_se3.npz')
    for tag, h, w, seed, n in SE3_CASES:
        params = synth.se3_scene(h, w, seed)
        kp = torch.from_numpy(g[f'{tag}__kp']).to(DEV)
        a, b, ids, ids_out = warp(kp, params)
        assert np.array_equal(ids.cpu().numpy(), g[f'{tag}__ids']), tag
        assert np.array_equal(ids_out.cpu().numpy(), g[f'{tag}__ids_out']), tag
        assert np.allclose(a.cpu().numpy(), g[f'{tag}__valid'], rtol=1e-5, atol=1e-6)
        assert np.allclose(b.cpu().numpy(), g[f'{tag}__proj'], rtol=1e-5, atol=1e-5)
    # batched entry point, ragged counts, against the oracle
    params = [synth.se3_scene(120, 160, 50 + i) for i in range(3)]
    gen = torch.Generator().manual_seed(9)
    pts = torch.rand(3, 300, 2, generator=gen)
    cnt = torch.tensor([300, 17, 0], dtype=torch.int32)
    st = lambda k: torch.stack([p[k] for p in params]).to(DEV)      # noqa: E731
    kv, kw, ids, ids_out, nv, no = ops().warp_se3_batched(pts.to(DEV), cnt.to(DEV), st('depth0'), st('depth1'), st('intrinsics0'),
                                                          st('intrinsics1'), st('pose01'), st('bbox0'), st('bbox1'))
    for i in range(3):
        wa, wb, wids, wout = ref_ops.warp_se3(pts[i, :int(cnt[i])].numpy(), params[i])
        a, o = int(nv[i]), int(no[i])
        assert np.array_equal(ids[i, :a].cpu().numpy().astype(np.int64), wids), i
        assert np.array_equal(ids_out[i, :o].cpu().numpy().astype(np.int64), wout), i
        assert np.allclose(kv[i, :a].cpu().numpy(), wa, rtol=1e-5, atol=1e-6) and np.allclose(kw[i, :a].cpu().numpy(), wb, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ LightGlue-style extract

def test_simple_nms_and_lightglue_extract_match_reference_fixtures(golden):
    from keypoint_bench_b200.utils import lightglue_extract as lg
    from oracle.make_golden import LG_CASES
    from test_oracle_golden import _lg_compare
    g = golden('ref_lightglue.npz')
    for tag, kind, h, w, seed, c, s in LG_CASES:
        sc = synth.score_map(kind, h, w, seed).to(DEV)
        for r in (0, 2, 5):
            if f'{tag}__nms{r}' in g.files:
                got = lg.simple_nms(sc[0], r)
                assert got.shape == sc[0].shape
                assert np.array_equal(got[0].cpu().numpy(), g[f'{tag}__nms{r}']), (tag, r)
        dm = torch.from_numpy(g[f'{tag}__dm']).to(DEV)
        feats = lg.extract(lambda img: (sc, dm), None, s)
        assert feats['keypoints'].shape[0] == 1 and feats['descriptors'].shape[2] == c
        _lg_compare(feats['keypoints'][0].cpu().numpy(), feats['keypoint_scores'][0].cpu().numpy(),
                    feats['descriptors'][0].cpu().numpy(), g, tag, w)


@pytest.mark.parametrize('kind,h,w,r', [('uniform', 480, 640, 5), ('ties', 100, 333, 4), ('alike', 376, 1241, 11),
                                        ('negative', 70, 90, 3), ('mixed', 65, 129, 16), ('relu', 33, 65, 1)])
def test_simple_nms_against_oracle(kind, h, w, r):
    s = torch.cat([synth.score_map(kind, h, w, 300 + r), synth.score_map('uniform', h, w, 301 + r)], 0)
    got = ops().simple_nms_batched(s.to(DEV), r)
    assert np.array_equal(got.cpu().numpy()[:, 0], ref_ops.simple_nms(s.numpy()[:, 0], r))


def test_lightglue_extract_batched_against_oracle():
    from keypoint_bench_b200.utils import lightglue_extract as lg
    maps = torch.cat([synth.score_map(k, 240, 320, 70 + i) for i, k in enumerate(['uniform', 'alike', 'relu'])], 0)
    gen = torch.Generator().manual_seed(9)
    dm = torch.randn(3, 48, 30, 40, generator=gen)
    kps, val, desc, count = lg.extract_batched(maps.to(DEV), dm.to(DEV), 8, max_num_kps=400)
    for i in range(3):
        wkp, wval, wdesc, _ = ref_ops.lightglue_extract(maps[i].numpy(), dm[i].numpy(), 8, max_num_kps=400)
        n = int(count[i])
        assert n == wkp.shape[0]
        assert np.array_equal(kps[i, :n].cpu().numpy(), wkp) and np.array_equal(val[i, :n].cpu().numpy(), wval)
        assert np.allclose(desc[i, :n].cpu().numpy(), wdesc, rtol=1e-5, atol=1e-5)
        assert not kps[i, n:].any() and not desc[i, n:].any()


# ------------------------------------------------------------------------------------------------ Lucas-Kanade tracker

def test_lk_tracker_matches_reference_fixtures(golden):
    """optical_flow_tensor (utils/matcher.py:188-203) on the fixtures minted from the reference, the random start
    replayed.  Up to 40 float32 Gauss-Newton steps per level over win^2*C-sample sums: the reference's own CPU result and
    its numpy restatement differ by up to 1.5e-5 px (oracle/REFCHECK.log); LK_TOL is the bound asserted here."""
    from keypoint_bench_b200.utils.matcher import OpticalFlow
    from oracle.make_golden import LK_CASES
    g = golden('ref_lk.npz')
    for tag, c, h, w, win, levels, iters, dist, seed, n in LK_CASES:
        params = {'distance': dist, 'win_size': win, 'levels': levels, 'interation': iters, 'gray': c == 1}
        img0, img1 = torch.from_numpy(g[f'{tag}__img0']).to(DEV), torch.from_numpy(g[f'{tag}__img1']).to(DEV)
        pts = torch.from_numpy(g[f'{tag}__pts']).to(DEV)
        start = torch.from_numpy(g[f'{tag}__init'])[None].to(DEV)
        out, err = OpticalFlow(params)(img0, img1, pts, pts, start=start)
        assert out.shape == (1, n, 2) and err.shape == (1, n)
        d = np.abs(out[0].cpu().numpy() - g[f'{tag}__out']).max()
        print(f'LK max |gpu - reference| {tag}: {d:.3e} px')
        assert d < LK_TOL, (tag, d)


def test_optical_flow_tensor_dropin_and_batched_against_oracle():
    from keypoint_bench_b200.utils.matcher import optical_flow_tensor, optical_flow_cv
    params = {'distance': 4, 'win_size': 11, 'levels': 2, 'interation': 8, 'gray': False}
    scenes = [synth.lk_scene(3, 120, 160, 20 + i, shift=(1.5 + i, -2.0)) for i in range(2)]
    img0 = torch.cat([s[0] for s in scenes], 0)
    img1 = torch.cat([s[1] for s in scenes], 0)
    gen = torch.Generator().manual_seed(4)
    pts = torch.rand(2, 90, 2, generator=gen)
    scale = torch.tensor([159.0, 119.0])
    init = pts * scale + torch.randn(2, 90, 2, generator=gen) * 2
    cnt = torch.tensor([90, 41], dtype=torch.int32)
    out = ops().lk_track_batched(img0.to(DEV), img1.to(DEV), (pts * scale).to(DEV), init.to(DEV), cnt.to(DEV), 11, 2, 8)
    for b in range(2):
        k = int(cnt[b])
        want = ref_ops.lk_track(img0[b].numpy(), img1[b].numpy(), (pts[b, :k] * scale).numpy(), init[b, :k].numpy(), 11, 2, 8)
        d = np.abs(out[b, :k].cpu().numpy() - want).max()
        print(f'LK max |gpu - oracle| batch {b}: {d:.3e} px')
        assert d < LK_TOL, (b, d)
        assert not out[b, k:].any()
    # the drop-in draws its own random start (matcher.py:55): most points must land on the true shift
    torch.manual_seed(0)
    params = {'distance': 3, 'win_size': 21, 'levels': 3, 'interation': 40, 'gray': False}
    got = optical_flow_tensor(pts[0, :, :2].to(DEV), pts[0, :, :2].to(DEV), img0[0:1].to(DEV), img1[0:1].to(DEV), params)
    assert got.shape == (1, 90, 2)
    inner = ((pts[0] > 0.25) & (pts[0] < 0.75)).all(dim=1)
    err = (got[0].cpu() - (pts[0] * scale + torch.tensor([1.5, -2.0]))).norm(dim=1)[inner]
    assert (err < 0.5).float().mean() > 0.5
    with pytest.raises(NotImplementedError):
        optical_flow_cv(None, None, None, None)
    with pytest.raises(RuntimeError):
        optical_flow_tensor(pts[0].to(DEV), pts[0].to(DEV), img0[0:1, :1].to(DEV), img1[0:1, :1].to(DEV), params)


# ------------------------------------------------------------------------------------------------ frame streams

def test_stream_pipeline_equals_pairwise_pipeline_and_oracle():
    """extract_match_stream (every frame extracted once) gives, pair by pair, what the pairwise pipeline gives on
    (frame t-1, frame t) without the covisibility warp -- and what the oracle's brute_force_matcher gives."""
    import dataclasses
    from keypoint_bench_b200 import pipeline
    cfg = dataclasses.replace(synth.CONFIGS['cfg5'], height=120, width=200, top_k=300)
    f = 5
    gen = torch.Generator().manual_seed(21)
    score = torch.rand(f + 1, 1, cfg.height, cfg.width, generator=gen)
    desc = 2.67 * torch.nn.functional.normalize(torch.randn(f + 1, cfg.desc_dim, cfg.height, cfg.width, generator=gen), dim=1)
    for t in range(1, f + 1):                                   # consecutive frames share content: shift by 2 px + noise
        score[t, :, :, 2:] = score[t - 1, :, :, :-2]
        desc[t, :, :, 2:] = desc[t - 1, :, :, :-2] + 0.05 * torch.randn(cfg.desc_dim, cfg.height, cfg.width - 2, generator=gen)
    frames = pipeline.FrameBatch(score.to(DEV), desc.to(DEV))
    got = pipeline.extract_match_stream(frames, cfg)
    assert got['matches'].shape[0] == f
    # pairwise pipeline on the same frames: first f frames as image 0, last f as image 1
    eye = torch.eye(3).reshape(1, 9).repeat(2 * f, 1).to(DEV)
    wh = torch.tensor([[float(cfg.width), float(cfg.height)]]).repeat(2 * f, 1).to(DEV)
    pb = pipeline.PairBatch(torch.cat([score[:-1], score[1:]]).to(DEV), torch.cat([desc[:-1], desc[1:]]).to(DEV), eye, wh)
    want = pipeline.extract_match(pb, cfg, covisible_only=False)
    for i in range(f):
        k = int(got['n_matches'][i])
        assert k == int(want['n_matches'][i]) and k > 20
        assert torch.equal(got['matches'][i, :k], want['matches'][i, :k]), i
        assert int(got['n_kpts'][i]) == int(want['n_kpts'][i])
    # oracle on the first pair
    n0, n1 = int(got['n_kpts'][0]), int(got['n_kpts'][1])
    k0, k1 = got['kpts'][0, :n0].cpu().numpy(), got['kpts'][1, :n1].cpu().numpy()
    _, _, pairs = ref_ops.brute_force_matcher(k0, k1, desc[0:1].numpy(), desc[1:2].numpy(), cfg.matcher_params)
    k = int(got['n_matches'][0])
    assert np.array_equal(got['matches'][0, :k].cpu().numpy().astype(np.int64), pairs)


@pytest.mark.parametrize('seed', range(10))
def test_repeat_counts_randomised_sorted_and_exhaustive_paths(seed):
    """kb_repeat_counts against the oracle's dist_mutual / mutual_argmin arithmetic (tasks/repeatability.py:69-85) on
    random point sets: clustered partners, ragged counts, one-point sides, and coordinates scaled past the 99999
    diagonal mask (which sends the map to the exhaustive kernels instead of the sorted sweeps)."""
    rng = np.random.default_rng(7000 + seed)
    b = 3
    a_max, b_max = int(rng.integers(1, 400)), int(rng.integers(1, 400))
    scale = [1.0, 1.0, 5.0e4][seed % 3]                       # 5e4: distances beat the 99999 mask -> exhaustive path
    na = rng.integers(1, a_max + 1, size=b)
    nb = rng.integers(1, b_max + 1, size=b)
    k0c = rng.random((b, a_max, 2), dtype=np.float32) * scale
    k1c = rng.random((b, b_max, 2), dtype=np.float32) * scale
    k01c = (k0c + rng.normal(0, 0.002, k0c.shape).astype(np.float32) * scale).astype(np.float32)   # A points seen in image 1
    k10c = (k1c + rng.normal(0, 0.002, k1c.shape).astype(np.float32) * scale).astype(np.float32)
    m = min(a_max, b_max) // 2
    k1c[:, :m] = k01c[:, :m]                                   # true partners: B point j sits where A point j lands
    k10c[:, :m] = k0c[:, :m]
    if m > 3:                                                  # shuffle partners away from the masked diagonal
        perm = rng.permutation(m)
        k1c[:, :m] = k1c[:, perm]
        k10c[:, :m] = k10c[:, perm]
    th, s01, s10 = 3.0, 512.0, 512.0
    t = lambda x: torch.from_numpy(x).to(DEV)      # noqa: E731
    stats, errors, _ = ops().repeat_batched(t(k0c), t(k01c), t(na.astype(np.int32)), t(k1c), t(k10c),
                                            t(nb.astype(np.int32)), s01, s10, th, want_errors=True)
    stats, errors = stats.cpu().numpy(), errors.cpu().numpy()
    for i in range(b):
        A, Bn = int(na[i]), int(nb[i])
        d01 = ref_ops.keypoint_distance(k0c[i, :A], k10c[i, :Bn])
        d10 = ref_ops.keypoint_distance(k1c[i, :Bn], k01c[i, :A])
        dm = ((d01 + d10.T) / np.float32(2)).astype(np.float32)
        for q in range(min(A, Bn)):
            dm[q, q] = np.float32(99999)
        rows, cols = ref_ops.mutual_argmin(dm)
        dist = (dm[rows, cols] * np.float32(s01)).astype(np.float32)
        assert stats[i, 2] == rows.shape[0], (seed, i, scale)
        assert stats[i, 0] == int((dist <= th).sum()), (seed, i, scale)
        assert np.isclose(stats[i, 1], dist[dist <= th].astype(np.float64).sum(), rtol=1e-12, atol=1e-9)
        assert np.array_equal(errors[i, :A], (dm.min(axis=1) * np.float32(s10)).astype(np.float32)), (seed, i, scale)


def test_repeat_counts_sorted_and_tile_walking_kernels_agree():
    """The y-sorted sweeps (default up to 4096 points per side) and the tile-walking pruned kernels (KB_KNOB_REP_NO_SORT, also
    the path for larger sets) must give identical statistics and per-row minima."""
    rng = np.random.default_rng(99)
    b, a_max, b_max = 4, 1500, 1300
    k0c = rng.random((b, a_max, 2), dtype=np.float32)
    k1c = rng.random((b, b_max, 2), dtype=np.float32)
    k01c = (k0c + rng.normal(0, 0.001, k0c.shape)).astype(np.float32)
    k10c = (k1c + rng.normal(0, 0.001, k1c.shape)).astype(np.float32)
    k1c[:, 100:900] = k01c[:, 200:1000]
    k10c[:, 100:900] = k0c[:, 200:1000]
    na = torch.tensor([1500, 1200, 7, 1], dtype=torch.int32)
    nb = torch.tensor([1300, 1300, 900, 5], dtype=torch.int32)
    t = lambda x: torch.from_numpy(x).to(DEV)      # noqa: E731
    args = (t(k0c), t(k01c), na.to(DEV), t(k1c), t(k10c), nb.to(DEV), 512.0, 512.0, 3.0)
    s_sorted, e_sorted, _ = ops().repeat_batched(*args, want_errors=True)
    from keypoint_bench_b200 import _lib
    with ops().debug_knob(_lib.KB_KNOB_REP_NO_SORT, 1):
        s_tiles, e_tiles, _ = ops().repeat_batched(*args, want_errors=True)
    assert torch.equal(s_sorted, s_tiles)
    assert float(s_sorted[0, 0]) > 500
    for i in range(b):
        assert torch.equal(e_sorted[i, :int(na[i])], e_tiles[i, :int(na[i])]), i
