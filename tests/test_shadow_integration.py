"""Shadow-module integration run: the reference's OWN evaluators -- tasks/repeatability.py and
tasks/MHA.py, unmodified -- executed on top of this package's drop-ins for utils.extracter / utils.matcher /
utils.projection (installed in sys.modules as INTEGRATION.md section 3 shows), compared with the CPU oracle.

The reference modules come from KB_REFERENCE_ROOT, baseline/_ref (scripts/stage_reference.py copies the two task
modules and utils/visualization.py there; git-ignored) or /root/reference; without any of them the test skips."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from keypoint_bench_b200 import synth
from oracle import ref_ops

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = 'cuda'


def _reference_root():
    for cand in (os.environ.get('KB_REFERENCE_ROOT'), os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if cand and os.path.isfile(os.path.join(cand, 'tasks', 'repeatability.py')) and \
                os.path.isfile(os.path.join(cand, 'tasks', 'MHA.py')):
            return cand
    return None


@pytest.fixture()
def shadowed_reference():
    root = _reference_root()
    if root is None:
        pytest.skip('no reference task modules (KB_REFERENCE_ROOT / baseline/_ref / /root/reference): run scripts/stage_reference.py')
    import keypoint_bench_b200.utils.extracter as ex
    import keypoint_bench_b200.utils.matcher as ma
    import keypoint_bench_b200.utils.projection as pr
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == 'utils' or k.startswith('utils.') or k == 'tasks'
             or k.startswith('tasks.')}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, root)
    sys.modules['utils.extracter'] = ex
    sys.modules['utils.matcher'] = ma
    sys.modules['utils.projection'] = pr
    try:
        rep = importlib.import_module('tasks.repeatability')
        mha = importlib.import_module('tasks.MHA')
        assert os.path.abspath(rep.__file__).startswith(os.path.abspath(root))
        assert rep.detection is ex.detection and rep.warp is pr.warp and mha.brute_force_matcher is ma.brute_force_matcher
        yield rep, mha
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == 'utils' or k.startswith('utils.') or k == 'tasks' or k.startswith('tasks.')]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_reference_evaluators_run_unmodified_on_the_dropins(shadowed_reference, tmp_path):
    rep_mod, mha_mod = shadowed_reference
    cfg = synth.CONFIGS['cfg2']
    params = {'extractor_params': cfg.extractor_params,
              'matcher_params': {'type': 'brute_force', 'brute_force_params': cfg.matcher_params},
              'repeatability_params': {'th': 3, 'image': None, 'output': str(tmp_path) + '/'},
              'MHA_params': {'th': [3, 5, 7]}}
    for i in range(2):
        pair = synth.make_pair(cfg, 2, 40 + i)
        img = torch.zeros(1, 3, cfg.height, cfg.width, device=DEV)     # the batch's image lives where the maps do (MHA.py:41-42)
        s0, s1 = pair['score0'].to(DEV), pair['score1'].to(DEV)
        d0, d1 = pair['desc0'].to(DEV), pair['desc1'].to(DEV)
        w01, w10 = pair['warp01'], pair['warp10']                     # 0-dim int64 tensors, as the collate leaves them
        k0, _ = ref_ops.detection(pair['score0'], cfg.extractor_params, nms='greedy')
        k1, _ = ref_ops.detection(pair['score1'], cfg.extractor_params, nms='greedy')
        # ---- tasks/repeatability.py:95-122, whole function (PNG dumps included)
        got = rep_mod.repeatability(i, img, s0, img, s1, w01, w10, params)
        ora = ref_ops.val_key_points(k0, k1, w01, w10, th=3)
        assert int(got['num_feat']) == ora['num_feat']
        # the reference's own torch arithmetic (torch.norm on the GPU) vs numpy: the 2^-7 bucket of mutual_argmin
        # (repeatability.py:18-32) can move a single pair, see oracle/compare.py:explain_repeat_pair_diffs
        assert abs(float(got['repeatability']) * ora['num_feat'] - ora['gt_num']) <= 1.0 + 1e-6
        assert abs(float(got['mean_error']) - ora['mean_error']) < 2e-3
        assert np.allclose(got['errors'].cpu().numpy(), ora['errors'], rtol=1e-5, atol=1e-5 * 512)
        assert os.path.isfile(str(tmp_path / f'{i}_repeatability_0.png'))
        # ---- tasks/MHA.py:11-72, whole function: consumption order into cv2.findHomography, 0-dim width / height,
        #      img_0.device and img_0.shape uses
        flags = mha_mod.mha(i, img, s0, d0, img, s1, d1, w01, w10, params)
        want = ref_ops.mha_pair(pair['score0'], pair['desc0'].numpy(), pair['score1'], pair['desc1'].numpy(), w01, w10,
                                params, (cfg.height, cfg.width))
        want_flags = want[0] if isinstance(want, tuple) else want
        assert flags == list(want_flags), (flags, want_flags)
        assert flags == [1.0, 1.0, 1.0]
