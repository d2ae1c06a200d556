"""Where does sparse_kernel (one CTA per map) spend its time?  clock64 at the phase boundaries of map 0's CTA
(KB_KNOB_SPARSE_PROF).  python scripts/sparse_phase_profile.py [cfg1..cfg5]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import _lib, ops, synth  # noqa: E402

names = ['cut', 'load', 'grid', 'decide', 'compact', 'sort', 'emit']
for name in (sys.argv[1:] or ['cfg2', 'cfg3', 'cfg4', 'cfg5']):
    cfg = synth.CONFIGS[name]
    n = 2 * cfg.pairs_per_gpu
    s = torch.rand(n, 1, cfg.height, cfg.width, device='cuda')
    with ops.debug_knob(_lib.KB_KNOB_SPARSE_PROF, 1):
        for _ in range(3):
            ops.detect_batched(s, cfg.extractor_params)
        torch.cuda.synchronize()
        out = (ctypes.c_longlong * 16)()
        _lib.check(_lib.lib.kb_debug_sparse_prof(ctypes.cast(out, ctypes.c_void_p)), 'prof')
    t = list(out)
    d = {nm: t[i + 1] - t[i] for i, nm in enumerate(names)}
    print(json.dumps({'config': name, 'maps': n, 'cycles': d, 'total': t[7] - t[0], 'candidates': t[8], 'kept_interior': t[9],
                      'listM': t[10], 'listO': t[11], 'attempt': t[12], 'decide_list_build_cycles': t[13], 'bands': t[14], 'max_rounds_any_warp_in_a_band': t[15]}))
