import ctypes, json, os, sys
import torch
sys.path.insert(0, '/root/repo')
from keypoint_bench_b200 import _lib, ops, synth
names = ['cut', 'load', 'grid', 'decide', 'compact', 'sort', 'emit']
cfg = synth.CONFIGS['cfg2']
for kind in ('uniform', 'alike'):
    for seed in (100, 101, 102):
        s = torch.cat([synth.score_map(kind, cfg.height, cfg.width, seed + i, 'cuda') for i in range(4)])
        with ops.debug_knob(_lib.KB_KNOB_SPARSE_PROF, 1):
            for _ in range(3):
                out4 = ops.detect_batched(s, cfg.extractor_params)
            torch.cuda.synchronize()
            out = (ctypes.c_longlong * 16)()
            _lib.check(_lib.lib.kb_debug_sparse_prof(ctypes.cast(out, ctypes.c_void_p)), 'prof')
        t = list(out)
        d = {nm: t[i + 1] - t[i] for i, nm in enumerate(names)}
        print(kind, seed, json.dumps({'cycles': d, 'total': t[7] - t[0], 'candidates': t[8], 'kept_interior': t[9], 'listM': t[10], 'listO': t[11], 'attempt': t[12], 'bands': t[14], 'max_rounds': t[15]}))
