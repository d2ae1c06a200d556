"""Tiled vs packed round-1 kernel on other kinds of score maps (synth.score_map): threshold estimate + round 1, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops, synth

cfg = synth.CONFIGS['cfg2']
for kind in ('uniform', 'alike', 'relu', 'ties'):
    base = torch.cat([synth.score_map(kind, cfg.height, cfg.width, 100 + i, 'cuda') for i in range(16)])
    s = base.repeat(8, 1, 1, 1).contiguous()
    row = {}
    for tag, bit in (('tiled', 8), ('packed', 32)):
        st = []
        with ops.no_zero_fill():
            out = ops.detect_batched(s, cfg.extractor_params, phases=7 | bit, state=st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.detect_batched(s, cfg.extractor_params, phases=3 | bit, state=st)
            e1.record()
            torch.cuda.synchronize()
            row[tag] = e0.elapsed_time(e1) / 10 * 1000
            e0.record()
            for _ in range(10):
                ops.detect_batched(s, cfg.extractor_params, phases=7 | bit, state=st)
            e1.record()
            torch.cuda.synchronize()
            row[tag + '_detect'] = e0.elapsed_time(e1) / 10 * 1000
        row[tag + '_paths'] = torch.bincount(out[3].cpu(), minlength=3).tolist()
    print(kind, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in row.items()})
