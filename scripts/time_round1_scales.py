"""Packed round-1 kernel on maps whose scores live far from 1 (the per-map scale of its 16-bit image): CUDA-event times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops, synth

cfg = synth.CONFIGS['cfg2']
g = torch.Generator(device='cuda').manual_seed(1)
base = torch.rand(128, 1, 480, 640, generator=g, device='cuda')
for scale in (1e-12, 1e-6, 1.0, 1e5, 1e12):
    s = base * scale
    st = []
    with ops.no_zero_fill():
        ops.detect_batched(s, cfg.extractor_params, phases=7 | 32, state=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.detect_batched(s, cfg.extractor_params, phases=3 | 32, state=st)
        e1.record()
        torch.cuda.synchronize()
    print(f'scale {scale:g}: tau + round 1 {e0.elapsed_time(e1) / 10 * 1000:.1f} us')
