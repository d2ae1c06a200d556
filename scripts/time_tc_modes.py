"""Timing experiments on the tensor-core matcher main kernel (kb_debug_knob(KB_KNOB_TC_DEBUG, mode); results are wrong in modes != 0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import _lib, ops
from torch.profiler import ProfilerActivity, profile
shapes = [(64, 1000, 1000, 256), (16, 4096, 4096, 64), (8, 2048, 2048, 128)]
g = torch.Generator().manual_seed(1)
for B, n, m, D in shapes:
    a = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=2).cuda()
    b = torch.nn.functional.normalize(torch.randn(B, m, D, generator=g), dim=2).cuda()
    b[:, :n // 2] = a[:, :n // 2] + 0.05 * torch.randn(B, n // 2, D, generator=g).cuda()
    for mode in (0, 1, 2):
        _lib.lib.kb_debug_knob(_lib.KB_KNOB_TC_DEBUG, mode)
        for _ in range(2):
            ops.match_batched(a, b, None, None, 5.0, True, algo=1)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                ops.match_batched(a, b, None, None, 5.0, True, algo=1)
            torch.cuda.synchronize()
        t = [e.device_time_total / 3 for e in prof.key_averages() if 'nn_top2' in e.key]
        print(f'B={B} n={n} D={D} mode={mode}: nn_top2 {t[0]:.1f} us', flush=True)
_lib.lib.kb_debug_knob(_lib.KB_KNOB_TC_DEBUG, 0)
