"""A/B of the plane-staged sampler (four channels per CTA) against the eight-channel double-buffered form
(the default; kb_debug_knob(KB_KNOB_SAMPLE_4CH, 1) selects the older kernel) at the cfg2 shape: equal bits, CUDA-event times."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops
from keypoint_bench_b200._lib import lib


def time_ms(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device='cuda').manual_seed(3)
B = 128
d = torch.randn(B, 256, 60, 80, generator=g, device='cuda')
pts = torch.rand(B, 1000, 3, generator=g, device='cuda')
cnt = torch.randint(950, 1001, (B,), generator=g, device='cuda').to(torch.int32)
out = {}
with ops.no_zero_fill():
    got = ops.sample_batched(d, pts, cnt).clone()
    out['eight_ch_ms'] = time_ms(lambda: ops.sample_batched(d, pts, cnt))
    lib.kb_debug_knob(7, 1)                                    # KB_KNOB_SAMPLE_4CH
    ref = ops.sample_batched(d, pts, cnt).clone()
    out['four_ch_ms'] = time_ms(lambda: ops.sample_batched(d, pts, cnt))
    lib.kb_debug_knob(7, 0)
    out['equal'] = all(bool(torch.equal(ref[b, :int(cnt[b])], got[b, :int(cnt[b])])) for b in range(B))
print(json.dumps(out))
