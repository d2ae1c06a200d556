"""Warm + one more fused sampling / matching call at the cfg2 shape (driven under ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops

P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device='cuda').manual_seed(3)
d = torch.randn(2 * P, 256, 60, 80, generator=g, device='cuda')
pts = torch.rand(2 * P, 1000, 3, generator=g, device='cuda')
cnt = torch.randint(950, 1001, (2 * P,), generator=g, device='cuda').to(torch.int32)
for fused in (True, False):
    for _ in range(2):
        out = ops.sample_match_batched(d, pts, cnt, P, 5.0, True, want_dist=False, fused=fused)
torch.cuda.synchronize()
print(int(out[3].sum()))
