"""CUDA-event time of the matcher's operand preparation alone at the cfg2 shape (64 pairs x ~975 x 256)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops

g = torch.Generator(device='cuda').manual_seed(3)
a = torch.randn(64, 1000, 256, generator=g, device='cuda')
b = torch.randn(64, 1000, 256, generator=g, device='cuda')
n = torch.randint(950, 1001, (64,), generator=g, device='cuda').to(torch.int32)
st = []
ops.match_batched(a, b, n, n, 5.0, True, algo=1, state=st, want_dist=False)
for ph, name in ((1, 'prep'), (2, 'search'), (4, 'tail')):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ops.match_batched(a, b, n, n, 5.0, True, algo=1, phases=ph, state=st, want_dist=False)
    e0.record()
    for _ in range(10):
        ops.match_batched(a, b, n, n, 5.0, True, algo=1, phases=ph, state=st, want_dist=False)
    e1.record()
    torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 10 * 1000, 1), 'us')
