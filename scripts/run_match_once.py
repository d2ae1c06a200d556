"""One warm call + one profiled call of the matcher per shape (driven under ncu for launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops

algo = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shapes = [(64, 1000, 1000, 256), (16, 4096, 4096, 64), (8, 2048, 2048, 128)]
g = torch.Generator().manual_seed(1)
for B, n, m, D in shapes:
    a = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=2).cuda()
    b = torch.nn.functional.normalize(torch.randn(B, m, D, generator=g), dim=2).cuda()
    b[:, :n // 2] = a[:, :n // 2] + 0.05 * torch.randn(B, n // 2, D, generator=g).cuda()
    for _ in range(2):
        p, d, c = ops.match_batched(a, b, None, None, 5.0, True, algo=algo)
    torch.cuda.synchronize()
    print(B, n, m, D, int(c.sum()))
