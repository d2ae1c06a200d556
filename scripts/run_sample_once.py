import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops
g = torch.Generator(device='cuda').manual_seed(3)
d = torch.randn(128, 256, 60, 80, generator=g, device='cuda')
p = torch.rand(128, 1000, 2, generator=g, device='cuda')
cnt = torch.full((128,), 950, dtype=torch.int32, device='cuda')
for _ in range(2):
    out = ops.sample_batched(d, p, cnt)
torch.cuda.synchronize()
print(float(out.abs().sum()))
