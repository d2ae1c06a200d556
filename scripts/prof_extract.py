import os, sys, torch
sys.path.insert(0, '/root/repo')
from keypoint_bench_b200.utils import lightglue_extract as lg
from torch.profiler import ProfilerActivity, profile
g = torch.Generator(device='cuda').manual_seed(1)
score = torch.rand(128, 1, 480, 640, generator=g, device='cuda')
desc = torch.randn(128, 256, 60, 80, generator=g, device='cuda')
for _ in range(3): lg.extract_batched(score, desc, 8)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): lg.extract_batched(score, desc, 8)
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f'{e.key[:70]:70s} {e.device_time_total/3:9.1f} us x{e.count/3:.0f}')
