"""Per-kernel device time of one bench step (CUPTI via torch.profiler): shares, not bench numbers."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from keypoint_bench_b200 import pipeline, synth
from torch.profiler import ProfilerActivity, profile

cfg_name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
kind = sys.argv[2] if len(sys.argv) > 2 else 'uniform'
cfg = synth.CONFIGS[cfg_name]
P = int(sys.argv[3]) if len(sys.argv) > 3 else cfg.pairs_per_gpu
batch, hms = bench.make_batch(cfg, int(cfg_name[3:]), P, 0, torch.device('cuda'), kind)


def step():
    if cfg.desc_dim == 0:
        return pipeline.repeatability_counts(batch, cfg, 3.0)
    return pipeline.extract_match(batch, cfg)


for _ in range(3):
    r = step()
torch.cuda.synchronize()
iters = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / iters, e.count / iters) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[1] for r in rows)
print(f'{cfg.name} kind={kind} pairs={P}: {tot:.1f} us of kernels per step')
for k, t, c in sorted(rows, key=lambda r: -r[1]):
    print(f'  {k[:80]:80s} {t:9.1f} us  x{c:.0f}  {100 * t / tot:5.1f}%')
if 'path' in r:
    print('  detect paths:', torch.bincount(r['path'].flatten().cpu(), minlength=3).tolist())
