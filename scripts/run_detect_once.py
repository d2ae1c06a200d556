"""Warm + one more batched detection call (driven under ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops, synth

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'cfg2']
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g = torch.Generator(device='cuda').manual_seed(3)
s = torch.rand(n, 1, cfg.height, cfg.width, generator=g, device='cuda')
for _ in range(2):
    xyp, count, raster, path = ops.detect_batched(s, cfg.extractor_params)
torch.cuda.synchronize()
print(int(count.sum()), torch.bincount(path.cpu(), minlength=3).tolist())
