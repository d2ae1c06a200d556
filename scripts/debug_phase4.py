"""Why does detect phases=4 alone take milliseconds?  Prints path / counts after phase-only calls."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops, synth
cfg = synth.CONFIGS['cfg2']
s = torch.rand(128, 1, 480, 640, device='cuda')
st = []
xyp, count, raster, path = ops.detect_batched(s, cfg.extractor_params, state=st)
torch.cuda.synchronize()
print('full', path[:8].tolist(), count[:4].tolist())
for ph in (4, 4, 2, 4, 1, 4, 7, 4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    xyp, count, raster, path = ops.detect_batched(s, cfg.extractor_params, phases=ph, state=st)
    e1.record()
    torch.cuda.synchronize()
    print('phases', ph, 'ms', round(e0.elapsed_time(e1), 3), path[:8].tolist(), count[:4].tolist())
