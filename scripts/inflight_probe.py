"""Does keeping two steps in flight (one CUDA graph per slot, own stream, own batch) raise step throughput?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from keypoint_bench_b200 import pipeline, synth

cfg_name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = synth.CONFIGS[cfg_name]
P = cfg.pairs_per_gpu
dev = torch.device('cuda')
batches = [bench.make_batch(cfg, int(cfg_name[3:]), P, 1000 * s, dev, 'uniform')[0] for s in range(depth)]


def mk(b):
    def step():
        if cfg.desc_dim == 0:
            res = pipeline.repeatability_counts(b, cfg, 3.0)
            return res, pipeline.accumulate_repeatability(res)
        res = pipeline.extract_match(b, cfg)
        return res, pipeline.accumulate_matches(res)
    return step


steps = [mk(b) for b in batches]
for s in steps:
    for _ in range(3):
        s()
torch.cuda.synchronize()
graphs = [pipeline.GraphedStep(s) for s in steps]
streams = [torch.cuda.Stream() for _ in range(depth)]
K = 40


def run(mode):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    e0.record()
    if mode == 'serial':
        for i in range(K):
            graphs[i % depth]()
    else:
        for s in streams:
            s.wait_stream(cur)
        for i in range(K):
            with torch.cuda.stream(streams[i % depth]):
                graphs[i % depth]()
        for s in streams:
            cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


for mode in ('serial', 'inflight', 'serial', 'inflight'):
    ms = run(mode)
    print(f'{cfg.name} {mode:9s} depth={depth}: {ms:.4f} ms/step  {P / ms * 1e3:,.0f} pairs/s')
