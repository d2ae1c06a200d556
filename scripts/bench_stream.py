"""cfg5 as a frame stream: F + 1 consecutive frames per step, every frame detected and sampled once,
matched against its predecessor -- against the pairwise form of the same config (both frames of every pair
extracted, what bench.py --config cfg5 times).  CUDA events, graphs, `depth` steps in flight, one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import pipeline, synth  # noqa: E402

cfg = synth.CONFIGS['cfg5']
F = int(sys.argv[1]) if len(sys.argv) > 1 else cfg.pairs_per_gpu
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
K = 30
dev = torch.device('cuda')


def make_frames(seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    score = torch.rand(F + 1, 1, cfg.height, cfg.width, generator=g, device=dev)
    desc = 2.67 * torch.nn.functional.normalize(
        torch.randn(F + 1, cfg.desc_dim, cfg.height, cfg.width, generator=g, device=dev), dim=1)
    for t in range(1, F + 1):           # camera pans 3 px per frame; new content enters on the left
        score[t, :, :, 3:] = score[t - 1, :, :, :-3]
        desc[t, :, :, 3:] = desc[t - 1, :, :, :-3] + 0.05 * torch.randn(cfg.desc_dim, cfg.height, cfg.width - 3,
                                                                       generator=g, device=dev)
    return pipeline.FrameBatch(score, desc)


def timed(fns):
    flight = pipeline.StepsInFlight(fns)
    def run(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flight.fork()
        for i in range(n):
            flight.launch(i)
        flight.join()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    run(2 * depth)
    return run(K), flight.slots[0].out


frames = [make_frames(100 + s) for s in range(depth)]
ms_stream, out = timed([(lambda f=f: pipeline.extract_match_stream(f, cfg)) for f in frames])
eye = torch.eye(3, device=dev).reshape(1, 9).repeat(2 * F, 1)
wh = torch.tensor([[float(cfg.width), float(cfg.height)]], device=dev).repeat(2 * F, 1)
pbs = [pipeline.PairBatch(torch.cat([f.score[:-1], f.score[1:]]), torch.cat([f.desc[:-1], f.desc[1:]]), eye, wh)
       for f in frames]
ms_pair, out_p = timed([(lambda b=b: pipeline.extract_match(b, cfg, covisible_only=False)) for b in pbs])
same = bool(torch.equal(out['n_matches'], out_p['n_matches']))
print(json.dumps({'workload': cfg.name, 'pairs_per_step': F, 'steps_in_flight': depth,
                  'stream': {'ms_per_step': ms_stream, 'pairs_per_s': F / ms_stream * 1e3},
                  'pairwise': {'ms_per_step': ms_pair, 'pairs_per_s': F / ms_pair * 1e3},
                  'mean_matches_per_pair': float(out['n_matches'].float().mean()), 'same_match_counts': same}))
