"""A/B timing of the round-1 kernels of the sparse detection path (tiled / fp32 streaming / packed streaming) and of the whole
detect stage, CUDA events on the launch stream.  python scripts/ab_round1.py [cfg2|cfg3|cfg4|cfg5] [maps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops, synth  # noqa: E402


def time_ms(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    cfg = synth.CONFIGS[name]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * cfg.pairs_per_gpu
    g = torch.Generator(device='cuda').manual_seed(1)
    s = torch.rand(n, 1, cfg.height, cfg.width, generator=g, device='cuda')
    out = {'config': name, 'maps': n, 'bytes': s.numel() * 4}
    with ops.no_zero_fill():
        for tag, bit in (('tiled', 8), ('stream', 16), ('packed', 32)):
            st = []
            ops.detect_batched(s, cfg.extractor_params, phases=7 | bit, state=st)
            out[f'round1_{tag}_ms'] = time_ms(lambda: ops.detect_batched(s, cfg.extractor_params, phases=2 | bit, state=st))
            # (repeated round-1 launches alone keep growing the list counters, so later repeats skip their stores:
            # threshold estimate + round 1 together, minus the estimate, is the unbiased figure)
            out[f'tau_round1_{tag}_ms'] = time_ms(lambda: ops.detect_batched(s, cfg.extractor_params, phases=3 | bit, state=st))
            ops.detect_batched(s, cfg.extractor_params, phases=7 | bit, state=st)
            out[f'resolve_after_{tag}_ms'] = time_ms(lambda: ops.detect_batched(s, cfg.extractor_params, phases=4, state=st))
            out[f'detect_{tag}_ms'] = time_ms(lambda: ops.detect_batched(s, cfg.extractor_params, phases=7 | bit, state=st))
            out[f'round1_{tag}_gbs'] = out['bytes'] / out[f'round1_{tag}_ms'] / 1e6
        st = []
        ops.detect_batched(s, cfg.extractor_params, phases=7, state=st)
        out['tau_ms'] = time_ms(lambda: ops.detect_batched(s, cfg.extractor_params, phases=1, state=st))
        out['detect_auto_ms'] = time_ms(lambda: ops.detect_batched(s, cfg.extractor_params, phases=7, state=st))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
