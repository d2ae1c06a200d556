"""Timing of the SURVEY 8(f) rows (LightGlue-style extract, tensor Lucas-Kanade matcher, SE(3) warp) at reference
sizes on one GPU: CUDA events on the launch stream, 3 warm-ups, L2 flushed between iterations.  Prints one JSON
line per row; `--cpu` adds the oracle port timed on a bounded sample (host cores, numpy/torch)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops, synth  # noqa: E402
from keypoint_bench_b200.utils import lightglue_extract as lg  # noqa: E402


def timed(fn, iters=10, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cpu', action='store_true')
    ap.add_argument('--maps', type=int, default=128)
    ap.add_argument('--lk-pairs', type=int, default=8)
    args = ap.parse_args()
    dev = 'cuda'
    rows = []

    # LightGlue-style extract: 480x640 score maps, radius 5, 1000 keypoints, 256-d descriptors at 1/8 resolution
    g = torch.Generator(device=dev).manual_seed(1)
    score = torch.rand(args.maps, 1, 480, 640, generator=g, device=dev)
    desc = torch.randn(args.maps, 256, 60, 80, generator=g, device=dev)
    ms_nms = timed(lambda: ops.simple_nms_batched(score, 5))
    ms_all = timed(lambda: lg.extract_batched(score, desc, 8))
    from keypoint_bench_b200 import pipeline
    graphed = pipeline.GraphedStep(lambda: lg.extract_batched(score, desc, 8))      # the same launches as one CUDA graph
    ms_graph = timed(graphed)
    row = {'row': 'lightglue extract (simple_nms r=5 + top-1000 + normalised sampling)', 'maps': args.maps,
           'ms': ms_all, 'maps_per_s': args.maps / ms_all * 1e3, 'ms_cuda_graph': ms_graph,
           'maps_per_s_cuda_graph': args.maps / ms_graph * 1e3, 'simple_nms_ms': ms_nms,
           'simple_nms_GBps_algorithmic': args.maps * 480 * 640 * 8 / ms_nms / 1e6}
    if args.cpu:
        from oracle import ref_ops
        t0 = time.time()
        for i in range(2):
            ref_ops.lightglue_extract(score[i].cpu().numpy(), desc[i].cpu().numpy(), 8)
        row['cpu_port_maps_per_s'] = 2 / (time.time() - t0)
    rows.append(row)

    # Tensor Lucas-Kanade matcher: config_fund.yaml:72-77 (win 21, 3 levels, 40 iterations), 1000 keypoints per pair
    P, n = args.lk_pairs, 1000
    scenes = [synth.lk_scene(3, 480, 640, 40 + i, shift=(2.0 + 0.3 * i, -1.5)) for i in range(P)]
    img0 = torch.cat([s[0] for s in scenes], 0).to(dev)
    img1 = torch.cat([s[1] for s in scenes], 0).to(dev)
    pts = torch.rand(P, n, 2, generator=g, device=dev) * torch.tensor([639.0, 479.0], device=dev)
    init = pts + torch.randn(P, n, 2, generator=g, device=dev) * 3
    ms_lk = timed(lambda: ops.lk_track_batched(img0, img1, pts, init, None, 21, 3, 40), iters=5)
    taps = P * n * 3 * 40 * 3 * 441 * 3 * 4
    row = {'row': 'optical_flow_tensor (win 21, 3 levels x 40 iterations, 1000 keypoints)', 'pairs': P, 'ms': ms_lk,
           'pairs_per_s': P / ms_lk * 1e3, 'G_taps_per_s': taps / ms_lk / 1e6}
    if args.cpu:
        from oracle import ref_ops
        k = 50
        t0 = time.time()
        ref_ops.lk_track(img0[0].cpu().numpy(), img1[0].cpu().numpy(), pts[0, :k].cpu().numpy(), init[0, :k].cpu().numpy(),
                         21, 3, 40)
        row['cpu_port_pairs_per_s'] = (k / n) / (time.time() - t0)
        row['cpu_sample'] = f'{k} of {n} keypoints of one pair through oracle/ref_ops.lk_track'
    rows.append(row)

    # SE(3) warp: 1000 keypoints per pair, 480x640 depth maps
    B = 64
    sc = synth.se3_scene(480, 640, 3, device=dev)
    kp = torch.rand(B, 1000, 3, generator=g, device=dev)
    rep = lambda t: t[None].expand(B, *t.shape).contiguous()    # noqa: E731
    args_se3 = (kp, None, rep(sc['depth0']), rep(sc['depth1']), rep(sc['intrinsics0']), rep(sc['intrinsics1']),
                rep(sc['pose01']), rep(sc['bbox0']), rep(sc['bbox1']))
    try:
        ms_se3 = timed(lambda: ops.warp_se3_batched(*args_se3))
        rows.append({'row': 'warp_se3 (1000 keypoints, 480x640 depth)', 'pairs': B, 'ms': ms_se3,
                     'pairs_per_s': B / ms_se3 * 1e3})
    except Exception as e:      # signature drift must not hide the other rows
        rows.append({'row': 'warp_se3', 'error': repr(e)})
    for r in rows:
        print(json.dumps(r))


if __name__ == '__main__':
    main()
