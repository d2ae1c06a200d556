#!/usr/bin/env python
"""Benchmark of the post-network hot path (BASELINE.json metric: image pairs/sec, extract+match).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

A "step" is one pass of the hot path over one batch of synthetic image pairs per GPU:
detection x2 -> covisibility warp x2 -> descriptor sampling x2 -> mutual-NN matching
(tasks/MHA.py:29-39 up to, not including, the host-side cv2 RANSAC).  N>1 = one process per GPU
under torchrun; pairs are independent so every rank processes its own batch (weak scaling) and the
only collective is the all-reduce of the count vector at the end of the timed region.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'image pairs/sec (extract+match)'
UNIT = 'pairs/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg2', help='cfg1..cfg5 (BASELINE.json configs); cfg2 is the headline')
    ap.add_argument('--pairs', type=int, default=0, help='pairs per GPU per step (0 = config default)')
    ap.add_argument('--algo', type=int, default=-1, help='matcher: 0 = float64 SIMT, 1 = tcgen05; -1 = best available')
    ap.add_argument('--kind', default='uniform', help='synthetic score-map kind (uniform | alike)')
    ap.add_argument('--cpu-pairs', type=int, default=2, help='pairs timed on the host for cpu_baseline (0 = skip)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--in-flight', type=int, default=3,
                    help='steps kept in flight (one CUDA graph + stream + batch per slot); 1 = strictly serial steps')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying a CUDA graph')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic batch (device)
# ------------------------------------------------------------------------------------------------

def make_batch(cfg, cfg_index, n_pairs, first_pair, device, kind='uniform'):
    """P pairs resident on `device`, generated per SURVEY 8(d): score A, score B = nearest-warp of A,
    unit-norm descriptor map A, map B = bilinear warp of A + 0.05 noise."""
    from keypoint_bench_b200 import synth
    from keypoint_bench_b200.pipeline import PairBatch
    H, W = cfg.height, cfg.width
    g = torch.Generator(device=device)
    g.manual_seed(synth.pair_seed(cfg_index, first_pair))
    hms = torch.stack([synth.homography(synth.pair_seed(cfg_index, first_pair + i) + 7) for i in range(n_pairs)])
    if kind == 'uniform':
        s0 = torch.rand(n_pairs, 1, H, W, generator=g, device=device)
    else:
        s0 = torch.cat([synth.score_map(kind, H, W, synth.pair_seed(cfg_index, first_pair + i), device)
                        for i in range(n_pairs)])
    s1 = torch.empty_like(s0)
    chunk = 8
    for i in range(0, n_pairs, chunk):
        grids = torch.cat([synth._inverse_grid(hms[j], H, W, device) for j in range(i, min(i + chunk, n_pairs))])
        s1[i:i + chunk] = torch.nn.functional.grid_sample(s0[i:i + chunk], grids, mode='nearest',
                                                          padding_mode='zeros', align_corners=True)
    desc = None
    if cfg.desc_dim:
        dh, dw = H // cfg.desc_stride, W // cfg.desc_stride
        d0 = torch.nn.functional.normalize(torch.randn(n_pairs, cfg.desc_dim, dh, dw, generator=g, device=device), dim=1)
        if not cfg.desc_normalized:
            d0 = 2.67 * d0
        d1 = torch.empty_like(d0)
        for i in range(0, n_pairs, chunk):
            grids = torch.cat([synth._inverse_grid(synth.rescale_homography(hms[j], cfg.desc_stride), dh, dw, device)
                               for j in range(i, min(i + chunk, n_pairs))])
            d1[i:i + chunk] = torch.nn.functional.grid_sample(d0[i:i + chunk], grids, mode='bilinear',
                                                              padding_mode='zeros', align_corners=True)
        d1 += 0.05 * torch.randn(d1.shape, generator=g, device=device)
        desc = torch.cat([d0, d1])
    h01 = hms.reshape(n_pairs, 9)
    h10 = torch.linalg.inv(hms.double()).float().reshape(n_pairs, 9)
    h33 = torch.cat([h01, h10]).to(device)
    wh = torch.tensor([[float(W), float(H)]], device=device).expand(2 * n_pairs, 2).contiguous()
    return PairBatch(score=torch.cat([s0, s1]), desc=desc, h33=h33, wh=wh, resize=512), hms


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the benchmark runs (B200_PROFILING.md)."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.proc = None
        self.samples = []
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith('GPU-'):
                uuid = 'GPU-' + uuid
            self.proc = subprocess.Popen(['nvidia-smi', '-i', uuid, f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'power_w_max': max(pw),
                'samples': len(sm), 'reasons': sorted(reasons)}


class StageTimer:
    """CUDA-event timing of the pipeline stages on the launching (current) stream."""
    def __init__(self):
        self.events = []      # (name, event)
        self.cur = None

    def __call__(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.events.append((name, ev))

    def totals(self):
        tot = {}
        for (n0, e0), (n1, e1) in zip(self.events[:-1], self.events[1:]):
            if n0 is None:
                continue
            tot[n0] = tot.get(n0, 0.0) + e0.elapsed_time(e1)
        return tot


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's algorithm (oracle port) on the host
# ------------------------------------------------------------------------------------------------

def cpu_pair(cfg, score0, score1, desc0, desc1, hm):
    """One pair through the oracle exactly as the reference would run it on CPU: detection with the
    im2col/argmax/col2im rounds (utils/extracter.py:49-98), warp, grid_sample, float64 cdist matcher."""
    from keypoint_bench_b200 import synth
    from oracle import ref_ops
    H, W = cfg.height, cfg.width
    w01, w10 = synth.warp_params(hm, H, W)
    k0, _ = ref_ops.detection(score0, cfg.extractor_params, nms='im2col')
    k1, _ = ref_ops.detection(score1, cfg.extractor_params, nms='im2col')
    k0c, _, ids0, _ = ref_ops.warp(k0, w01)
    k1c, _, ids1, _ = ref_ops.warp(k1, w10)
    pairs = None
    if desc0 is not None and k0c.shape[0] and k1c.shape[0]:
        _, _, pairs = ref_ops.brute_force_matcher(k0c, k1c, desc0, desc1, cfg.matcher_params)
    return k0, k1, k0c, k1c, pairs


def run_cpu(cfg, batch_cpu, hms, n_pairs):
    torch.set_num_threads(os.cpu_count() or 1)
    P = batch_cpu['score'].shape[0] // 2
    t0 = time.perf_counter()
    results = []
    for i in range(n_pairs):
        d0 = batch_cpu['desc'][i:i + 1].numpy() if batch_cpu['desc'] is not None else None
        d1 = batch_cpu['desc'][P + i:P + i + 1].numpy() if batch_cpu['desc'] is not None else None
        results.append(cpu_pair(cfg, batch_cpu['score'][i:i + 1], batch_cpu['score'][P + i:P + i + 1], d0, d1, hms[i]))
    dt = time.perf_counter() - t0
    return dt, results


def main():
    args = parse_args()
    from keypoint_bench_b200 import synth
    cfg = synth.CONFIGS[args.config]
    cfg_index = int(args.config[3:])
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    P = args.pairs or cfg.pairs_per_gpu
    config = {'workload': cfg.name, 'pairs_per_gpu_per_step': P, 'height': cfg.height, 'width': cfg.width,
              'desc_dim': cfg.desc_dim, 'desc_stride': cfg.desc_stride, 'nms_dist': cfg.nms_dist, 'top_k': cfg.top_k,
              'border_dist': cfg.border_dist, 'max_distance': cfg.max_distance, 'cross_check': cfg.cross_check,
              'score_map': args.kind, 'sharding': f'by-pair x{world}',
              'timed_stages': 'detect x2, warp x2, sample x2, mutual-NN match (host cv2 RANSAC excluded)',
              'l2': 'working set per step > 126 MB L2 (no flush needed)'}

    # -------------------------------------------------------------------------------- reference arm
    if args.impl == 'reference':
        if rank != 0:
            return 0
        dev = 'cpu'
        batch, hms = make_batch(cfg, cfg_index, 1, 0, dev, args.kind)
        bc = {'score': batch.score, 'desc': batch.desc}
        for _ in range(min(args.warmup, 1)):                     # one warm pair is enough to page in torch
            run_cpu(cfg, bc, hms, 1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run_cpu(cfg, bc, hms, 1)
        dt = time.perf_counter() - t0
        value = args.steps / dt
        line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000 * dt / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32+f64',
                'data': 'synthetic', 'config': dict(config, pairs_per_gpu_per_step=1),
                'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                                 'sample': f'{args.steps} steps x 1 pair of {cfg.name} through oracle/ref_ops.py '
                                           f'(im2col NMS rounds, grid_sample, scipy cdist float64)'},
                'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return 0

    # -------------------------------------------------------------------------------- our arm
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the product path)')
    from keypoint_bench_b200 import ops, parallel, pipeline
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        parallel.init('nccl')
    algo = args.algo
    tc = algo == 1 or (algo < 0 and cfg.desc_dim <= 256)
    config['matcher'] = 'tcgen05 split-bf16 Gram + float64 certify' if tc else 'float64 SIMT'

    batch, hms = make_batch(cfg, cfg_index, P, rank * P, device, args.kind)
    task_rep = cfg.desc_dim == 0

    def step(timer=None, b=None):
        b = batch if b is None else b
        if task_rep:
            res = pipeline.repeatability_counts(b, cfg, 3.0, timer)
            return res, pipeline.accumulate_repeatability(res)
        res = pipeline.extract_match(b, cfg, algo=algo, timer=timer)
        return res, pipeline.accumulate_matches(res)

    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    # per-stage device times: every stage alone, 5 back-to-back calls between two CUDA events on the launch
    # stream (the stage's kernels queue up behind each other, so host launch gaps do not count)
    l0 = ops.launches()
    res0, _ = step()
    launches_per_step = ops.launches() - l0
    torch.cuda.synchronize()

    def time_stage(fn, reps=5):
        # like the pipeline (pipeline.extract_match), outputs are not zero-filled: the fill kernels of a
        # 131 MB descriptor buffer would otherwise be charged to the sampling stage
        with ops.no_zero_fill():
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    stage_ms = {'detect': time_stage(lambda: ops.detect_batched(batch.score, cfg.extractor_params)),
                'warp': time_stage(lambda: ops.warp_batched(res0['kpts'], res0['n_kpts'], batch.h33, batch.wh))}
    if task_rep:
        kv, kw, nv = res0['kcov'], res0['kwarp'], res0['n_cov']
        stage_ms['repeat'] = time_stage(lambda: ops.repeat_batched(kv[:P], kw[:P], nv[:P], kv[P:], kw[P:], nv[P:], 512.0,
                                                                   512.0, 3.0, want_errors=True))
    else:
        kv, nv, dd = res0['kcov'], res0['n_cov'], res0['desc']
        stage_ms['sample'] = time_stage(lambda: ops.sample_batched(batch.desc, kv, nv))
        stage_ms['match'] = time_stage(lambda: ops.match_batched(dd[:P], dd[P:], nv[:P], nv[P:], cfg.max_distance,
                                                                 cfg.cross_check, algo=algo, want_dist=False))
    # steady state: the launch sequence of one step replayed as a CUDA graph (same kernels, same buffers);
    # `--in-flight` slots, each with its own batch, graph and stream, so consecutive steps overlap
    depth = 1 if args.no_graph else max(1, args.in_flight)
    batches = [batch] + [make_batch(cfg, cfg_index, P, (rank + world * s) * P, device, args.kind)[0]
                         for s in range(1, depth)]
    graphed = flight = None
    if not args.no_graph:
        try:
            flight = pipeline.StepsInFlight([(lambda b=b: step(None, b)) for b in batches])
            graphed = flight.slots[0]
        except Exception as e:      # noqa: BLE001 -- fall back to eager launches, say so in the JSON line
            config['cuda_graph_error'] = repr(e)[:200]
            graphed = flight = None
            depth = 1
            torch.cuda.synchronize()
    config['launch'] = (f'cuda graph replay of the step, {depth} step(s) in flight on {depth} stream(s)'
                        if graphed is not None else 'eager launches from Python')
    config['steps_in_flight'] = depth
    run = graphed if graphed is not None else step

    def timed_steps(n_steps, in_flight):
        """n_steps steps between two events on the launch stream; returns (ms, accumulated counters)."""
        torch.cuda.synchronize()
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        accs = [None] * depth
        e0.record()
        if in_flight and flight is not None and depth > 1:
            def after(k, out):
                accs[k] = out[1].clone() if accs[k] is None else accs[k] + out[1]
            flight.fork()
            for i in range(n_steps):
                flight.launch(i, after=after)
            flight.join()
        else:
            for _ in range(n_steps):
                _, a = run()
                accs[0] = a.clone() if accs[0] is None else accs[0] + a
        tot = None
        for a in accs:
            if a is not None:
                a.record_stream(torch.cuda.current_stream())
                tot = a if tot is None else tot + a
        tot = parallel.reduce_counts(tot)             # the run's single collective
        e1.record()
        torch.cuda.synchronize()
        parallel.barrier()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=device)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            t = float(tt.item())
        return t, tot

    timed_steps(2 * depth, True)                      # graph / stream warm-up
    serial_ms = None
    if depth > 1:
        serial_ms, _ = timed_steps(args.steps, False)  # the same steps strictly one after another, for reference
    ms, acc = timed_steps(args.steps, True)
    value = world * P * args.steps / (ms / 1000.0)
    n_launch = launches_per_step * args.steps
    res, _ = run()                                     # slot 0 = `batch`: the outputs checked against the CPU port below
    torch.cuda.synchronize()

    # ---- end to end through the public API with HOST buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        # every slot: pinned host inputs -> device copy -> the step's graph -> device -> pinned host results, all on the
        # slot's stream; the host waits for a slot's previous results before it relaunches that slot (the caller
        # consumes the result of every step), so with two slots one step's copies overlap the other's kernels
        top = cfg.top_k
        in_bytes = batch.score.numel() * 4 + (batch.desc.numel() * 4 if batch.desc is not None else 0)
        slots = []
        for k, bk in enumerate(batches):
            share = k > 0 and in_bytes > (2 << 30)     # do not pin a second multi-GB host copy: reuse slot 0's
            sl = {'host_score': slots[0]['host_score'] if share else bk.score.cpu().pin_memory(),
                  'host_desc': None if bk.desc is None else (slots[0]['host_desc'] if share else bk.desc.cpu().pin_memory()),
                  'dev_score': torch.empty_like(bk.score),
                  'dev_desc': torch.empty_like(bk.desc) if bk.desc is not None else None,
                  'out_pairs': torch.empty((P, top, 2), dtype=torch.int32).pin_memory(),
                  'out_n': torch.empty((P,), dtype=torch.int32).pin_memory(),
                  'out_stats': torch.empty((P, 4), dtype=torch.float64).pin_memory(),
                  'done': torch.cuda.Event()}
            sl['batch'] = pipeline.PairBatch(sl['dev_score'], sl['dev_desc'], bk.h33, bk.wh, bk.resize)
            slots.append(sl)
        e2e_flight = None
        if flight is not None:
            try:
                e2e_flight = pipeline.StepsInFlight([(lambda b=sl['batch']: step(None, b)) for sl in slots])
            except Exception:       # noqa: BLE001
                e2e_flight = None
                torch.cuda.synchronize()

        def copy_in(k):
            sl = slots[k]
            sl['dev_score'].copy_(sl['host_score'], non_blocking=True)
            if sl['dev_desc'] is not None:
                sl['dev_desc'].copy_(sl['host_desc'], non_blocking=True)

        def copy_out(k, out):
            sl, r = slots[k], out[0]
            if task_rep:
                sl['out_stats'].copy_(r['stats'], non_blocking=True)
            else:
                sl['out_pairs'].copy_(r['matches'], non_blocking=True)
                sl['out_n'].copy_(r['n_matches'], non_blocking=True)
            sl['done'].record()

        def e2e_steps_run(n_steps):
            if e2e_flight is not None:
                e2e_flight.fork()
                for i in range(n_steps):
                    if i >= depth:
                        slots[i % depth]['done'].synchronize()      # results of this slot's previous step are consumed
                    e2e_flight.launch(i, before=copy_in, after=copy_out)
                e2e_flight.join()
                torch.cuda.synchronize()
            else:
                for _ in range(n_steps):
                    copy_in(0)
                    copy_out(0, step(None, slots[0]['batch']))
                    torch.cuda.synchronize()

        e2e_steps = max(3, min(args.steps, 10))
        e2e_steps_run(depth)
        parallel.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_steps_run(e2e_steps)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            dt = float(t.item())
        h2d = in_bytes
        d2h = slots[0]['out_stats'].numel() * 8 if task_rep else (slots[0]['out_pairs'].numel() * 4 +
                                                                 slots[0]['out_n'].numel() * 4)
        e2e = {'value': world * P * e2e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'steps': e2e_steps, 'steps_in_flight': depth if e2e_flight is not None else 1,
               'note': 'pinned host -> device copy of score+descriptor maps, hot path, device -> host copy of match '
                       'index pairs and counts, every step; the host waits for a slot\'s results before reusing the slot'}
    clocks = sampler.stop()

    if rank != 0:
        parallel.shutdown()
        return 0

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    pk = peaks()
    n_img = 2 * P
    n_out = float(res['n_kpts'].float().mean().item())
    # the detect stage is tau + round-1 + resolve: its streaming kernel (round1_kernel) is timed alone
    st = []
    ops.detect_batched(batch.score, cfg.extractor_params, state=st)
    round1_ms = time_stage(lambda: ops.detect_batched(batch.score, cfg.extractor_params, phases=2, state=st))
    kernels = {'round1_kernel (detect, streaming NMS round 1)': {
        'ms': round1_ms, 'bound': 'hbm', 'algorithmic_bytes': n_img * 4.0 * cfg.height * cfg.width}}
    if not task_rep:
        npts = float(res['n_cov'].float().sum().item())
        hw = (cfg.height // cfg.desc_stride) * (cfg.width // cfg.desc_stride)
        kernels['sample kernel'] = {'ms': stage_ms['sample'], 'bound': 'hbm', 'algorithmic_bytes':
                                    min(16 * npts * cfg.desc_dim, 4.0 * cfg.desc_dim * hw * n_img) + 4 * npts * cfg.desc_dim + 8 * npts}
        ncov = res['n_cov'].float()
        flops = float((2.0 * ncov[:P] * ncov[P:] * cfg.desc_dim).sum().item()) * (2 if cfg.cross_check else 1)
        if tc:      # the tcgen05 search kernel alone (kb_match_mnn_phases), on the buffers of a full call
            mst = []
            margs = (res['desc'][:P], res['desc'][P:], res['n_cov'][:P], res['n_cov'][P:], cfg.max_distance, cfg.cross_check)
            ops.match_batched(*margs, algo=1, state=mst, want_dist=False)
            top2_ms = time_stage(lambda: ops.match_batched(*margs, algo=1, phases=2, state=mst, want_dist=False))
            kernels['nn_top2_kernel (match, tcgen05 Gram + fused top-3 epilogue)'] = {
                'ms': top2_ms, 'bound': 'tensor', 'algorithmic_flops': flops}
        else:
            kernels['match stage (float64 SIMT)'] = {'ms': stage_ms['match'], 'bound': 'tensor', 'algorithmic_flops': flops}
    for k in kernels.values():
        if k['bound'] == 'hbm':
            k['achieved'] = k['algorithmic_bytes'] / (k['ms'] / 1e3) / 1e9
            k['unit'], k['peak'] = 'GB/s', pk['hbm_gbs']
        else:
            k['achieved'] = k['algorithmic_flops'] / (k['ms'] / 1e3) / 1e12
            k['unit'], k['peak'] = 'TFLOP/s', pk['bf16_tflops_sustained']
        k['frac'] = k['achieved'] / k['peak']
    dom = max(kernels, key=lambda n: kernels[n]['ms'])
    # ncu --set full figures captured for this workload (profiles/traffic.json): DRAM bytes per launch, tensor-pipe %
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        for name, k in kernels.items():
            t = tj.get(name.split(' ')[0].replace('sample', 'sample_planes_kernel') if name.startswith('sample') else name.split(' ')[0], {})
            if t.get('workload') == cfg.name and t.get('maps') == n_img:
                k['ncu'] = {kk: vv for kk, vv in t.items() if kk not in ('workload', 'maps', 'source')}
                if name == dom:
                    traffic = t.get('dram_bytes_per_launch')
    d = kernels[dom]
    roof = {'kernel': dom, 'bound': d['bound'], 'achieved': d['achieved'], 'peak': d['peak'], 'unit': d['unit'],
            'frac': d['frac'], 'traffic': traffic, 'stage_ms': stage_ms, 'kernels': kernels,
            'note': f'algorithmic bytes (or one-pass 2nmD flops per direction) of one launch / CUDA-event time of the kernel '
                    f'launched alone 5x back to back; peaks {pk["source"]} (HBM copy, bf16 sustained); traffic = dram bytes of '
                    f'one ncu --set full capture (profiles/)'}

    # ---- CPU baseline on a bounded sample + parity spot check ------------------------------------
    cpu = None
    if world == 1 and args.cpu_pairs > 0:
        bc = {'score': batch.score.cpu(), 'desc': batch.desc.cpu() if batch.desc is not None else None}
        dt, results = run_cpu(cfg, bc, hms, args.cpu_pairs)
        ok = True
        for i, (k0, k1, k0c, k1c, pairs) in enumerate(results):
            n0 = int(res['n_kpts'][i])
            ok &= bool(np.array_equal(np.sort(res['kpts'][i, :n0, 2].cpu().numpy()), np.sort(k0[:, 2])))
            if pairs is not None:
                got = res['matches'][i, :int(res['n_matches'][i])].cpu().numpy()
                ok &= bool(abs(got.shape[0] - pairs.shape[0]) <= 2)
        cpu = {'value': args.cpu_pairs / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': f'{args.cpu_pairs} pairs of {cfg.name} through oracle/ref_ops.py (the reference\'s im2col NMS '
                         f'rounds, grid_sample, scipy float64 cdist); scipy cdist is single-threaded',
               'gpu_matches_oracle_on_sample': ok}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32+f64', 'data': 'synthetic', 'config': config, 'clocks': clocks,
            'e2e': e2e, 'gpu_launches': n_launch, 'roofline': roof, 'cpu_baseline': cpu,
            'counts': {'sum_matches_or_rep': float(acc[0]), 'pairs': float(acc[1])}}
    if serial_ms is not None:       # the same K steps strictly one after another (what the per-kernel shares refer to)
        line['serial'] = {'ms_per_step': serial_ms / args.steps, 'value': world * P * args.steps / (serial_ms / 1000.0)}
    print(json.dumps(line))
    parallel.shutdown()
    return 0


if __name__ == '__main__':
    sys.exit(main())
