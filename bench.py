#!/usr/bin/env python
"""Benchmark of the post-network hot path (BASELINE.json metric: image pairs/sec, extract+match).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

A "step" is one pass of the hot path over one batch of synthetic image pairs per GPU, as the task the config names
runs it in the reference (keypoint_bench_b200.pipeline.run_task):

    cfg1  repeatability   detect x2 -> warp x2 -> val_key_points counting            (tasks/repeatability.py:95-122)
    cfg2  MHA (headline)  detect x2 -> warp x2 -> sample + match the covisible keypoints (tasks/MHA.py:29-39,
                          up to, not including, the host-side cv2 RANSAC)
    cfg3 / cfg4  match    detect x2 -> sample + match ALL keypoints                  (tasks/AUC.py:115-120)
    cfg5  stream          F+1 frames: every frame extracted once, matched with its predecessor
                          (models/model_interface.py:217-228 + tasks/FundamentalMatrix.py:53-57)

N>1 = one process per GPU under torchrun; pairs are independent so every rank processes its own batch (weak scaling;
cfg5: contiguous frame chunks with a one-frame halo) and the only collective is the all-reduce of the count vector at
the end of the timed region.  The default run times the headline config (cfg2) in full -- device-timed value, e2e with
host buffers, roofline of the dominant kernel, CPU baseline -- and then every other config briefly (`configs` key),
because the driver's command line cannot pass --config.  Rank 0 prints ONE JSON line.
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
HEADLINE = 'cfg2'
STREAM_SEED = 5150


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default=HEADLINE, help='cfg1..cfg5 (BASELINE.json configs); cfg2 is the headline')
    ap.add_argument('--pairs', type=int, default=0, help='pairs per GPU per step (0 = config default)')
    ap.add_argument('--algo', type=int, default=-1, help='matcher: 0 = float64 SIMT, 1 = tcgen05; -1 = best available')
    ap.add_argument('--kind', default='uniform', help='synthetic score-map kind (uniform | alike)')
    ap.add_argument('--cpu-pairs', type=int, default=2, help='pairs timed on the host for cpu_baseline (0 = skip)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-others', action='store_true', help='skip the brief runs of the other four configs')
    ap.add_argument('--fused-table', action='store_true', help='per-kernel table of the opt-in fused sampling + operand preparation (A/B)')
    ap.add_argument('--in-flight', type=int, default=4,
                    help='steps kept in flight (one CUDA graph + stream + batch per slot); 1 = strictly serial steps')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from Python instead of replaying a CUDA graph')
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the benchmark runs (B200_PROFILING.md)."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.proc = None
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


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


def bind_to_gpu_cpus(local_rank):
    """Pin this process to the cores NVML names as local to its GPU BEFORE pinned host buffers are allocated, so the
    e2e copies read host memory of the GPU's own NUMA node.  Returns the number of cores bound (None if unavailable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's algorithm (oracle port) on the host, one pair as the config's task runs it
# ------------------------------------------------------------------------------------------------

def cpu_pair(cfg, score0, score1, desc0, desc1, hm):
    """One pair through the oracle exactly as the reference would run it on CPU: detection with the im2col / argmax /
    col2im rounds (utils/extracter.py:49-98), then the task's own continuation."""
    from keypoint_bench_b200 import synth
    from oracle import ref_ops
    H, W = cfg.height, cfg.width
    k0, _ = ref_ops.detection(score0, cfg.extractor_params, nms='im2col')
    k1, _ = ref_ops.detection(score1, cfg.extractor_params, nms='im2col')
    out = {'k0': k0, 'k1': k1}
    if cfg.task == 'repeatability':
        w01, w10 = synth.warp_params(hm, H, W)
        out['rep'] = ref_ops.val_key_points(k0, k1, w01, w10, th=3)
        return out
    m0, m1 = k0, k1
    if cfg.task == 'mha':                                       # tasks/MHA.py:33-34
        w01, w10 = synth.warp_params(hm, H, W)
        m0, _, _, _ = ref_ops.warp(k0, w01)
        m1, _, _, _ = ref_ops.warp(k1, w10)
    out['m0'], out['m1'] = m0, m1
    if m0.shape[0] and m1.shape[0]:
        out['d0'] = ref_ops.sample_brute_force(desc0, m0)
        out['d1'] = ref_ops.sample_brute_force(desc1, m1)
        out['pairs'] = ref_ops.match_descriptors(out['d0'], out['d1'], metric='euclidean', max_distance=cfg.max_distance,
                                                 cross_check=cfg.cross_check)
    return out


def cpu_inputs(cfg, cfg_index, n_pairs, batch=None, hms=None):
    """Host tensors of `n_pairs` pairs: (score0, score1, desc0, desc1, H) per pair.  `batch` (+ `hms`) = the device
    batch the GPU arm processed: its first pairs are copied to the host so both sides see identical inputs; without
    it the pairs are generated on the host (the reference arm)."""
    from keypoint_bench_b200 import synth
    if cfg.task == 'stream':
        fr = batch if batch is not None else synth.make_frames(cfg, STREAM_SEED, 0, n_pairs + 1, n_pairs + 1, 'cpu')
        sc, de = fr.score[:n_pairs + 1].cpu(), fr.desc[:n_pairs + 1].cpu()
        return [(sc[i:i + 1], sc[i + 1:i + 2], de[i:i + 1].numpy(), de[i + 1:i + 2].numpy(), None) for i in range(n_pairs)]
    if batch is None:
        batch, hms = synth.make_batch(cfg, cfg_index, n_pairs, 0, 'cpu')
    P = batch.pairs
    idx = list(range(n_pairs)) + list(range(P, P + n_pairs))
    sc = batch.score[idx].cpu()
    de = None if batch.desc is None else batch.desc[idx].cpu()
    n = n_pairs
    return [(sc[i:i + 1], sc[n + i:n + i + 1], None if de is None else de[i:i + 1].numpy(),
             None if de is None else de[n + i:n + i + 1].numpy(), hms[i].cpu()) for i in range(n)]


def run_cpu(cfg, inputs):
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    results = [cpu_pair(cfg, *inp) for inp in inputs]
    return time.perf_counter() - t0, results


def base_config(cfg, P, world, kind):
    return {'workload': cfg.name, 'task': cfg.task, 'pairs_per_gpu_per_step': P, 'height': cfg.height, 'width': cfg.width,
            'desc_dim': cfg.desc_dim, 'desc_stride': cfg.desc_stride, 'nms_dist': cfg.nms_dist, 'top_k': cfg.top_k,
            'border_dist': cfg.border_dist, 'max_distance': cfg.max_distance, 'cross_check': cfg.cross_check,
            'score_map': kind, 'sharding': (f'contiguous frame chunks + 1-frame halo x{world}' if cfg.task == 'stream'
                                            else f'by-pair x{world}'),
            'timed_stages': {'repeatability': 'detect x2, warp x2, val_key_points counting',
                             'mha': 'detect x2, warp x2, sample x2, mutual-NN match of the covisible keypoints (host cv2 RANSAC excluded)',
                             'match': 'detect x2, sample x2, mutual-NN match of all keypoints',
                             'stream': 'detect + sample once per frame, mutual-NN match of consecutive frames'}[cfg.task],
            'l2': 'working set per step > 126 MB L2 (no flush needed)'}


# ------------------------------------------------------------------------------------------------
# one config on this rank's GPU
# ------------------------------------------------------------------------------------------------

def time_ms(fn, reps=5):
    """`reps` back-to-back calls between two CUDA events on the launch (current) stream, after one warm call."""
    from keypoint_bench_b200 import ops
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


FUSED_TABLE = False


def kernel_table(cfg, inputs, res, algo, tc, pk):
    """Every kernel group of one step timed ALONE (5 back-to-back launches between CUDA events on the launch stream,
    through the library's measurement hooks), with its algorithmic bytes / flops and roofline fraction.
    Peaks: measured HBM copy bandwidth; bf16 BURST for the isolated tensor-core kernel."""
    from keypoint_bench_b200 import ops
    P = inputs.pairs
    score, desc = inputs.score, inputs.desc
    n_img = score.shape[0]
    stream = cfg.task == 'stream'
    k = {}
    st = []
    ops.detect_batched(score, cfg.extractor_params, state=st)
    hbm_img = n_img * 4.0 * cfg.height * cfg.width
    n_out = float(res['n_kpts'].float().sum().item())
    k['detect.tau (threshold estimate)'] = {'ms': time_ms(lambda: ops.detect_batched(score, cfg.extractor_params, phases=1, state=st)),
                                            'bound': 'latency'}
    k['detect.round1 (NMS round 1, the one pass over the score maps)'] = {
        'ms': time_ms(lambda: ops.detect_batched(score, cfg.extractor_params, phases=2, state=st)), 'bound': 'hbm',
        'algorithmic_bytes': hbm_img}
    ops.detect_batched(score, cfg.extractor_params, phases=7, state=st)      # (the lists the repeated round-1 launches grew) restored
    k['detect.resolve (per-map greedy resolve + top-k sort + fallback stubs)'] = {
        'ms': time_ms(lambda: ops.detect_batched(score, cfg.extractor_params, phases=4, state=st)), 'bound': 'latency',
        'algorithmic_bytes': 16.0 * n_out}
    k['detect (whole stage)'] = {'ms': time_ms(lambda: ops.detect_batched(score, cfg.extractor_params)), 'bound': 'hbm',
                                 'algorithmic_bytes': hbm_img + 16.0 * n_out, 'stage': True}
    pts, n_pts = res['kpts'], res['n_kpts']
    if cfg.task in ('repeatability', 'mha'):
        k['warp_homography'] = {'ms': time_ms(lambda: ops.warp_batched(res['kpts'], res['n_kpts'], inputs.h33, inputs.wh)),
                                'bound': 'latency', 'algorithmic_bytes': 8.0 * n_out + 36.0 * n_img}
        pts, n_pts = res['kcov'], res['n_cov']
    if cfg.task == 'repeatability':
        kv, kw, nv = res['kcov'], res['kwarp'], res['n_cov']
        k['repeat_counts (val_key_points core)'] = {
            'ms': time_ms(lambda: ops.repeat_batched(kv[:P], kw[:P], nv[:P], kv[P:], kw[P:], nv[P:], 512.0, 512.0, 3.0,
                                                     want_errors=True)),
            'bound': 'latency', 'algorithmic_bytes': 16.0 * float(nv.float().sum().item())}
        return k
    npts = float(n_pts.float().sum().item())
    hw = (cfg.height // cfg.desc_stride) * (cfg.width // cfg.desc_stride)
    sample_bytes = min(16 * npts * cfg.desc_dim, 4.0 * cfg.desc_dim * hw * n_img) + 4 * npts * cfg.desc_dim + 8 * npts
    dd = res['desc']
    a, b = (dd[:-1], dd[1:]) if stream else (dd[:P], dd[P:])
    na, nb = (n_pts[:-1], n_pts[1:]) if stream else (n_pts[:P], n_pts[P:])
    flops = float((2.0 * na.double() * nb.double() * cfg.desc_dim).sum().item())        # ONE pass of 2 n m D per pair
    margs = (a, b, na, nb, cfg.max_distance, cfg.cross_check)
    search_name = 'match.search (tcgen05 Gram + fused top-3 epilogue)'
    tail_name = 'match.tail (certify / rescan / gate / pairs)'
    passes = ops.match_issue_factor(bool(cfg.cross_check), cfg.desc_dim) if tc else 0
    note = f'{passes}x the one-pass flops are issued to the tensor pipe (directions x products of the split operands)'
    fused = (FUSED_TABLE and tc and not stream and pts.shape[1] > 0 and
             bool(ops.lib.kb_sample_desc_operands_supported(cfg.desc_dim, cfg.height // cfg.desc_stride, cfg.width // cfg.desc_stride,
                                                            pts.shape[1])))
    if fused:
        # --fused-table: the opt-in form in which the sampler writes the matcher's operand rows (kb_sample_desc_operands);
        # the step itself runs the two calls (measured faster), so this table is for A/B only
        fst = []
        fargs = (desc, pts, n_pts, P, cfg.max_distance, cfg.cross_check)
        ops.sample_match_batched(*fargs, algo=1, want_dist=False, fused=True, state=fst)
        k['sample (bilinear descriptor sampling + 16-bit operand rows of the matcher, fused)'] = {
            'ms': time_ms(lambda: ops.sample_match_batched(*fargs, algo=1, want_dist=False, state=fst, part='sample')), 'bound': 'hbm',
            'algorithmic_bytes': sample_bytes + 4 * npts * cfg.desc_dim,
            'bytes_note': 'maps once + float32 rows + hi/lo 16-bit operand rows + coordinates'}
        k['match.finish (row norms from the sampler\'s partial sums, c, padding)'] = {
            'ms': time_ms(lambda: ops.sample_match_batched(*fargs, algo=1, want_dist=False, state=fst, part='finish')), 'bound': 'latency',
            'algorithmic_bytes': 0.0}
        k[search_name] = {
            'ms': time_ms(lambda: ops.sample_match_batched(*fargs, algo=1, want_dist=False, state=fst, part='search')), 'bound': 'tensor',
            'algorithmic_flops': flops, 'issued_flops': flops * passes, 'issued_note': note}
        k[tail_name] = {
            'ms': time_ms(lambda: ops.sample_match_batched(*fargs, algo=1, want_dist=False, state=fst, part='tail')), 'bound': 'latency',
            'algorithmic_bytes': 12.0 * float(res['n_matches'].float().sum().item())}
        k['sample + match (whole stage)'] = {
            'ms': time_ms(lambda: ops.sample_match_batched(*fargs, algo=1, want_dist=False, state=fst)), 'bound': 'tensor',
            'algorithmic_flops': flops, 'stage': True}
        return k
    k['sample (bilinear descriptor sampling)'] = {
        'ms': time_ms(lambda: ops.sample_batched(desc, pts, n_pts)), 'bound': 'hbm',
        'algorithmic_bytes': sample_bytes,
        'sector_bytes': (64.0 * npts * cfg.desc_dim if cfg.desc_stride == 1 else None)}
    if tc:
        mst = []
        ops.match_batched(*margs, algo=1, state=mst, want_dist=False)
        k['match.prep (hi/lo 16-bit split + norms)'] = {
            'ms': time_ms(lambda: ops.match_batched(*margs, algo=1, phases=1, state=mst, want_dist=False)), 'bound': 'hbm',
            'algorithmic_bytes': 0.0}
        k[search_name] = {
            'ms': time_ms(lambda: ops.match_batched(*margs, algo=1, phases=2, state=mst, want_dist=False)), 'bound': 'tensor',
            'algorithmic_flops': flops, 'issued_flops': flops * passes, 'issued_note': note}
        k[tail_name] = {
            'ms': time_ms(lambda: ops.match_batched(*margs, algo=1, phases=4, state=mst, want_dist=False)), 'bound': 'latency',
            'algorithmic_bytes': 12.0 * float(res['n_matches'].float().sum().item())}
    k['match (whole stage)'] = {'ms': time_ms(lambda: ops.match_batched(*margs, algo=algo, want_dist=False)),
                                'bound': 'tensor', 'algorithmic_flops': flops, 'stage': True}
    return k


def finish_table(kernels, pk):
    for v in kernels.values():
        if v['bound'] == 'tensor':
            v['achieved'] = v['algorithmic_flops'] / (v['ms'] / 1e3) / 1e12
            v['unit'], v['peak'] = 'TFLOP/s', pk['bf16_tflops']           # burst: the kernel is timed alone
        else:
            v['achieved'] = v.get('algorithmic_bytes', 0.0) / (v['ms'] / 1e3) / 1e9
            v['unit'], v['peak'] = 'GB/s', pk['hbm_gbs']
        v['frac'] = v['achieved'] / v['peak']
    single = {n: v for n, v in kernels.items() if not v.get('stage')}
    return max(single, key=lambda n: single[n]['ms'])


def run_config(args, name, rank, local_rank, world, device, full):
    """Times one config on this rank.  full: headline treatment (stage table, e2e, CPU baseline)."""
    from keypoint_bench_b200 import ops, parallel, pipeline, synth
    cfg = synth.CONFIGS[name]
    cfg_index = int(name[3:])
    P = (args.pairs if full and args.pairs else cfg.pairs_per_gpu)
    steps = args.steps if full else max(3, min(args.steps, 10))
    config = base_config(cfg, P, world, args.kind)
    algo = args.algo
    tc = cfg.desc_dim > 0 and (algo == 1 or (algo < 0 and cfg.desc_dim <= 256))
    if cfg.desc_dim:
        config['matcher'] = ('tcgen05 Gram of split 16-bit operands (fp16 x 2 products for D > 64, bf16 x 3 for D <= 64) + float64 certify'
                             if tc else 'float64 SIMT')
    stream = cfg.task == 'stream'
    depth = 1 if args.no_graph else max(1, args.in_flight)
    total_frames = world * P + 1                                  # one sequence per step slot, sharded over the ranks

    hms_of = {}

    def make_inputs(slot):
        if stream:
            lo, hi = parallel.shard_stream(total_frames, rank, world)
            return synth.make_frames(cfg, STREAM_SEED + slot, lo, hi - lo, total_frames, device)
        b, hms_of[slot] = synth.make_batch(cfg, cfg_index, P, (rank + world * slot) * P, device, args.kind)
        return b

    inputs = make_inputs(0)

    def step(b=None, timer=None):
        return pipeline.run_task(inputs if b is None else b, cfg, algo=algo, timer=timer)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = ops.launches()
    res0, _ = step()
    launches_per_step = ops.launches() - l0
    torch.cuda.synchronize()

    batches = [inputs] + [make_inputs(s) for s in range(1, depth)]
    graphed = flight = None
    if not args.no_graph:
        try:
            flight = pipeline.StepsInFlight([(lambda b=b: step(b)) for b in batches])
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
        serial_ms, _ = timed_steps(steps, False)      # the same steps strictly one after another, for reference
    ms, acc = timed_steps(steps, True)
    value = world * P * steps / (ms / 1000.0)
    res, _ = run()                                    # slot 0: the outputs checked against the CPU port below
    torch.cuda.synchronize()
    out = {'config': config, 'value': value, 'ms_per_step': ms / steps, 'steps': steps, 'pairs_per_step': world * P,
           'gpu_launches': launches_per_step * steps,
           'counts': {'sum_matches_or_rep': float(acc[0]), 'pairs': float(acc[1])}}
    if serial_ms is not None:
        out['serial'] = {'ms_per_step': serial_ms / steps, 'value': world * P * steps / (serial_ms / 1000.0)}

    # ---- cfg5 on N ranks: the reduced match count equals a single-rank run over the same frames ----------------
    if stream:
        _, acc1 = timed_steps(1, False)               # one step of slot 0 on every rank, reduced: [sum matches, pairs]
        if rank == 0:
            # the same sequence on rank 0 alone, cut into chunks that do NOT coincide with the rank shards
            tot, npairs, n_chunks = 0.0, 0, 2 * world + 1
            for c in range(n_chunks):
                lo, hi = parallel.shard_stream(total_frames, c, n_chunks)
                if hi - lo < 2:
                    continue
                fr = synth.make_frames(cfg, STREAM_SEED, lo, hi - lo, total_frames, device)
                r1, a1 = pipeline.run_task(fr, cfg, algo=algo)
                tot += float(a1[0]); npairs += int(a1[1])
                del fr, r1
            out['stream_check'] = {'frames': total_frames, 'ranks': world, 'reduced_matches': float(acc1[0]),
                                   'reduced_pairs': float(acc1[1]), 'single_rank_matches': tot, 'single_rank_pairs': npairs,
                                   'equal': bool(tot == float(acc1[0]) and npairs == int(acc1[1]))}

    # ---- end to end through the public API with HOST buffers (headline only) -----------------------------------
    if full and not args.no_e2e and not stream:
        out['e2e'] = run_e2e(args, cfg, batches, step, flight, depth, P, world, device, cfg.task == 'repeatability')

    if rank != 0:
        return out

    # ---- per-kernel table + roofline of the dominant kernel ----------------------------------------------------
    pk = peaks()
    kernels = kernel_table(cfg, inputs, res, algo, tc, pk)
    dom = finish_table(kernels, pk)
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        for kn, kv in kernels.items():
            t = tj.get(kn.split(' ')[0], {})
            if t.get('workload') == cfg.name and t.get('maps') == inputs.score.shape[0]:
                kv['ncu'] = {kk: vv for kk, vv in t.items() if kk not in ('workload', 'maps', 'source')}
                if kn == dom:
                    traffic = t.get('dram_bytes_per_launch')
    d = kernels[dom]
    out['roofline'] = {
        'kernel': dom, 'bound': 'tensor' if d['bound'] == 'tensor' else 'hbm', 'achieved': d['achieved'], 'peak': d['peak'],
        'unit': d['unit'], 'frac': d['frac'], 'traffic': traffic, 'kernels': kernels,
        **({'bound_note': 'the largest kernel group of this step is LATENCY-bound (a chain of small launches over KB-sized '
                          'inputs): its fraction of the HBM peak says nothing about it'} if d['bound'] == 'latency' else {}),
        'note': f'dominant = the kernel group with the largest CUDA-event time when launched alone 5x back to back (every '
                f'group of the step is timed); achieved = algorithmic bytes, or ONE pass of 2nmD flops per pair, per launch / '
                f'that time; peaks {pk["source"]}: HBM copy {pk["hbm_gbs"]:.0f} GB/s, bf16 burst {pk["bf16_tflops"]:.0f} TFLOP/s '
                f'(isolated kernel); traffic = dram bytes of one ncu --set full capture (profiles/)'}

    # ---- CPU baseline on a bounded sample + exact parity check of that sample ----------------------------------
    if full and world == 1 and args.cpu_pairs > 0:
        out['cpu_baseline'] = cpu_baseline(cfg, cfg_index, args.cpu_pairs, res, P, inputs, hms_of.get(0))
    return out


def cpu_baseline(cfg, cfg_index, n_pairs, res, P, batch, hms):
    """The oracle port timed on `n_pairs` pairs of the batch the GPU just processed (same seeds), and an EXACT check of
    the GPU outputs on them: keypoint rows of both images and the match pair sets (modulo north_star's 1e-5 near-tie
    rule, evaluated on float64 distances)."""
    from oracle.compare import check_detection_rows, exact_pairs_or_near_tie
    stream = cfg.task == 'stream'
    inputs = cpu_inputs(cfg, cfg_index, n_pairs, batch, hms)
    dt, results = run_cpu(cfg, inputs)
    ok, why = True, None
    try:
        for i, r in enumerate(results):
            i0, i1 = (i, i + 1) if stream else (i, P + i)
            for idx, want in ((i0, r['k0']), (i1, r['k1'])):
                n = int(res['n_kpts'][idx])
                check_detection_rows(res['kpts'][idx, :n].cpu().numpy(), want)
            if cfg.task == 'repeatability':
                st = res['stats'][i].cpu().numpy()
                assert int(st[0]) == r['rep']['gt_num'] and int(res['num_feat'][i]) == r['rep']['num_feat'], (st, r['rep']['gt_num'])
            elif 'pairs' in r:
                got = res['matches'][i, :int(res['n_matches'][i])].cpu().numpy().astype(np.int64)
                exact_pairs_or_near_tie(got, r['d0'], r['d1'], cfg.max_distance, cfg.cross_check)
    except AssertionError as e:
        ok, why = False, repr(e)[:300]
    cpu = {'value': n_pairs / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
           'sample': f'{n_pairs} pairs of {cfg.name} through oracle/ref_ops.py (the reference\'s im2col NMS rounds, '
                     f'grid_sample, scipy float64 cdist); scipy cdist is single-threaded',
           'gpu_matches_oracle_on_sample': ok,
           'check': 'exact keypoint rows of both images and exact match pairs (1e-5 near-tie rule) / exact gt_num, num_feat'}
    if why:
        cpu['mismatch'] = why
    return cpu


def run_e2e(args, cfg, batches, step, flight, depth, P, world, device, task_rep):
    """Every slot: pinned host inputs -> device copy -> the step's graph -> device -> pinned host results, all on the
    slot's stream; the host waits for a slot's previous results before it relaunches that slot (the caller consumes the
    result of every step), so with several slots one step's copies overlap another's kernels."""
    from keypoint_bench_b200 import parallel, pipeline
    top = cfg.top_k
    batch = batches[0]
    in_bytes = batch.score.numel() * 4 + (batch.desc.numel() * 4 if batch.desc is not None else 0)
    share_host = in_bytes > (256 << 20)        # one pinned copy of large inputs serves every slot (the bytes copied are the same)
    slots = []
    for k, bk in enumerate(batches):
        share = k > 0 and share_host
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
            e2e_flight = pipeline.StepsInFlight([(lambda b=sl['batch']: step(b)) for sl in slots])
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
                copy_out(0, step(slots[0]['batch']))
                torch.cuda.synchronize()

    e2e_steps = max(3, min(args.steps, 10))
    e2e_steps_run(depth)
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps_run(e2e_steps)
    dt_rank = time.perf_counter() - t0
    dt = dt_rank
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        dt = float(t.item())
    d2h = slots[0]['out_stats'].numel() * 8 if task_rep else (slots[0]['out_pairs'].numel() * 4 + slots[0]['out_n'].numel() * 4)
    return {'value': world * P * e2e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': in_bytes, 'd2h_bytes_per_step': d2h,
            'steps': e2e_steps, 'steps_in_flight': depth if e2e_flight is not None else 1,
            'h2d_gbs_this_rank': in_bytes * e2e_steps / dt_rank / 1e9,
            'pinned_host_copies': 1 if share_host else depth,
            'note': 'pinned host -> device copy of score+descriptor maps, hot path, device -> host copy of match '
                    'index pairs and counts, every step; the host waits for a slot\'s results before reusing the slot; '
                    'this number is the H2D link (PCIe), not the kernels'}


def main():
    global FUSED_TABLE
    args = parse_args()
    FUSED_TABLE = bool(args.fused_table)
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    # -------------------------------------------------------------------------------- reference arm
    if args.impl == 'reference':
        # the reference's own CPU algorithm through the oracle port; nothing of the product (no CUDA library) is loaded
        if rank != 0:
            return 0
        from keypoint_bench_b200 import synth
        cfg = synth.CONFIGS[args.config]
        inputs = cpu_inputs(cfg, int(args.config[3:]), 1)
        for _ in range(min(args.warmup, 1)):                     # one warm pair is enough to page in torch
            run_cpu(cfg, inputs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run_cpu(cfg, inputs)
        dt = time.perf_counter() - t0
        value = args.steps / dt
        line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000 * dt / args.steps,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32+f64',
                'data': 'synthetic', 'config': dict(base_config(cfg, 1, 1, args.kind), pairs_per_gpu_per_step=1),
                'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                                 'sample': f'{args.steps} steps x 1 pair of {cfg.name} through oracle/ref_ops.py '
                                           f'(im2col NMS rounds, grid_sample, scipy cdist float64)'},
                'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'product_library_loaded': 'keypoint_bench_b200._lib' in sys.modules}
        print(json.dumps(line))
        return 0

    # -------------------------------------------------------------------------------- our arm
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback for the product path)')
    bound = bind_to_gpu_cpus(local_rank)
    from keypoint_bench_b200 import parallel
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        parallel.init('nccl')
    sampler = ClockSampler(local_rank)
    head = run_config(args, args.config, rank, local_rank, world, device, full=True)
    clocks = sampler.stop()
    others = {}
    if not args.no_others and args.config == HEADLINE:
        for name in ('cfg1', 'cfg3', 'cfg4', 'cfg5'):
            torch.cuda.empty_cache()
            r = run_config(args, name, rank, local_rank, world, device, full=False)
            roof = r.get('roofline')
            others[name] = {'workload': r['config']['workload'], 'task': r['config']['task'], 'value': r['value'], 'unit': UNIT,
                            'ms_per_step': r['ms_per_step'], 'steps': r['steps'], 'pairs_per_step': r['pairs_per_step'],
                            'serial': r.get('serial'), 'counts': r['counts'], 'stream_check': r.get('stream_check'),
                            'roofline': None if roof is None else {
                                'kernel': roof['kernel'], 'bound': roof['bound'], 'achieved': roof['achieved'], 'peak': roof['peak'],
                                'unit': roof['unit'], 'frac': roof['frac'],
                                'kernels': {k: {kk: v[kk] for kk in ('ms', 'bound', 'achieved', 'unit', 'frac') if kk in v}
                                            for k, v in roof['kernels'].items()}}}
    if rank != 0:
        parallel.shutdown()
        return 0
    cfg_line = dict(head['config'])
    cfg_line['cpu_affinity_cores'] = bound
    line = {'metric': METRIC, 'value': head['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32+f64', 'data': 'synthetic', 'config': cfg_line, 'clocks': clocks,
            'e2e': head.get('e2e'), 'gpu_launches': head['gpu_launches'], 'roofline': head.get('roofline'),
            'cpu_baseline': head.get('cpu_baseline'), 'counts': head['counts']}
    for k in ('serial', 'stream_check'):
        if head.get(k) is not None:
            line[k] = head[k]
    if others:
        line['configs'] = others
    print(json.dumps(line))
    parallel.shutdown()
    return 0


if __name__ == '__main__':
    sys.exit(main())
