"""GPU probe of the tcgen05 matcher (algo 1) against the float64 SIMT matcher (algo 0):
agreement of pairs, accuracy of the split-bf16 scores against float64, rescan counts, timing."""
import math
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops


def case(B, n, m, D, maxd=5.0, cc=True, seed=0, norm=True, verbose=True):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, n, D, generator=g)
    b = torch.randn(B, m, D, generator=g)
    if norm:
        a = torch.nn.functional.normalize(a, dim=2)
        b = torch.nn.functional.normalize(b, dim=2)
    k = min(n, m) // 2
    b[:, :k] = a[:, :k] + 0.05 * torch.randn(B, k, D, generator=g)
    a, b = a.cuda(), b.cuda()
    p0, d0, c0 = ops.match_batched(a, b, None, None, maxd, cc, algo=0)
    torch.cuda.synchronize()
    p1, d1, c1, ws = ops.match_batched(a, b, None, None, maxd, cc, algo=1, return_ws=True)
    torch.cuda.synchronize()
    ok = bool(torch.equal(c0, c1))
    for i in range(B):
        k0 = int(c0[i])
        ok &= bool(torch.equal(p0[i, :k0], p1[i, :k0])) and bool(torch.allclose(d0[i, :k0], d1[i, :k0], rtol=1e-13, atol=1e-13))
    dbg = ops.match_tc_debug(ws, B, n, m, D)
    best, second, idx = dbg['res0']
    # exact t = x.y - |y|^2/2 of the reported argbest, float64
    a64, b64 = a.double(), b.double()
    idx_l = idx.long().reshape(B, n).clamp(0, m - 1)
    yb = torch.gather(b64, 1, idx_l[..., None].expand(B, n, D))
    t_exact = (a64 * yb).sum(-1) - 0.5 * (yb * yb).sum(-1)
    err = (best.reshape(B, n).double() - t_exact).abs()
    na = a64.norm(dim=2)
    nbmax = b64.norm(dim=2).max(dim=1).values[:, None]
    bound = 6.2e-5 * na * nbmax + 3.1e-5 * nbmax * nbmax
    ratio = float((err / bound).max())
    if verbose:
        print(f'B={B} n={n} m={m} D={D} maxd={maxd} cc={cc}: equal={ok} matches={int(c1.sum())} '
              f'n_exact={int(dbg["n_exact"][0])} n_pair={int(dbg["n_pair"][0])} max_err={float(err.max()):.3e} err/bound={ratio:.3f}', flush=True)
    return ok and ratio < 1.0


def kernel_table(fn, iters=5):
    """Per-kernel device time (CUPTI via torch.profiler; not a bench number, just shares)."""
    from torch.profiler import profile, ProfilerActivity
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / iters, e.count / iters) for e in prof.key_averages() if e.device_time_total > 0]
    for k, t, c in sorted(rows, key=lambda r: -r[1]):
        print(f'    {k[:70]:70s} {t:10.1f} us/iter  x{c:.0f}', flush=True)


def timing(B, n, m, D):
    g = torch.Generator().manual_seed(1)
    a = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=2).cuda()
    b = torch.nn.functional.normalize(torch.randn(B, m, D, generator=g), dim=2).cuda()
    b[:, :n // 2] = a[:, :n // 2] + 0.05 * torch.randn(B, n // 2, D, generator=g).cuda()
    kernel_table(lambda: ops.match_batched(a, b, None, None, 5.0, True, algo=1))
    for algo in (0, 1):
        for _ in range(3):
            ops.match_batched(a, b, None, None, 5.0, True, algo=algo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.match_batched(a, b, None, None, 5.0, True, algo=algo)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'timing B={B} n={n} m={m} D={D} algo={algo}: {ms:.3f} ms  {2.0 * B * n * m * D / ms / 1e9:.1f} TFLOP/s (one-pass count)',
              flush=True)


if __name__ == '__main__':
    allok = True
    allok &= case(1, 128, 128, 64)
    allok &= case(1, 300, 280, 64)
    allok &= case(2, 1000, 977, 256)
    allok &= case(3, 257, 511, 48, maxd=0.9)
    allok &= case(1, 2048, 2048, 128, cc=False, maxd=math.inf)
    allok &= case(1, 1, 5, 32)
    allok &= case(1, 130, 1, 64)
    allok &= case(2, 1000, 1000, 256, norm=False, seed=5)
    allok &= case(4, 4096, 4096, 64, seed=7)
    print('ALL OK' if allok else 'MISMATCH', flush=True)
    timing(64, 1000, 1000, 256)
    timing(64, 4096, 4096, 64)
    timing(8, 2048, 2048, 128)
