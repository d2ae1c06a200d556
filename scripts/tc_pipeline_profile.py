"""Where does nn_top2_kernel wait?  kb_debug_knob(KB_KNOB_TC_DEBUG, 4) makes the producer / MMA / epilogue lanes accumulate the cycles
they spend in mbarrier waits (clock64) and store them per CTA; this prints the median over CTAs as a share of the
role's total loop time.  Diagnostic only (the counters perturb the kernel by a few percent)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import _lib, ops
from keypoint_bench_b200._lib import lib, check
lib.kb_debug_knob(_lib.KB_KNOB_TC_DEBUG, 4 | int(os.environ.get('KB_TC_DEBUG_EXTRA', '0')))

shapes = [(64, 1000, 1000, 256), (16, 4096, 4096, 64)]
g = torch.Generator().manual_seed(1)
for B, n, m, D in shapes:
    a = torch.nn.functional.normalize(torch.randn(B, n, D, generator=g), dim=2).cuda()
    b = torch.nn.functional.normalize(torch.randn(B, m, D, generator=g), dim=2).cuda()
    b[:, :n // 2] = a[:, :n // 2] + 0.05 * torch.randn(B, n // 2, D, generator=g).cuda()
    for _ in range(3):
        _, _, _, ws = ops.match_batched(a, b, None, None, 5.0, True, algo=1, return_ws=True)
    torch.cuda.synchronize()
    off = (ctypes.c_size_t * 6)()
    check(lib.kb_match_tc_debug_offsets(B, n, m, D, ctypes.cast(off, ctypes.c_void_p)), 'offsets')
    prof = ws[off[5]:off[5] + 8 * 512 * 8].view(torch.int64).reshape(512, 8).cpu()
    prof = prof[:148]
    lead = prof[prof[:, 3] > 0]          # CTAs whose MMA warp ran (pair mode: leaders only)
    med = lambda t: float(t.float().median())
    print(f'B={B} n={n} D={D}  (cycles, median over CTAs)')
    print(f'  producer loop {med(prof[:, 0]):10.0f}  wait a_free {med(prof[:, 1]):9.0f}  wait b_empty {med(prof[:, 2]):9.0f}')
    print(f'  mma loop      {med(lead[:, 3]):10.0f}  wait t_empty {med(lead[:, 4]):8.0f}  wait a_full {med(lead[:, 5]):9.0f}  wait b_full {med(lead[:, 6]):9.0f}')
    print(f'  epilogue wait t_full {med(prof[:, 7]):10.0f}')
