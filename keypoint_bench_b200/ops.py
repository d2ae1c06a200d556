"""Batched device-resident ops over the C ABI (include/kb_b200.h).

These are the throughput entry points: every function takes and returns CUDA tensors, launches
asynchronously on the current stream and never synchronises the host.  The per-image drop-ins
with the reference's own signatures live in ``keypoint_bench_b200.utils`` / ``.tasks`` and are
thin wrappers over these.

PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import check, lib

_launches = 0          # kernels launched through this module (bench.py reports it)
_zero_fill = True      # outputs are zero-initialised (rows beyond `count` read as zeros)


class no_zero_fill:
    """Context manager for steady-state callers (pipeline.py): outputs are allocated uninitialised, rows
    beyond the per-map counts are then unspecified.  Saves a dozen fill kernels per step."""

    def __enter__(self):
        global _zero_fill
        self.prev, _zero_fill = _zero_fill, False

    def __exit__(self, *exc):
        global _zero_fill
        _zero_fill = self.prev


def _out(*shape, dtype, device):
    return torch.zeros(*shape, dtype=dtype, device=device) if _zero_fill else torch.empty(*shape, dtype=dtype, device=device)


class debug_knob:
    """Context manager around ``kb_debug_knob`` (include/kb_b200.h): experiment switches for tests and A/B timings."""

    def __init__(self, knob: int, value: int):
        self.knob, self.value = knob, value

    def __enter__(self):
        self.prev = lib.kb_debug_knob(self.knob, self.value)
        if self.prev == _lib.KB_ERR_BAD_ARG:
            raise _lib.KbError(f'unknown knob {self.knob}')
        return self

    def __exit__(self, *exc):
        lib.kb_debug_knob(self.knob, self.prev)


def launches() -> int:
    return _launches


def _count(n: int) -> None:
    global _launches
    _launches += n


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.KbError(f'{name} must be a CUDA tensor (keypoint_bench_b200 has no CPU path)')


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous() if t.dtype == torch.float32 else t.to(torch.float32).contiguous()


def _i32(t):
    if t is None:
        return None
    return t.contiguous() if t.dtype == torch.int32 else t.to(torch.int32).contiguous()


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _maps3(score: torch.Tensor) -> torch.Tensor:
    """[B,1,H,W] / [B,H,W] / [H,W] -> contiguous float32 [B,H,W] (channels fold into the batch)."""
    s = _f32(score)
    if s.dim() == 2:
        s = s[None]
    elif s.dim() == 4:
        s = s.reshape(-1, s.shape[-2], s.shape[-1])
    elif s.dim() != 3:
        raise _lib.KbError(f'score map must have 2-4 dims, got {tuple(score.shape)}')
    return s


# ------------------------------------------------------------------------------------------------
# Stage 1
# ------------------------------------------------------------------------------------------------

def fast_nms_batched(score: torch.Tensor, nms_dist: int = 4, max_iter: int = -1, min_value: float = 0.0,
                     return_rounds: bool = False):
    """Round-faithful ``fast_nms`` (utils/extracter.py:6-100) on every map of the batch jointly."""
    _require_cuda(score, 'score')
    s = _maps3(score)
    b, h, w = s.shape
    out = torch.empty_like(s)
    rounds = torch.zeros(1, dtype=torch.int32, device=s.device)
    ws = _ws(lib.kb_fast_nms_workspace_bytes(b, h, w), s.device)
    with torch.cuda.device(s.device):
        check(lib.kb_fast_nms(s.data_ptr(), out.data_ptr(), b, h, w, int(nms_dist), int(max_iter), float(min_value),
                              rounds.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), 'kb_fast_nms')
    _count(1)
    out = out.reshape(score.shape) if score.dim() != 2 else out[0]
    return (out, rounds) if return_rounds else out


def simple_nms_batched(score: torch.Tensor, nms_radius: int) -> torch.Tensor:
    """``simple_nms`` of the LightGlue-style extractor (models/lightglue.py:904-920) on every map of the batch."""
    _require_cuda(score, 'score')
    if nms_radius < 0:
        raise AssertionError('nms_radius must be >= 0')       # lightglue.py:906
    s = _maps3(score)
    b, h, w = s.shape
    out = torch.empty_like(s)
    ws = _ws(lib.kb_simple_nms_workspace_bytes(b, h, w), s.device)
    with torch.cuda.device(s.device):
        check(lib.kb_simple_nms(s.data_ptr(), out.data_ptr(), b, h, w, int(nms_radius), ws.data_ptr(), ws.numel(),
                                _stream()), 'kb_simple_nms')
    _count(3)
    return out.reshape(score.shape) if score.dim() != 2 else out[0]


def select_batched(nms_map: torch.Tensor, border_dist: int, threshold: float, min_score: float, top_k: int,
                   cap: int | None = None):
    """Border + threshold + (top-k) + min_score on already-suppressed maps
    (utils/extracter.py:164-190, 129-161, 217-220).  -> xyp[B,cap,3], count[B], raster[B,cap], total[B]."""
    _require_cuda(nms_map, 'nms_map')
    s = _maps3(nms_map)
    b, h, w = s.shape
    if cap is None:
        cap = top_k if top_k > 0 else h * w
    xyp = _out(b, cap, 3, dtype=torch.float32, device=s.device)
    raster = _out(b, cap, dtype=torch.int32, device=s.device)
    count = torch.zeros(b, dtype=torch.int32, device=s.device)
    total = torch.zeros(b, dtype=torch.int32, device=s.device)
    ws = _ws(lib.kb_select_workspace_bytes(b, h, w, int(top_k)), s.device)
    with torch.cuda.device(s.device):
        check(lib.kb_select(s.data_ptr(), b, h, w, int(border_dist), float(threshold), float(min_score), int(top_k),
                            int(cap), xyp.data_ptr(), raster.data_ptr(), count.data_ptr(), total.data_ptr(),
                            ws.data_ptr(), ws.numel(), _stream()), 'kb_select')
    _count(1)
    return xyp, count, raster, total


SORT_CAP = 8192        # largest top_k the selection kernels sort on chip (kb_select.cu)
DETECT_MAX_B = 2048    # maps per kb_detect call (larger batches are split here)


def nms_keep_bound(h: int, w: int, nms_dist: int) -> int:
    """Kept pixels are pairwise more than ``nms_dist`` apart (Chebyshev): at most ceil(h/(r+1)) * ceil(w/(r+1))."""
    if nms_dist <= 0:
        return h * w
    return ((h + nms_dist) // (nms_dist + 1)) * ((w + nms_dist) // (nms_dist + 1))


def detect_batched(score: torch.Tensor, params: dict | None = None, phases: int = 7, state=None):
    """``detection`` (utils/extracter.py:193-221) for every map of the batch independently.
    -> xyp[B,top_k,3] (x,y,p), count[B], raster[B,top_k], path[B].
    ``phases`` / ``state`` are a measurement hook (bench.py): ``state=[]`` receives the buffers of a full call,
    a later call with the same ``state`` and ``phases`` in {1,2,4} re-runs just that kernel on them; ``phases | 8`` /
    ``phases | 16`` / ``phases | 32`` force the tiled / the fp32 streaming / the packed streaming round-1 kernel
    (identical output, include/kb_b200.h).

    Like the reference, any ``top_k`` is accepted: 0 gives no rows, and a ``top_k`` the NMS can never exceed (at least
    ``nms_keep_bound``) means raster order for every map (extracter.py:217).  Only SORT_CAP < top_k < keep bound -- a
    sort of more than 8192 survivors -- is not implemented and raises."""
    _require_cuda(score, 'score')
    if params is None:
        nms_dist, threshold, border_dist, top_k, min_score = 4, 0.0, 8, 300, 0.0   # extracter.py:200-205
    else:
        nms_dist, threshold, border_dist = params['nms_dist'], params['threshold'], params['border_dist']
        top_k, min_score = params['top_k'], params['min_score']
    s = _maps3(score)
    b, h, w = s.shape
    if top_k <= 0 and state is None:                           # pts[argsort(...)[:0]] -- an empty result for every map
        z = torch.zeros(b, dtype=torch.int32, device=s.device)
        return (torch.zeros(b, 0, 3, dtype=torch.float32, device=s.device), z,
                torch.zeros(b, 0, dtype=torch.int32, device=s.device), z.clone())
    if top_k > SORT_CAP and state is None:
        bound = nms_keep_bound(h, w, int(nms_dist)) if threshold >= 0 else h * w
        if top_k < bound:
            raise _lib.KbError(f'top_k = {top_k}: sorting more than {SORT_CAP} survivors per map is not implemented '
                               f'(top_k >= {bound}, the most this NMS can keep on a {h}x{w} map, is accepted)')
        # the top-k can never bind: round-faithful NMS, then every surviving pixel in raster order
        nms = fast_nms_batched(s, int(nms_dist)) if nms_dist > 0 else s
        xyp, count, raster, _ = select_batched(nms, int(border_dist), float(threshold), float(min_score), 0, cap=bound)
        return xyp, count, raster, torch.full_like(count, 2)
    if b > DETECT_MAX_B and state is None:
        parts = [detect_batched(s[i:i + DETECT_MAX_B], params, phases) for i in range(0, b, DETECT_MAX_B)]
        return tuple(torch.cat([p[k] for p in parts]) for k in range(4))
    if state:
        xyp, raster, count, path, ws = state
    else:
        xyp = _out(b, top_k, 3, dtype=torch.float32, device=s.device)
        raster = _out(b, top_k, dtype=torch.int32, device=s.device)
        count = _out(b, dtype=torch.int32, device=s.device)         # assigned for every map by the kernels
        path = _out(b, dtype=torch.int32, device=s.device)
        ws = _ws(lib.kb_detect_workspace_bytes(b, h, w, int(nms_dist), int(top_k), float(threshold)), s.device)
        if state is not None:
            state.extend([xyp, raster, count, path, ws])
    with torch.cuda.device(s.device):
        check(lib.kb_detect_phases(s.data_ptr(), b, h, w, int(nms_dist), int(border_dist), float(threshold),
                                   float(min_score), int(top_k), xyp.data_ptr(), raster.data_ptr(), count.data_ptr(),
                                   path.data_ptr(), ws.data_ptr(), ws.numel(), int(phases), _stream()), 'kb_detect')
    _count(5 if (phases & 7) == 7 else 1)       # threshold estimate, round-1, resolve, (fallback NMS, fallback select: early exit)
    return xyp, count, raster, path


# ------------------------------------------------------------------------------------------------
# Stage 2
# ------------------------------------------------------------------------------------------------

def sample_batched(desc: torch.Tensor, pts: torch.Tensor, count: torch.Tensor | None = None, normalize: bool = False,
                   coord_mode: int = 0, s: int = 8) -> torch.Tensor:
    """Bilinear descriptor sampling (utils/matcher.py:221-226; models/lightglue.py:24-41 with
    ``normalize=True, coord_mode=1``).  desc [B,C,h,w]; pts [B,n,>=2]  ->  [B,n,C]."""
    _require_cuda(desc, 'desc')
    _require_cuda(pts, 'pts')
    d = _f32(desc)
    p = _f32(pts)
    b, c, h, w = d.shape
    if p.dim() != 3 or p.shape[0] != b or p.shape[2] < 2:
        raise _lib.KbError(f'pts must be [B,n,>=2] with B={b}, got {tuple(p.shape)}')
    n = p.shape[1]
    out = _out(b, max(n, 1), c, dtype=torch.float32, device=d.device)
    if n == 0:
        return out[:, :0]
    cnt = _i32(count)
    with torch.cuda.device(d.device):
        check(lib.kb_sample_desc(d.data_ptr(), b, c, h, w, p.data_ptr(), p.shape[2], _ptr(cnt), n, int(bool(normalize)),
                                 int(coord_mode), int(s), out.data_ptr(), _stream()), 'kb_sample_desc')
    _count(1)
    return out


# ------------------------------------------------------------------------------------------------
# Stage 3
# ------------------------------------------------------------------------------------------------

def lk_track_batched(img0: torch.Tensor, img1: torch.Tensor, pts0_px: torch.Tensor, init_px: torch.Tensor,
                     count: torch.Tensor | None = None, win_size: int = 3, levels: int = 1, iterations: int = 40):
    """Pyramidal Gauss-Newton patch tracking of ``OpticalFlow`` (utils/matcher.py:49-142) for B image pairs.
    img [B,C,H,W]; pts0_px / init_px [B,n,2] in pixels -> tracked points [B,n,2] in pixels."""
    _require_cuda(img0, 'img0')
    _require_cuda(img1, 'img1')
    a, bm = _f32(img0), _f32(img1)
    if a.dim() != 4 or a.shape != bm.shape:
        raise ValueError(f'images must be two [B,C,H,W] tensors of one shape, got {tuple(img0.shape)} / {tuple(img1.shape)}')
    b, c, h, w = a.shape
    p0, pi = _f32(pts0_px), _f32(init_px)
    if p0.shape != pi.shape or p0.dim() != 3 or p0.shape[0] != b or p0.shape[2] != 2:
        raise ValueError('pts0_px / init_px must both be [B,n,2]')
    n = p0.shape[1]
    out = _out(b, max(n, 1), 2, dtype=torch.float32, device=a.device)[:, :n]
    if n == 0:
        return out
    nbytes = lib.kb_lk_workspace_bytes(b, c, h, w, int(levels))
    if nbytes == 0:
        raise _lib.KbError(f'kb_lk_track: unsupported pyramid ({levels} levels of a {h}x{w} image)')
    ws = _ws(nbytes, a.device)
    cnt = _i32(count)
    with torch.cuda.device(a.device):
        check(lib.kb_lk_track(a.data_ptr(), bm.data_ptr(), b, c, h, w, p0.data_ptr(), pi.data_ptr(), _ptr(cnt), n,
                              int(win_size), int(levels), int(iterations), out.data_ptr(), ws.data_ptr(), ws.numel(),
                              _stream()), 'kb_lk_track')
    _count(4 * int(levels) - 1)
    return out


def match_batched(d0: torch.Tensor, d1: torch.Tensor, n0: torch.Tensor | None = None, n1: torch.Tensor | None = None,
                  max_distance: float = math.inf, cross_check: bool = True, algo: int = -1, return_ws: bool = False,
                  phases: int = 7, state=None, want_dist: bool = True):
    """Mutual-NN matching (utils/matcher.py:227-234).  d0 [B,n,D], d1 [B,m,D]
    -> pairs[B,n,2] int32 (sorted by first index), dist[B,n] float64 (None with ``want_dist=False``: the
    tensor-core path then evaluates float64 distances only where the max_distance gate needs them), count[B]."""
    _require_cuda(d0, 'd0')
    _require_cuda(d1, 'd1')
    a, bm = _f32(d0), _f32(d1)
    b, n, dd = a.shape
    m = bm.shape[1]
    if bm.shape[0] != b or bm.shape[2] != dd:
        raise ValueError('Descriptor length must equal.')
    if state:           # measurement hook (bench.py): re-run one part on the buffers of an earlier full call
        pairs, dist, count, ws = state
    else:
        pairs = _out(b, max(n, 1), 2, dtype=torch.int32, device=a.device)
        dist = _out(b, max(n, 1), dtype=torch.float64, device=a.device) if want_dist else None
        count = _out(b, dtype=torch.int32, device=a.device)         # assigned for every pair by the compaction kernel
        ws = None
    if n == 0 or m == 0:
        return pairs[:, :n], (dist[:, :n] if dist is not None else None), torch.zeros_like(count)
    c0, c1 = _i32(n0), _i32(n1)
    if ws is None:
        ws = _ws(lib.kb_match_workspace_bytes(b, n, m, dd, int(algo)), a.device)
        if state is not None:
            state.extend([pairs, dist, count, ws])
    with torch.cuda.device(a.device):
        check(lib.kb_match_mnn_phases(a.data_ptr(), bm.data_ptr(), _ptr(c0), _ptr(c1), b, n, m, dd, float(max_distance),
                                      int(bool(cross_check)), int(algo), pairs.data_ptr(), _ptr(dist), count.data_ptr(),
                                      ws.data_ptr(), ws.numel(), int(phases), _stream()), 'kb_match_mnn')
    _count(2 if (algo == 0 or dd > 256) else (8 if phases == 7 else 1))   # prep, search, resolve, rescan, gate, rescan, gate, pairs
    if return_ws:
        return pairs, dist, count, ws
    return pairs, dist, count


def sample_match_batched(desc: torch.Tensor, pts: torch.Tensor, count: torch.Tensor | None, pairs: int,
                         max_distance: float = math.inf, cross_check: bool = True, algo: int = -1, want_dist: bool = True,
                         fused: bool | None = None, state=None, part: str | None = None):
    """``sample_batched`` for the 2 * pairs maps of a batch of image pairs (first ``pairs`` maps = image 0) followed by
    ``match_batched(d[:pairs], d[pairs:], count[:pairs], count[pairs:])`` -- what brute_force_matcher does per pair
    (utils/matcher.py:221-234).  Where the library supports it (low-resolution maps sampled densely, C a multiple of 64 up
    to 256, at most 1024 keypoints per map) the sampler writes the tensor-core matcher's operand rows itself
    (kb_sample_desc_operands) and the matcher skips its preparation pass; otherwise the two calls are made as they are.
    ``fused=True`` asks for the fused form (an error where unsupported); the default is the two calls, which measured
    faster on B200.  -> (d [2P,n,C], pairs, dist, count).
    ``state`` / ``part`` are a measurement hook (bench.py, fused form only): ``state=[]`` receives the buffers of a full
    call; a later call with the same ``state`` and ``part`` in {'sample', 'finish', 'search', 'tail'} re-runs that part."""
    _require_cuda(desc, 'desc')
    _require_cuda(pts, 'pts')
    d = _f32(desc)
    p = _f32(pts)
    b, c, h, w = d.shape
    n = p.shape[1] if p.dim() == 3 else 0
    P = int(pairs)
    ok = (p.dim() == 3 and b == 2 * P and p.shape[0] == b and p.shape[2] >= 2 and n > 0 and algo in (-1, 1)
          and bool(lib.kb_sample_desc_operands_supported(c, h, w, n)))
    if fused is True and not ok:
        raise _lib.KbError('kb_sample_desc_operands: unsupported shape')
    if not ok or not fused:                 # (measured slower than the two calls on B200, DESIGN.md section 5: opt-in)
        dd = sample_batched(desc, pts, count)
        cnt = _i32(count)
        pr, dist, cm = match_batched(dd[:P], dd[P:], None if cnt is None else cnt[:P], None if cnt is None else cnt[P:],
                                     max_distance, cross_check, algo=algo, want_dist=want_dist)
        return dd, pr, dist, cm
    cnt = _i32(count)
    if state:
        out, mpairs, dist, mcount, ws = state
    else:
        out = _out(b, n, c, dtype=torch.float32, device=d.device)
        ws = _ws(lib.kb_match_workspace_bytes(P, n, n, c, 1), d.device)
        mpairs = _out(P, n, 2, dtype=torch.int32, device=d.device)
        dist = _out(P, n, dtype=torch.float64, device=d.device) if want_dist else None
        mcount = _out(P, dtype=torch.int32, device=d.device)
        if state is not None:
            state.extend([out, mpairs, dist, mcount, ws])
    c0 = None if cnt is None else cnt[:P]
    c1 = None if cnt is None else cnt[P:]
    phases = {None: 6 | 8, 'finish': 8, 'search': 2, 'tail': 4}.get(part)
    with torch.cuda.device(d.device):
        if part in (None, 'sample'):
            check(lib.kb_sample_desc_operands(d.data_ptr(), P, c, h, w, p.data_ptr(), p.shape[2], _ptr(cnt), n, 0, 8,
                                              out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), 'kb_sample_desc_operands')
        if phases is not None:
            check(lib.kb_match_mnn_phases(out[:P].data_ptr(), out[P:].data_ptr(), _ptr(c0), _ptr(c1), P, n, n, c,
                                          float(max_distance), int(bool(cross_check)), 1, mpairs.data_ptr(), _ptr(dist),
                                          mcount.data_ptr(), ws.data_ptr(), ws.numel(), phases, _stream()), 'kb_match_mnn')
    _count(1 + 8 if part is None else 1)        # fused sampler; finish, search, resolve, rescan, gate, pairs (+ the zeroing memset)
    return out, mpairs, dist, mcount


def _knob(k: int) -> int:
    v = lib.kb_debug_knob(k, 0)
    lib.kb_debug_knob(k, v)
    return v


def match_issue_factor(cross_check: bool, dim: int = 256) -> int:
    """How many times the one-pass 2*n*m*D flops of a pair the tensor-core search issues (bench.py reports it beside the
    algorithmic figure): two fp16 products (hi + lo of the query against the fp16 rounding of the database; three bf16
    products under KB_KNOB_TC_BF16X3) per direction, one Gram per direction (one direction only under the one-pass
    cross-check, KB_KNOB_TC_ONE_PASS)."""
    k = _knob(_lib.KB_KNOB_TC_BF16X3)
    products = 3 if k == 1 else (2 if k == 2 else (2 if dim > 64 else 3))
    return products * (2 if (cross_check and not _knob(_lib.KB_KNOB_TC_ONE_PASS)) else 1)


def match_tc_debug(ws: torch.Tensor, b: int, n: int, m: int, dd: int) -> dict:
    """Views into an algo=1 workspace (tests only): per-row top-2 records, exact-rescan count, norms."""
    import ctypes
    off = (ctypes.c_size_t * 6)()
    check(lib.kb_match_tc_debug_offsets(b, n, m, dd, ctypes.cast(off, ctypes.c_void_p)), 'kb_match_tc_debug_offsets')
    def rec(o, rows, slices=4):
        # per row: `slices` records of 32 bytes (float best, second, third, pad; int argbest, argsecond, pad, pad),
        # one per column slice of the epilogue; the resolver merges them -- here: best over the slices
        raw = ws[o:o + rows * slices * 32].view(torch.float32).reshape(rows, slices, 8)
        idx = ws[o:o + rows * slices * 32].view(torch.int32).reshape(rows, slices, 8)[:, :, 4]
        best, arg = raw[:, :, 0].max(dim=1)
        second = torch.where(torch.arange(slices, device=ws.device)[None, :] == arg[:, None],
                             raw[:, :, 1], raw[:, :, 0]).max(dim=1).values
        return best, second, idx.gather(1, arg[:, None])[:, 0]
    return {'res0': rec(off[0], b * n), 'res1': rec(off[1], b * m),
            'n_exact': ws[off[2]:off[2] + 4].view(torch.int32), 'n_pair': ws[off[2] + 4:off[2] + 8].view(torch.int32),
            'n_col_rescan': ws[off[2] + 8:off[2] + 12].view(torch.int32),
            'norm2_0': ws[off[3]:off[3] + 4 * b * n].view(torch.float32),
            'norm2_1': ws[off[4]:off[4] + 4 * b * m].view(torch.float32)}


# ------------------------------------------------------------------------------------------------
# Stage 4
# ------------------------------------------------------------------------------------------------

def warp_batched(pts: torch.Tensor, count: torch.Tensor | None, h33: torch.Tensor, wh: torch.Tensor):
    """``warp_homography`` (utils/projection.py:137-167) per map.  pts [B,n,>=2]; h33 [B,3,3] or [B,9];
    wh [B,2] = (width, height).  -> kp_valid[B,n,2], kp_warp[B,n,2], ids[B,n], ids_out[B,n], n_valid[B]."""
    _require_cuda(pts, 'pts')
    p = _f32(pts)
    b, n = p.shape[0], p.shape[1]
    hm = _f32(h33).reshape(b, 9)
    whf = _f32(wh).reshape(b, 2)
    kv = _out(b, max(n, 1), 2, dtype=torch.float32, device=p.device)
    kw = _out(b, max(n, 1), 2, dtype=torch.float32, device=p.device)
    ids = _out(b, max(n, 1), dtype=torch.int32, device=p.device)
    ids_out = _out(b, max(n, 1), dtype=torch.int32, device=p.device)
    if n == 0:
        return kv[:, :0], kw[:, :0], ids[:, :0], ids_out[:, :0], torch.zeros(b, dtype=torch.int32, device=p.device)
    nv = _out(b, dtype=torch.int32, device=p.device)                # assigned for every map by the kernel
    cnt = _i32(count)
    with torch.cuda.device(p.device):
        check(lib.kb_warp_homography(p.data_ptr(), p.shape[2], _ptr(cnt), b, n, hm.data_ptr(), whf.data_ptr(),
                                     kv.data_ptr(), kw.data_ptr(), ids.data_ptr(), ids_out.data_ptr(), nv.data_ptr(),
                                     _stream()), 'kb_warp_homography')
    _count(1)
    return kv, kw, ids, ids_out, nv


def warp_se3_batched(pts: torch.Tensor, count: torch.Tensor | None, depth0: torch.Tensor, depth1: torch.Tensor,
                     k0: torch.Tensor, k1: torch.Tensor, pose01: torch.Tensor, bbox0: torch.Tensor, bbox1: torch.Tensor):
    """``warp_se3`` (utils/projection.py:194-267) per map.  pts [B,n,>=2]; depth0 [B,h0,w0]; depth1 [B,h1,w1];
    k0 / k1 [B,3,3] intrinsics; pose01 [B,4,4]; bbox0 / bbox1 [B,2] (row, col).
    -> kp_valid[B,n,2], kp_warp[B,n,2], ids[B,n], ids_out[B,n], n_valid[B], n_out[B]."""
    _require_cuda(pts, 'pts')
    p = _f32(pts)
    b, n = p.shape[0], p.shape[1]
    d0, d1 = _f32(depth0).reshape(b, depth0.shape[-2], depth0.shape[-1]), _f32(depth1).reshape(b, depth1.shape[-2], depth1.shape[-1])
    # the reference inverts intrinsics0 in float32 on the fly (torch.inverse, projection.py:46): same precision here
    kinv = torch.linalg.inv(k0.to(device=p.device, dtype=torch.float32)).reshape(b, 9).contiguous()
    k1f = _f32(k1.to(p.device)).reshape(b, 9)
    pose = _f32(pose01.to(p.device)).reshape(b, 16)
    bb0, bb1 = _f32(bbox0.to(p.device)).reshape(b, 2), _f32(bbox1.to(p.device)).reshape(b, 2)
    kv = _out(b, max(n, 1), 2, dtype=torch.float32, device=p.device)
    kw = _out(b, max(n, 1), 2, dtype=torch.float32, device=p.device)
    ids = _out(b, max(n, 1), dtype=torch.int32, device=p.device)
    ids_out = _out(b, max(n, 1), dtype=torch.int32, device=p.device)
    nv = torch.zeros(b, dtype=torch.int32, device=p.device)
    no = torch.zeros(b, dtype=torch.int32, device=p.device)
    if n == 0:
        return kv[:, :0], kw[:, :0], ids[:, :0], ids_out[:, :0], nv, no
    cnt = _i32(count)
    with torch.cuda.device(p.device):
        check(lib.kb_warp_se3(p.data_ptr(), p.shape[2], _ptr(cnt), b, n, d0.data_ptr(), d0.shape[1], d0.shape[2], d1.data_ptr(),
                              d1.shape[1], d1.shape[2], kinv.data_ptr(), k1f.data_ptr(), pose.data_ptr(), bb0.data_ptr(),
                              bb1.data_ptr(), kv.data_ptr(), kw.data_ptr(), ids.data_ptr(), ids_out.data_ptr(), nv.data_ptr(),
                              no.data_ptr(), _stream()), 'kb_warp_se3')
    _count(1)
    return kv, kw, ids, ids_out, nv, no


def repeat_batched(k0c, k01c, na, k1c, k10c, nb, scale01: float, scale10: float, th: float, want_errors: bool = True,
                   pair_cap: int = 0):
    """Counting core of ``val_key_points`` (tasks/repeatability.py:69-85).
    -> stats[B,4] float64 (gt_num, sum of errors <= th, n mutual pairs, 0), errors[B,a] or None, pairs or None."""
    for t, nm in ((k0c, 'k0c'), (k01c, 'k01c'), (k1c, 'k1c'), (k10c, 'k10c')):
        _require_cuda(t, nm)
    a0, a1, b0, b1 = _f32(k0c), _f32(k01c), _f32(k1c), _f32(k10c)
    b, a_max, b_max = a0.shape[0], a0.shape[1], b0.shape[1]
    pairs = torch.full((b, pair_cap, 2), -1, dtype=torch.int32, device=a0.device) if pair_cap > 0 else None
    if a_max == 0 or b_max == 0:
        return (torch.zeros(b, 4, dtype=torch.float64, device=a0.device),
                torch.zeros(b, a_max, dtype=torch.float32, device=a0.device) if want_errors else None, pairs)
    stats = _out(b, 4, dtype=torch.float64, device=a0.device)           # both are cleared by the library's init kernel
    errors = _out(b, a_max, dtype=torch.float32, device=a0.device) if want_errors else None
    ca, cb = _i32(na), _i32(nb)
    ws = _ws(lib.kb_repeat_workspace_bytes(b, a_max, b_max), a0.device)
    with torch.cuda.device(a0.device):
        check(lib.kb_repeat_counts(a0.data_ptr(), a1.data_ptr(), _ptr(ca), b0.data_ptr(), b1.data_ptr(), _ptr(cb), b,
                                   a_max, b_max, float(scale01), float(scale10), float(th), stats.data_ptr(),
                                   _ptr(errors), _ptr(pairs), int(pair_cap), ws.data_ptr(), ws.numel(), _stream()),
              'kb_repeat_counts')
    _count(8 + (1 if pair_cap > 0 else 0))     # init, bound, sort, 2 sorted minima, sorted mutual, 2 exhaustive (early exit)
    return stats, errors, pairs


def corner_error_batched(h_est: torch.Tensor, h_real: torch.Tensor, valid: torch.Tensor | None, w: int, h: int,
                         resize_h: int, resize_w: int, th) -> tuple[torch.Tensor, torch.Tensor]:
    """MHA corner error (tasks/MHA.py:51-72).  h_est/h_real [B,3,3] -> mean_dist[B], flags[B,len(th)] (float64)."""
    _require_cuda(h_est, 'h_est')
    he = h_est.to(torch.float64).contiguous().reshape(-1, 9)
    hr = h_real.to(device=he.device, dtype=torch.float64).contiguous().reshape(-1, 9)
    b = he.shape[0]
    tht = torch.as_tensor(list(th), dtype=torch.float64, device=he.device)
    md = torch.zeros(b, dtype=torch.float64, device=he.device)
    flags = torch.zeros(b, tht.numel(), dtype=torch.float64, device=he.device)
    v = _i32(valid)
    with torch.cuda.device(he.device):
        check(lib.kb_corner_error(he.data_ptr(), hr.data_ptr(), _ptr(v), b, int(w), int(h), int(resize_h),
                                  int(resize_w), tht.data_ptr(), tht.numel(), md.data_ptr(), flags.data_ptr(),
                                  _stream()), 'kb_corner_error')
    _count(1)
    return md, flags
