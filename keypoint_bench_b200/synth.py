"""Seeded synthetic inputs of the shapes the CNN backbones hand to the post-network path.

The backbones themselves (ALIKE, SuperPoint, XFeat, DISK ...) are out of scope; what
matters is the shape / value range of their outputs (e.g.
models/ALike.py:159-164, models/SuperPoint.py:61-71, models/XFeat.py:136-140,
models/disk.py:309-313) and the ``warp01_params`` schema of datasets/hpatches.py:76-81.

Everything is generated from ``torch.Generator`` seeds so the oracle, the tests and the
benchmark see identical inputs.  CPU generation is bit-reproducible across hosts for the
iid kinds ('uniform', 'ties', 'ramp', 'negative'); the blurred 'alike' kind goes through a
convolution and is only used where inputs are copied, not regenerated.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class PathConfig:
    """One of BASELINE.json's five configs, reduced to what the hot path sees."""
    name: str
    height: int
    width: int
    desc_dim: int          # 0 = no descriptor stage (repeatability only)
    desc_stride: int       # descriptor map is (H/stride, W/stride)
    desc_normalized: bool
    nms_dist: int
    top_k: int
    border_dist: int = 8
    threshold: float = 0.0
    min_score: float = 0.0
    max_distance: float = 5.0
    cross_check: bool = True
    task: str = "match"    # 'repeatability' | 'mha' | 'match' | 'stream'
    pairs_per_gpu: int = 64

    @property
    def extractor_params(self) -> dict:
        return {'nms_dist': self.nms_dist, 'threshold': self.threshold, 'border_dist': self.border_dist,
                'top_k': self.top_k, 'min_score': self.min_score}

    @property
    def matcher_params(self) -> dict:
        return {'metric': 'euclidean', 'max_distance': self.max_distance, 'cross_check': self.cross_check}


# config/config.yaml:17-22,38 ; config/config_MHA.yaml:82-85 ; SURVEY.md section 8(d)
CONFIGS = {
    'cfg1': PathConfig('cfg1-alike-t-repeatability-480x640', 480, 640, 0, 1, False, 6, 1000, task='repeatability'),
    'cfg2': PathConfig('cfg2-superpoint256-mha-480x640', 480, 640, 256, 8, True, 6, 1000, task='mha'),
    # nms_dist=4 (config_vo.yaml:88) so that top_k=4096 binds on a 480x640 map (SURVEY 8(d) warning)
    'cfg3': PathConfig('cfg3-xfeat64-top4096-480x640', 480, 640, 64, 8, True, 4, 4096, task='match'),
    'cfg4': PathConfig('cfg4-disk128-top2048-1024x1024', 1024, 1024, 128, 1, True, 6, 2048, task='match',
                       pairs_per_gpu=4),
    'cfg5': PathConfig('cfg5-alike-t-stream-376x1241', 376, 1241, 64, 1, False, 6, 1000, task='stream',
                       pairs_per_gpu=32),
}


def pair_seed(cfg_index: int, pair_idx: int) -> int:
    return 1234 + 1000 * cfg_index + pair_idx


def _gen(seed: int, device="cpu") -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _gauss_kernel(sigma: float, device) -> torch.Tensor:
    r = int(math.ceil(3 * sigma))
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
    k = torch.exp(-(x * x) / (2 * sigma * sigma))
    return k / k.sum()


def score_map(kind: str, h: int, w: int, seed: int, device="cpu", sigma: float = 2.0) -> torch.Tensor:
    """[1,1,h,w] float32 score map.

    'uniform'  U(0,1) iid (6-8 NMS rounds)              'alike'   sigmoid(2*blur(N(0,1))/std)
    'ties'     floor(8U)/8 (heavy ties)                 'ramp'    increasing along x (W/r rounds)
    'negative' U(-1,0)                                  'mixed'   U(-0.5,0.5)
    'relu'     max(U(-1,1),0) (exact zeros, KeyNet-like)
    """
    g = _gen(seed, device)
    if kind == 'uniform':
        return torch.rand(1, 1, h, w, generator=g, device=device)
    if kind == 'ties':
        return torch.floor(torch.rand(1, 1, h, w, generator=g, device=device) * 8) / 8
    if kind == 'ramp':
        row = (torch.arange(w, dtype=torch.float32, device=device) + 1) / (w + 1)
        return row.view(1, 1, 1, w).expand(1, 1, h, w).contiguous()
    if kind == 'negative':
        return torch.rand(1, 1, h, w, generator=g, device=device) - 1.0
    if kind == 'mixed':
        return torch.rand(1, 1, h, w, generator=g, device=device) - 0.5
    if kind == 'relu':
        return torch.clamp(torch.rand(1, 1, h, w, generator=g, device=device) * 2 - 1, min=0)
    if kind == 'alike':
        z = torch.randn(1, 1, h, w, generator=g, device=device)
        k = _gauss_kernel(sigma, device)
        r = k.numel() // 2
        z = F.conv2d(F.pad(z, (r, r, 0, 0), mode='reflect'), k.view(1, 1, 1, -1))
        z = F.conv2d(F.pad(z, (0, 0, r, r), mode='reflect'), k.view(1, 1, -1, 1))
        return torch.sigmoid(2 * z / z.std())
    raise ValueError(f'unknown score-map kind {kind!r}')


def homography(seed: int, jitter: float = 0.10) -> torch.Tensor:
    """Pixel-unit 3x3 homography near identity, jittered +-``jitter`` per pair."""
    g = _gen(seed)
    base = torch.tensor([[1.02, 0.03, 5.0], [-0.02, 0.98, 3.0], [1e-5, -2e-5, 1.0]], dtype=torch.float64)
    ident = torch.eye(3, dtype=torch.float64)
    j = 1 + jitter * (2 * torch.rand(3, 3, generator=g, dtype=torch.float64) - 1)
    hm = ident + (base - ident) * j
    hm[2, 2] = 1.0
    return hm.to(torch.float32)


def _inverse_grid(hm: torch.Tensor, h: int, w: int, device) -> torch.Tensor:
    """Normalised sampling grid that pulls image B's pixel (u,v) from A at H^-1 (u,v)."""
    hinv = torch.linalg.inv(hm.double()).to(device)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64, device=device),
                            torch.arange(w, dtype=torch.float64, device=device), indexing='ij')
    q = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1) @ hinv.T
    q = q[..., :2] / q[..., 2:]
    gx = q[..., 0] / (w - 1) * 2 - 1
    gy = q[..., 1] / (h - 1) * 2 - 1
    return torch.stack([gx, gy], dim=-1).float().unsqueeze(0)


def warp_map(src: torch.Tensor, hm: torch.Tensor, mode: str = 'nearest') -> torch.Tensor:
    """Image-B view of a [1,C,h,w] map under pixel homography ``hm`` (of ITS OWN resolution)."""
    _, _, h, w = src.shape
    grid = _inverse_grid(hm, h, w, src.device)
    return F.grid_sample(src, grid, mode=mode, padding_mode='zeros', align_corners=True)


def rescale_homography(hm: torch.Tensor, stride: int) -> torch.Tensor:
    """Same homography expressed in the pixel units of a 1/stride-resolution map."""
    if stride == 1:
        return hm
    s = torch.diag(torch.tensor([1.0 / stride, 1.0 / stride, 1.0], dtype=torch.float64))
    return (s @ hm.double() @ torch.linalg.inv(s)).float()


def desc_map(c: int, h: int, w: int, seed: int, normalized: bool, device="cpu") -> torch.Tensor:
    """[1,c,h,w]: unit-norm per pixel (SuperPoint/XFeat/DISK) or raw ~2.67-norm (ALIKE)."""
    g = _gen(seed, device)
    d = torch.randn(1, c, h, w, generator=g, device=device)
    d = F.normalize(d, dim=1)
    return d if normalized else 2.67 * d


def warp_params(hm: torch.Tensor, h: int, w: int, resize: int = 512, as_tensors: bool = True):
    """warp01 / warp10 dicts with the schema of datasets/hpatches.py:76-81 after the collate
    unwrapping of models/model_interface.py:176-180 (0-dim int64 tensors)."""
    def wrap(v):
        return torch.tensor(v, dtype=torch.int64) if as_tensors else int(v)
    inv = torch.linalg.inv(hm.double()).float()
    w01 = {'mode': 'homo', 'width': wrap(w), 'height': wrap(h), 'homography_matrix': hm.clone(), 'resize': wrap(resize)}
    w10 = {'mode': 'homo', 'width': wrap(w), 'height': wrap(h), 'homography_matrix': inv, 'resize': wrap(resize)}
    return w01, w10


def make_pair(cfg: PathConfig, cfg_index: int, pair_idx: int, kind: str = 'uniform', device="cpu") -> dict:
    """One synthetic image pair for ``cfg``: score maps, descriptor maps, warp dicts."""
    seed = pair_seed(cfg_index, pair_idx)
    h, w = cfg.height, cfg.width
    hm = homography(seed + 7)
    s0 = score_map(kind, h, w, seed, device)
    s1 = warp_map(s0, hm.to(device), 'nearest')
    out = {'score0': s0, 'score1': s1, 'H': hm}
    out['warp01'], out['warp10'] = warp_params(hm, h, w)
    if cfg.desc_dim:
        dh, dw = h // cfg.desc_stride, w // cfg.desc_stride
        d0 = desc_map(cfg.desc_dim, dh, dw, seed + 13, cfg.desc_normalized, device)
        hd = rescale_homography(hm, cfg.desc_stride).to(device)
        g = _gen(seed + 17, device)
        d1 = warp_map(d0, hd, 'bilinear') + 0.05 * torch.randn(d0.shape, generator=g, device=device)
        out['desc0'], out['desc1'] = d0, d1
    return out


def se3_scene(h: int, w: int, seed: int, device="cpu") -> dict:
    """Synthetic two-view depth scene with the ``warp01_params`` schema of datasets/megadepth.py:333-352
    (mode 'se3'): a smooth depth map for view 0, a small rigid motion, and view 1's depth rendered by
    forward-warping view 0 (nearest pixel, nearest depth wins) -- which leaves holes (depth 0), the
    'no depth' branch of utils/projection.py:interpolate_depth -- plus a closer occluder patch."""
    g = _gen(seed)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing='ij')
    depth0 = 4.0 + 0.8 * torch.sin(xs / 37.0) * torch.cos(ys / 29.0) + 0.002 * xs
    holes = torch.rand(h, w, generator=g, dtype=torch.float64) < 0.03
    depth0 = torch.where(holes, torch.zeros_like(depth0), depth0)
    f = 0.9 * w
    k0 = torch.tensor([[f, 0.0, w / 2.0 + 3.0], [0.0, f * 1.02, h / 2.0 - 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    k1 = torch.tensor([[f * 0.97, 0.0, w / 2.0 - 4.0], [0.0, f, h / 2.0 + 1.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    ang = torch.tensor([0.02, -0.03, 0.015], dtype=torch.float64) * (1 + 0.2 * (2 * torch.rand(3, generator=g, dtype=torch.float64) - 1))
    cx, cy, cz = torch.cos(ang)
    sx, sy, sz = torch.sin(ang)
    rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float64)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
    rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=torch.float64)
    pose = torch.eye(4, dtype=torch.float64)
    pose[:3, :3] = rz @ ry @ rx
    pose[:3, 3] = torch.tensor([0.25, -0.1, 0.15], dtype=torch.float64)
    bbox0 = torch.tensor([3.0, 5.0])          # (row, col) offsets of the crops, COLMAP convention
    bbox1 = torch.tensor([2.0, 7.0])
    # render depth1: unproject every valid pixel of view 0, move it, project, keep the nearest depth
    valid = depth0 > 0
    u = xs[valid] + bbox0[1].double() + 0.5
    v = ys[valid] + bbox0[0].double() + 0.5
    d = depth0[valid]
    pts = torch.linalg.inv(k0) @ torch.stack([u * d, v * d, d])
    pts1 = pose[:3, :3] @ pts + pose[:3, 3:4]
    q = k1 @ pts1
    z = q[2]
    px = torch.round(q[0] / z - bbox1[1].double() - 0.5).long()
    py = torch.round(q[1] / z - bbox1[0].double() - 0.5).long()
    ok = (px >= 0) & (px < w) & (py >= 0) & (py < h) & (z > 0)
    depth1 = torch.full((h * w,), float('inf'), dtype=torch.float64)
    depth1.scatter_reduce_(0, (py[ok] * w + px[ok]), z[ok], reduce='amin')
    depth1 = torch.where(torch.isinf(depth1), torch.zeros_like(depth1), depth1).reshape(h, w)
    depth1[h // 3:h // 3 + h // 8, w // 2:w // 2 + w // 6] *= 0.7          # an occluder in view 1
    f32 = lambda t: t.to(torch.float32).to(device)      # noqa: E731
    return {'mode': 'se3', 'width': w, 'height': h, 'pose01': f32(pose), 'bbox0': bbox0.to(device), 'bbox1': bbox1.to(device),
            'depth0': f32(depth0), 'depth1': f32(depth1), 'intrinsics0': f32(k0), 'intrinsics1': f32(k1)}


def lk_scene(c: int, h: int, w: int, seed: int, shift=(2.3, -1.7), sigma: float = 3.0, device="cpu"):
    """Two smooth images for the Lucas-Kanade tracker (utils/matcher.py:7-142): Gaussian-filtered noise scaled to
    [0,1] (an 'image' the tracker can converge on) and the same field translated by ``shift`` pixels
    (img1(x, y) = img0(x - dx, y - dy), bilinear).  -> img0, img1 [1,c,h,w]."""
    g = _gen(seed)
    x = torch.randn(1, c, h, w, generator=g)
    k = int(4 * sigma) | 1
    ax = torch.arange(k, dtype=torch.float32) - k // 2
    ker = torch.exp(-ax ** 2 / (2 * sigma ** 2))
    ker = ker / ker.sum()
    x = torch.nn.functional.conv2d(x, ker.view(1, 1, 1, k).repeat(c, 1, 1, 1), padding=(0, k // 2), groups=c)
    x = torch.nn.functional.conv2d(x, ker.view(1, 1, k, 1).repeat(c, 1, 1, 1), padding=(k // 2, 0), groups=c)
    img0 = (x - x.min()) / (x.max() - x.min())
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    grid = torch.stack([(xs - shift[0]) / (w - 1) * 2 - 1, (ys - shift[1]) / (h - 1) * 2 - 1], dim=-1)[None]
    img1 = torch.nn.functional.grid_sample(img0, grid, align_corners=True, padding_mode='border')
    return img0.to(device), img1.to(device)


# ------------------------------------------------------------------------------------------------
# Batches (what the backbone + dataset hand to the path).  They live here, not in pipeline.py, so that
# code which only needs INPUTS (bench.py's reference arm, the oracle-side tests) never loads the CUDA library.
# ------------------------------------------------------------------------------------------------

@dataclass
class PairBatch:
    """Device-resident inputs for P pairs (what the backbone + dataset would hand over)."""
    score: torch.Tensor          # [2P,1,H,W]   first P = image 0, last P = image 1
    desc: torch.Tensor | None    # [2P,C,h,w]
    h33: torch.Tensor            # [2P,9]  first P = H01 (pixels of image 1), last P = H10
    wh: torch.Tensor             # [2P,2]  (width,height) of the TARGET image of each warp
    resize: int = 512

    @property
    def pairs(self) -> int:
        return self.score.shape[0] // 2


@dataclass
class FrameBatch:
    """Device-resident inputs of a frame stream (KITTI-like sequences, tasks/visual_odometer.py,
    tasks/FundamentalMatrix.py): F + 1 consecutive frames give F pairs (t-1, t)."""
    score: torch.Tensor          # [F+1,1,H,W]
    desc: torch.Tensor           # [F+1,C,h,w]

    @property
    def pairs(self) -> int:
        return self.score.shape[0] - 1


def make_batch(cfg: PathConfig, cfg_index: int, n_pairs: int, first_pair: int, device, kind: str = 'uniform'):
    """P pairs resident on ``device`` per SURVEY 8(d): score A, score B = nearest-warp of A, descriptor map A
    (unit-norm per pixel, or 2.67x for ALIKE-like configs), map B = bilinear warp of A + 0.05 noise.
    -> (PairBatch, homographies [P,3,3])."""
    H, W = cfg.height, cfg.width
    g = torch.Generator(device=device)
    g.manual_seed(pair_seed(cfg_index, first_pair))
    hms = torch.stack([homography(pair_seed(cfg_index, first_pair + i) + 7) for i in range(n_pairs)])
    if kind == 'uniform':
        s0 = torch.rand(n_pairs, 1, H, W, generator=g, device=device)
    else:
        s0 = torch.cat([score_map(kind, H, W, pair_seed(cfg_index, first_pair + i), device) for i in range(n_pairs)])
    s1 = torch.empty_like(s0)
    chunk = 8
    for i in range(0, n_pairs, chunk):
        grids = torch.cat([_inverse_grid(hms[j], H, W, device) for j in range(i, min(i + chunk, n_pairs))])
        s1[i:i + chunk] = F.grid_sample(s0[i:i + chunk], grids, mode='nearest', padding_mode='zeros', align_corners=True)
    desc = None
    if cfg.desc_dim:
        dh, dw = H // cfg.desc_stride, W // cfg.desc_stride
        d0 = F.normalize(torch.randn(n_pairs, cfg.desc_dim, dh, dw, generator=g, device=device), dim=1)
        if not cfg.desc_normalized:
            d0 = 2.67 * d0
        d1 = torch.empty_like(d0)
        for i in range(0, n_pairs, chunk):
            grids = torch.cat([_inverse_grid(rescale_homography(hms[j], cfg.desc_stride), dh, dw, device)
                               for j in range(i, min(i + chunk, n_pairs))])
            d1[i:i + chunk] = F.grid_sample(d0[i:i + chunk], grids, mode='bilinear', padding_mode='zeros', align_corners=True)
        d1 += 0.05 * torch.randn(d1.shape, generator=g, device=device)
        desc = torch.cat([d0, d1])
    h01 = hms.reshape(n_pairs, 9)
    h10 = torch.linalg.inv(hms.double()).float().reshape(n_pairs, 9)
    h33 = torch.cat([h01, h10]).to(device)
    wh = torch.tensor([[float(W), float(H)]], device=device).expand(2 * n_pairs, 2).contiguous()
    return PairBatch(score=torch.cat([s0, s1]), desc=desc, h33=h33, wh=wh, resize=512), hms


def _hash_uniform(idx: torch.Tensor, seed: int) -> torch.Tensor:
    """Counter-based U[0,1): integer mixing of an int64 index tensor (same bits on every device / rank)."""
    def lsr(v, k):                               # logical shift right of an int64 tensor
        return (v >> k) & ((1 << (64 - k)) - 1)
    h = idx * 6364136223846793005 + (1442695040888963407 + 2 * int(seed) + 1)
    h = h ^ lsr(h, 29)
    h = h * -4658895280553007687                 # 0xBF58476D1CE4E5B9 as int64
    h = h ^ lsr(h, 32)
    h = h * -7723592293110705685                 # 0x94D049BB133111EB as int64
    h = h ^ lsr(h, 29)
    return lsr(h, 40).to(torch.float32) / 16777216.0


PAN_PX = 3        # the synthetic camera pans 3 px per frame; new content enters on the left


def make_frames(cfg: PathConfig, seed: int, first_frame: int, n_frames: int, total_frames: int, device) -> FrameBatch:
    """Frames [first_frame, first_frame + n_frames) of a synthetic sequence of ``total_frames`` frames: windows of a
    procedural panorama (a counter-based hash of the panorama pixel, so every rank generates exactly its own frames and
    a frame's content depends only on its global index) plus per-frame descriptor noise."""
    H, W, C = cfg.height, cfg.width, cfg.desc_dim
    dh, dw = H // cfg.desc_stride, W // cfg.desc_stride
    pano_w = W + PAN_PX * (total_frames - 1)
    ys = torch.arange(H, dtype=torch.int64, device=device)[:, None]
    xs = torch.arange(W, dtype=torch.int64, device=device)[None, :]
    score = torch.empty(n_frames, 1, H, W, dtype=torch.float32, device=device)
    desc = torch.empty(n_frames, C, dh, dw, dtype=torch.float32, device=device)
    cs = torch.arange(C, dtype=torch.int64, device=device)[:, None, None]
    dys = torch.arange(dh, dtype=torch.int64, device=device)[None, :, None]
    dxs = torch.arange(dw, dtype=torch.int64, device=device)[None, None, :]
    for i in range(n_frames):
        t = first_frame + i
        off = PAN_PX * (total_frames - 1 - t)                 # panorama column of the frame's column 0
        score[i, 0] = _hash_uniform(ys * pano_w + (xs + off), seed)
        doff = off // cfg.desc_stride
        pidx = (cs * dh + dys) * (pano_w // cfg.desc_stride + 1) + (dxs + doff)
        base = _hash_uniform(pidx, seed + 1) + _hash_uniform(pidx, seed + 2) + _hash_uniform(pidx, seed + 3) - 1.5
        base = F.normalize(base, dim=0)
        fidx = ((cs * dh + dys) * dw + dxs) * total_frames + t
        noise = (_hash_uniform(fidx, seed + 4) + _hash_uniform(fidx, seed + 5) - 1.0) * 0.1225     # std ~0.05
        d = base + noise
        desc[i] = d if cfg.desc_normalized else 2.67 * d
    return FrameBatch(score, desc)
