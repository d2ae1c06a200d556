"""One call each of simple_nms (128 maps, r = 5) and the Lucas-Kanade tracker (1 pair, 1000 keypoints, win 21,
3 levels x 40 iterations) -- driven under ncu for the --set full captures of the SURVEY 8(f) kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keypoint_bench_b200 import ops, synth  # noqa: E402

g = torch.Generator(device='cuda').manual_seed(1)
score = torch.rand(128, 1, 480, 640, generator=g, device='cuda')
out = ops.simple_nms_batched(score, 5)
img0, img1 = synth.lk_scene(3, 480, 640, 40, shift=(2.0, -1.5))
img0, img1 = img0.cuda(), img1.cuda()
pts = torch.rand(1, 1000, 2, generator=g, device='cuda') * torch.tensor([639.0, 479.0], device='cuda')
init = pts + torch.randn(1, 1000, 2, generator=g, device='cuda') * 3
trk = ops.lk_track_batched(img0, img1, pts, init, None, 21, 3, 40)
torch.cuda.synchronize()
print(float(out.sum()), float(trk.sum()))
