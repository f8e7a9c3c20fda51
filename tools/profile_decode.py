"""Decode + pre-process kernels on inputs larger than L2, for an ncu launch list (tools only)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from upliftingtabletennis_b200 import ops, synthetic  # noqa: E402

dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
hm = torch.randn((64, 704, 1280), device=dev, generator=g) * 0.05     # 231 MB > 126 MB L2
yy, xx = torch.meshgrid(torch.arange(704, device=dev), torch.arange(1280, device=dev), indexing='ij')
for i in range(64):
    hm[i] += torch.exp(-((xx - (100.3 + 17 * i)) ** 2 + (yy - (50.7 + 9 * i)) ** 2) / (2 * 1.5 ** 2))
frames = torch.from_numpy(synthetic.frames_1080p(34, seed=1)).to(dev)
for _ in range(3):
    pos = ops.decode_heatmaps(hm, 1920, 1080, 'table')
    x = ops.preprocess_stacks(frames, 3, 1, 32, 1280, 704, layout='nhwc16', dtype=torch.bfloat16)
torch.cuda.synchronize()
print('ok', pos[0].tolist())
