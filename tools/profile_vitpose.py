"""Time the ViTPose detector per kernel family (development aid; numbers under a profiler are not bench values).
    python tools/profile_vitpose.py [batch] [subbatch]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vitpose as ov
from upliftingtabletennis_b200._lib import lib, check
from upliftingtabletennis_b200.vitpose import VitPose

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device('cuda')
m = VitPose(in_frames=3, resolution=(1152, 640)).to(dev).eval()
m.load_state_dict(ov.random_state_dict(7, 9, 2880, 1), strict=True)
check(lib.ttk_vit_set_subbatch(m.engine.h, sub))
x = torch.randn(batch, 9, 640, 1152, device=dev)
for dt in (torch.bfloat16, torch.float32):
    m.compute_dtype = dt
    n = 3 if dt == torch.bfloat16 else 1
    m(x[:sub])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print('%s: %.2f ms for %d stacks = %.1f stacks/s, %.1f TFLOP/s (313.5 GFLOP/stack), %d launches' %
          (dt, ms, batch, batch / ms * 1e3, 313.5 * batch / ms, m.engine.last_launches()))
