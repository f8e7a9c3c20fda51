"""Time the ViTPose detector network (device-resident input) per arithmetic class.  usage: python tools/time_vitpose.py [batch] [classes]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vitpose as ov                                             # noqa: E402  (random weights only)
from upliftingtabletennis_b200.vitpose import VitPose                        # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
classes = sys.argv[2].split(',') if len(sys.argv) > 2 else ['bf16', 'tf32x3']
dev = torch.device('cuda:0')
hp, wp = ov.tokens_hw(640, 1152)
sd = ov.random_state_dict(7, 9, hp * wp, 1)
x = torch.from_numpy(np.random.default_rng(1).standard_normal((batch, 9, 640, 1152)).astype(np.float32)).to(dev)
ref = None
for c in classes:
    m = VitPose(in_frames=3, resolution=(1152, 640), dtype=c).to(dev).eval()
    m.load_state_dict(sd, strict=True)
    for _ in range(2):
        y, _n = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        y, _n = m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('%s: %.2f ms per %d stacks = %.0f stacks/s, %.1f TFLOP/s (313.5 GFLOP/stack), launches %d' %
          (c, ms, batch, batch / ms * 1e3, 313.5e9 * batch / ms / 1e9, m.engine.last_launches()))
    if ref is None:
        ref = y.clone()
    else:
        print('   max |y - y[%s]| = %.3e (max |y| %.3f)' % (classes[0], (y - ref).abs().max().item(), ref.abs().max().item()))
